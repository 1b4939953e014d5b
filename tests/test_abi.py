"""CPU tier: the C-ABI library loads and exports every symbol include/*.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import affaction_b200 as ab


def _declared(header):
    txt = open(os.path.join(ab.REPO_ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(aff_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built):
    L = ab.lib()
    names = _declared("affpredict.h") + _declared("affpredict_host.h")
    assert len(names) > 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(ab.ABI_SYMBOLS + ab.HOST_SYMBOLS) <= set(names)


def test_struct_layout_matches_header(built):
    assert C.sizeof(ab.Result) == 88
    assert C.sizeof(ab.Candidate) == 72
    assert C.sizeof(ab.Constraint) == 64
    assert C.sizeof(ab.TaskDesc) == 88
    o = ab.default_opts()
    assert (o.dt, o.alpha, o.lam, o.q_filt_decay, o.n_after_steps) == (0.01, 0.05, 1e-6, 0.1, 500)


def test_no_cpu_fallback(built):
    """Without a CUDA device the compute calls return AFF_ERR_CUDA; nothing is computed on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    scene = ab.HostScene()
    batch = ab.HostBatch(scene)
    batch.add_action("get fresh_tomatoes")
    with pytest.raises(ab.AffError):
        ab.predict_batch(scene, batch)
    with pytest.raises(ab.AffError):
        batch.predict()
    assert ab.lib().aff_device_count() < 0


def test_product_does_not_reference_oracle():
    """Nothing under affaction_b200/, include/ or the Makefile's product target touches oracle/."""
    bad = []
    for base in ("affaction_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ab.REPO_ROOT, base)):
            for f in files:
                if f.endswith((".py", ".h", ".cuh", ".cu", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r'#include\s+"[^"]*oracle|import\s+oracle|liboracle|libaffemu', txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def _build_c_client():
    import subprocess
    libdir = os.path.join(ab.REPO_ROOT, "affaction_b200")
    exe = os.path.join(ab.REPO_ROOT, "tests", "c_client", "abi_client")
    src = exe + ".c"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ab.REPO_ROOT, "include"), src, "-o", exe,
                    "-L" + libdir, "-laffhost", "-laffpredict", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_compiled_c_client(built):
    """include/*.h compile as C11 and link against the two libraries; a C program loads the scene, builds the
    candidates of a text command and reaches the device boundary (which fails loudly without a GPU)."""
    import subprocess
    exe = _build_c_client()
    out = subprocess.run([exe, ab.PIZZA_SCENE, "get fresh_tomatoes"], check=True, capture_output=True, text=True).stdout
    assert "bodies 455 ik_joints 23 candidates 2" in out
    assert "device present" in out or "no device: status %d" % -2 in out


def _build_cpp_client():
    import subprocess
    libdir = os.path.join(ab.REPO_ROOT, "affaction_b200")
    exe = os.path.join(ab.REPO_ROOT, "tests", "cpp_client", "shim_client")
    inc = ["-I" + os.path.join(ab.REPO_ROOT, d) for d in ("include", "affaction_b200/csrc", "affaction_b200/csrc/host")]
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror"] + inc + [exe + ".cpp", "-o", exe, "-L" + libdir, "-laffhost", "-laffpredict",
                    "-Wl,-rpath," + libdir], check=True)
    return exe


def test_shim_links_beside_classes_named_like_the_reference(built):
    """The C++ shim (namespace affgpu) in one binary with classes named aff::ActionBase / aff::ActionScene /
    aff::TrajectoryPredictor (the reference's names): INTEGRATION.md section 2 as a compiled program."""
    import subprocess
    exe = _build_cpp_client()
    out = subprocess.run([exe, ab.PIZZA_SCENE, "get fresh_tomatoes"], check=True, capture_output=True, text=True).stdout
    assert out.startswith("solutions 2 tasks ")
