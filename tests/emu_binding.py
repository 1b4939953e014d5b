"""Binding of the TEST-ONLY warp emulator (tests/emu/libaffemu.so): the kernel
source compiled for the CPU, used to check the plan compiler and kernel logic
in the CPU test tier.  Not a product path and not a GPU result."""
import ctypes as C
import os

import numpy as np

from affaction_b200 import Result, REPO_ROOT

_lib = None


def emu_lib():
    global _lib
    if _lib is None:
        path = os.path.join(REPO_ROOT, "tests", "emu", "libaffemu.so")
        if not os.path.isfile(path):
            raise RuntimeError("%s missing: run `make emu`" % path)
        _lib = C.CDLL(path)
        vp, sz = C.c_void_p, C.c_size_t
        _lib.emu_predict_batch.restype = C.c_int
        _lib.emu_predict_batch.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, vp, vp, sz, vp, sz]
        _lib.emu_predict_batch_sweep.restype = C.c_int
        _lib.emu_predict_batch_sweep.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, vp, vp, sz, vp, sz]
        _lib.emu_predict_batch_cta.restype = C.c_int
        _lib.emu_predict_batch_cta.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, vp, vp, sz, vp, sz]
        _lib.emu_plan_info.restype = C.c_int
        _lib.emu_plan_info.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, vp, vp, sz, vp]
        _lib.emu_count_flops.restype = C.c_int
        _lib.emu_count_flops.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, vp, vp, vp, sz]
    return _lib


def predict(scene, batch, opts, first=0, count=None, log_q=False, n_steps=None, sweep=False):
    """Run candidates [first, first+count) through the emulated kernel (sweep=True: the
    thread-per-candidate kernel source, sweep.cuh; sweep="cta": the block-per-candidate arrangement,
    warp 0 + collision worker warps; else the warp-per-candidate one)."""
    n = batch.n_candidates - first if count is None else count
    res = (Result * n)()
    q = None
    rows = 0
    if log_q:
        rows = n_steps + 1
        q = np.zeros((n, rows, scene.nJ))
    err = C.create_string_buffer(512)
    fn = (emu_lib().emu_predict_batch_cta if sweep == "cta"
          else emu_lib().emu_predict_batch_sweep if sweep else emu_lib().emu_predict_batch)
    rc = fn(scene.desc, batch.tasks_ptr, batch.n_tasks, batch.constraints_ptr, batch.n_constraints,
                                     batch.candidate_ptr(first), n, C.byref(opts), res,
                                     q.ctypes.data_as(C.c_void_p) if q is not None else None, rows, err, 512)
    if rc != 0:
        raise RuntimeError("emu_predict_batch failed (%d): %s" % (rc, err.value.decode()))
    return list(res), q


def count_flops(scene, batch, opts, first=0, count=None):
    """FP64 operation counts of candidates [first, first+count): the kernel source (sweep.cuh) run on the
    CPU with a counting scalar.  -> (results, counts[n, 5]) with columns add/sub, mul, fma, div, sqrt."""
    n = batch.n_candidates - first if count is None else count
    res = (Result * n)()
    counts = np.zeros((n, 5), dtype=np.uint64)
    err = C.create_string_buffer(512)
    rc = emu_lib().emu_count_flops(scene.desc, batch.tasks_ptr, batch.n_tasks, batch.constraints_ptr, batch.n_constraints,
                                   batch.candidate_ptr(first), n, C.byref(opts), res, counts.ctypes.data_as(C.c_void_p), err, 512)
    if rc != 0:
        raise RuntimeError("emu_count_flops failed (%d): %s" % (rc, err.value.decode()))
    return list(res), counts


def plan_info(scene, batch, opts):
    """What the plan compiler made of a batch: dict(n_dyn, n_rounds, n_rounds_crit, warp_ok, sweep_ok,
    n_stat_levels, round_of=[...], parent_of=[...]) (FK schedule of the warp / block kernels)."""
    info = (C.c_int * 8)()
    rnd = (C.c_int * 256)()
    par = (C.c_int * 256)()
    rc = emu_lib().emu_plan_info(scene.desc, batch.tasks_ptr, batch.n_tasks, batch.constraints_ptr, batch.n_constraints,
                                 batch.candidate_ptr(0), batch.n_candidates, C.byref(opts), info, rnd, 256, par)
    if rc != 0:
        raise RuntimeError("emu_plan_info failed (%d)" % rc)
    n = info[0]
    return dict(n_dyn=n, n_rounds=info[1], n_rounds_crit=info[2], warp_ok=info[3], sweep_ok=info[4], n_stat_levels=info[5],
                round_of=list(rnd[:n]), parent_of=list(par[:n]))
