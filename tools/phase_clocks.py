"""Tuning aid: cycles per phase of the sweep kernel's step loop (a -DAFFT_PROFILE build of the library,
AFF_LIB_PATH=build_variants/prof.so).  usage: python tools/phase_clocks.py [n] [command]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affaction_b200 as ab
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
cmd = sys.argv[2] if len(sys.argv) > 2 else "get fresh_tomatoes"
os.environ["AFF_KERNEL"] = "sweep"
s = ab.HostScene(); b = ab.HostBatch(s); b.add_perturbed(cmd, n)
dev = ab.DeviceBatch(s, b, ab.default_opts())
st = torch.cuda.current_stream().cuda_stream
G = ab.lib().gpu()
out = (C.c_ulonglong * 16)()
dev.launch(st); torch.cuda.synchronize()
G.aff_debug_phase_clocks(out, 1)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); dev.launch(st); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
G.aff_debug_phase_clocks(out, 1)
names = ["A traj/constraints/dx", "B jacobian/nullspace", "C solve/integrate", "D fk", "E collision", "F costs/geo/tracking", "loop head"]
tot = sum(out[i] for i in range(7))
print("%s n=%d: %.1f ms = %.0f rollouts/s" % (cmd, n, ms, n / ms * 1e3))
for i, nm in enumerate(names):
    print("  %-24s %6.1f %%   %8.0f cycles/step" % (nm, 100.0 * out[i] / tot, out[i] / 1555.0))
sub = ["E1 shape records", "E2 body boxes", "E3 tree boxes", "E4 static near test", "E5 seg-seg pairs", "E6 seg-box/cyl pairs"]
for i, nm in enumerate(sub):
    print("    %-22s %6.1f %%   %8.0f cycles/step" % (nm, 100.0 * out[8 + i] / tot, out[8 + i] / 1555.0))
