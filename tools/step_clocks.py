"""Tuning aid: cycles per part of a step on warp 0 of the block-per-candidate (or the warp-per-candidate)
kernel; needs a -DAFFK_PROFILE build of the library (AFF_LIB_PATH=build_variants/kprof.so).
usage: AFF_KERNEL=cta|warp python tools/step_clocks.py [n] [command]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affaction_b200 as ab
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cmd = sys.argv[2] if len(sys.argv) > 2 else "get fresh_tomatoes"
s = ab.HostScene(); b = ab.HostBatch(s); b.add_perturbed(cmd, n)
dev = ab.DeviceBatch(s, b, ab.default_opts())
st = torch.cuda.current_stream().cuda_stream
G = ab.lib().gpu()
out = (C.c_ulonglong * 16)()
dev.launch(st); torch.cuda.synchronize()
G.aff_debug_step_clocks(out, 1)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); dev.launch(st); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
G.aff_debug_step_clocks(out, 1)
names = ["0 qFilt + trajectories", "1 constraints/activation/blend", "2 dx + Jacobian columns", "3 dH null space + masks", "4 rhs + J W^-1 J^T",
         "5 Cholesky", "6 wait + solve + limits + q", "7 FK", "8 collision|post, checks, jl cost", "9 task geo + tracking", "10 fold"]
steps = dev.num_steps(0)
tot = sum(out[i] for i in range(11))
print("%s kernel %s n=%d: %.1f ms, %d steps, %.0f cycles/step" % (cmd, dev.kernel, n, ms, steps, tot / steps))
for i, nm in enumerate(names):
    print("  %-36s %6.1f %%   %8.0f cycles/step" % (nm, 100.0 * out[i] / tot, out[i] / steps))
for i, nm in ((11, "  FK: sincos"), (12, "  task geo only")):
    print("  %-36s %6.1f %%   %8.0f cycles/step" % (nm, 100.0 * out[i] / tot, out[i] / steps))
