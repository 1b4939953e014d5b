"""Latency of the batch-of-one call sites (TrajectoryComponent re-check; single-candidate actions):
one rollout = one warp, so this is the step-chain latency of the kernel, not its throughput."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affaction_b200 as ab
import torch

s = ab.HostScene()
b = ab.HostBatch(s)
b.add_action("get fresh_tomatoes")
for n in (1, 2):
    one = ab.HostBatch(s)
    one.add_perturbed("get fresh_tomatoes", n)
    dev = ab.DeviceBatch(s, one, ab.default_opts())
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(4):
        e0.record(); dev.launch(st); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("kernel only, %d rollout(s): %.1f ms (%.1f us / step)" % (n, min(ts), min(ts) * 1e3 / dev.num_steps(0)))
for i in range(3):
    t = time.perf_counter()
    out = b.check_trajectory(0, simulate_only=True)
    print("check_trajectory e2e (plan + H2D + kernel + q log D2H): %.1f ms, ok=%s" % ((time.perf_counter() - t) * 1e3, out["ok"]))
