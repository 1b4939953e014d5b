"""Latency of the small-batch call sites (TrajectoryComponent re-check; the two candidates of a `get`):
the rollout is a dependent chain of steps, so this is the time per step of the kernel, not its throughput.
Prints the kernel-only time of 1 .. 592 rollouts on the block-per-candidate and the warp-per-candidate
kernel, and the end-to-end time of the host calls (plan + H2D + kernels + D2H)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affaction_b200 as ab
import torch

s = ab.HostScene()
for kern in ("cta", "warp"):
    os.environ["AFF_KERNEL"] = kern
    for n in (1, 2, 32, 148, 296, 592):
        one = ab.HostBatch(s)
        one.add_perturbed("get fresh_tomatoes", n)
        dev = ab.DeviceBatch(s, one, ab.default_opts())
        st = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for i in range(4):
            e0.record(); dev.launch(st); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("%-4s kernel only, %3d rollout(s): %6.1f ms (%.1f us / step), grid %d" % (kern, n, min(ts), min(ts) * 1e3 / dev.num_steps(0), dev.grid_blocks))
    b = ab.HostBatch(s)
    b.add_action("get fresh_tomatoes")
    for i in range(3):
        t = time.perf_counter()
        res, msgs = b.predict(sort=True)
        t1 = time.perf_counter()
        out = b.check_trajectory(0, simulate_only=True)
        t2 = time.perf_counter()
        print("%-4s get (2 candidates) through aff_host_predict: %.1f ms; check_trajectory (q log D2H): %.1f ms, ok=%s" % (kern, (t1 - t) * 1e3, (t2 - t1) * 1e3, out["ok"]))
del os.environ["AFF_KERNEL"]
os.environ["AFF_TIMING"] = "1"
b = ab.HostBatch(s)
b.add_action("get fresh_tomatoes")
t = time.perf_counter()
b.predict(sort=True)
print("default kernel choice, with stage timing: %.1f ms" % ((time.perf_counter() - t) * 1e3))
