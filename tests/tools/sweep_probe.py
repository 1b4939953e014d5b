"""GPU probe of the two kernels: parity of the thread-per-candidate ("sweep") kernel against the
warp-per-candidate kernel and the CPU oracle on a sample, then kernel time over batch sizes.
usage: python tests/tools/sweep_probe.py [sizes ...]"""
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import affaction_b200 as ab  # noqa: E402
import torch  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [32, 256, 2368, 9472, 37888]
s = ab.HostScene()
st = torch.cuda.current_stream().cuda_stream


def run(kernel, b, n_launch=2, log_q=False):
    os.environ["AFF_KERNEL"] = kernel
    dev = ab.DeviceBatch(s, b, ab.default_opts(log_q=log_q))
    best = 1e30
    for _ in range(n_launch):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dev.launch(st); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return dev, dev.results(), best


# parity: sweep == warp == oracle on 96 candidates per variant
import oracle_binding as ob  # noqa: E402
for cmd in ("get fresh_tomatoes", "get italian_seasoning_bowl", "get tomato_sauce_bottle"):
    b = ab.HostBatch(s); b.add_perturbed(cmd, 96)
    dw, rw, _ = run("warp", b, 1)
    ds, rs, _ = run("sweep", b, 1)
    ro = ob.predict_batch(s, b, ab.default_opts(), n_threads=16)
    bad_ws = sum(1 for x, y in zip(rw, rs) if x.as_tuple() != y.as_tuple())
    bad_so = sum(1 for x, y in zip(rs, ro) if x.as_tuple() != y.as_tuple())
    print("parity %-28s sweep!=warp %d  sweep!=oracle %d  of 96" % (cmd, bad_ws, bad_so), flush=True)

for cmd in ("get fresh_tomatoes", "get italian_seasoning_bowl", "get tomato_sauce_bottle"):
    for n in sizes:
        b = ab.HostBatch(s); b.add_perturbed(cmd, n)
        out = []
        for kernel in ("warp", "sweep"):
            if kernel == "warp" and n > 10000:
                out.append("warp      -    ")
                continue
            dev, res, ms = run(kernel, b)
            out.append("%s %8.1f ms %8.0f r/s" % (kernel, ms, n / ms * 1e3))
        print("%-28s n %6d  %s" % (cmd, n, "   ".join(out)), flush=True)
