"""throughput of the host-buffer call (aff_predict_batch) over batch sizes
1e3 ... 1e6 perturbed-goal candidates on one GPU.  Prints one line per size."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affaction_b200 as ab

sizes = [int(float(x)) for x in (sys.argv[1:] or ["1e3", "1e4", "1e5", "1e6"])]
scene = ab.HostScene()
warm = ab.HostBatch(scene)
warm.add_perturbed("get fresh_tomatoes", 64)
ab.predict_batch(scene, warm)
for n in sizes:
    t0 = time.perf_counter()
    batch = ab.HostBatch(scene)
    batch.add_perturbed("get fresh_tomatoes", n, seed0=0xAFFAC7)
    t1 = time.perf_counter()
    res = ab.predict_batch(scene, batch)
    t2 = time.perf_counter()
    ok = sum(r.success for r in res)
    print("N %8d  candidate generation %.2f s  predict %.3f s = %.0f rollouts/s  host tables %.1f MB  feasible %d (%.1f %%)"
          % (n, t1 - t0, t2 - t1, n / (t2 - t1), batch.host_bytes() / 1e6, ok, 100.0 * ok / n), flush=True)
    del batch, res
