"""Latency of BASELINE.json configs[1] at its stated size: `put block table` on the reference's 120-Supportable
unit-test table (120 candidates in one call), after `get block` has been predicted and applied on the GPU.
Prints the end-to-end time of the host call per kernel (block per candidate: what 120 candidates select by
themselves; warp per candidate for comparison)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affaction_b200 as ab

scene = ab.HostScene(os.path.join(ab.DATA_DIR, "unittest_predict_many.affscene"))
b = ab.HostBatch(scene)
b.add_action("get block")
opts = ab.default_opts(log_q=True)
dev = ab.DeviceBatch(scene, b, opts)
dev.launch()
res = dev.results()
best = ab.rank(res)[0]
# (the ranked-first candidate is applied whatever its verdict, as tests/chain_util.py does for the parity test)
scene.apply_rollout(b, best, dev.fetch_q(best))
for kern in ("cta", "warp"):
    os.environ["AFF_KERNEL"] = kern
    for i in range(3):
        b2 = ab.HostBatch(scene)
        t0 = time.perf_counter()
        n = b2.add_action("put block table")
        t1 = time.perf_counter()
        out, msgs = b2.predict(sort=True)
        t2 = time.perf_counter()
        ok = sum(1 for r in out if r.success)
        print("%-4s put block table: %d candidates (%d feasible): candidates %.1f ms, predict + rank + messages %.1f ms" % (kern, n, ok, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
del os.environ["AFF_KERNEL"]
