"""Breakdown of the host-buffer path (AFF_TIMING=1 prints the stages)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["AFF_TIMING"] = "1"
import affaction_b200 as ab
s = ab.HostScene(); b = ab.HostBatch(s); b.add_perturbed("get fresh_tomatoes", 9472)
for i in range(3):
    t = time.perf_counter(); ab.predict_batch(s, b); print("predict_batch %.1f ms" % (1e3 * (time.perf_counter() - t)))
