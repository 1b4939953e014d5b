"""SURVEY 8(d) agreement metric on >= 1e4 rollouts of the synthetic sweep: every record of the GPU
against the CPU oracle (all host threads), per grasp variant.  Prints one summary line per variant."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import affaction_b200 as ab
import oracle_binding as ob

per = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
scene = ab.HostScene()
threads = os.cpu_count() or 1
total = bad = 0
for v, cmd in enumerate(["get tomato_sauce_bottle", "get fresh_tomatoes", "get italian_seasoning_bowl"]):
    batch = ab.HostBatch(scene)
    batch.add_perturbed(cmd, per, seed0=0x5EED0000 + v * per)
    t0 = time.perf_counter()
    gpu = ab.predict_batch(scene, batch)
    t1 = time.perf_counter()
    ref = ob.predict_batch(scene, batch, ab.default_opts(), n_threads=threads)
    t2 = time.perf_counter()
    diff = [i for i in range(per) if gpu[i].as_tuple() != ref[i].as_tuple()]
    mix = {}
    for r in gpu:
        mix[ab.CODE_NAMES[r.code]] = mix.get(ab.CODE_NAMES[r.code], 0) + 1
    same_first = ab.rank(gpu)[0] == ab.rank(ref)[0]
    print("%-28s %d rollouts: %d records differ, ranked-first equal %s, GPU %.2f s, oracle %.1f s on %d threads, verdicts %s"
          % (cmd, per, len(diff), same_first, t1 - t0, t2 - t1, threads, mix), flush=True)
    total += per
    bad += len(diff)
print("TOTAL %d rollouts, %d differing records" % (total, bad))
