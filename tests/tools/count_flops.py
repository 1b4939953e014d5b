"""FP64 operations per rollout of the bench workload, counted by running the kernel source on the CPU
with a counting scalar (tests/emu/flop_count.cpp).  Writes profiles/flops_per_rollout.json, which
bench.py reads for the numerator of its FP64 roofline.
usage: python tests/tools/count_flops.py [candidates per variant, default 64]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import affaction_b200 as ab  # noqa: E402
import emu_binding as eb  # noqa: E402
import oracle_binding as ob  # noqa: E402
from bench import VARIANTS, PER_VARIANT  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scene = ab.HostScene()
out = {"how": "tests/emu/flop_count.cpp: sweep.cuh executed with a counting scalar; fma = 2 flops, add/sub/mul/div/sqrt = 1",
       "candidates_per_variant": n, "variants": {}}
tot = 0.0
for v, cmd in enumerate(VARIANTS):
    b = ab.HostBatch(scene)
    b.add_perturbed(cmd, n, seed0=0xAFFAC7 + v * PER_VARIANT)
    opts = ab.default_opts()
    res, cnt = eb.count_flops(scene, b, opts)
    ref = ob.predict_batch(scene, b, opts, n_threads=os.cpu_count() or 1)
    assert [r.as_tuple() for r in res] == [r.as_tuple() for r in ref], "counting run must reproduce the oracle"
    c = cnt.astype(float).mean(axis=0)
    flops = c[0] + c[1] + 2 * c[2] + c[3] + c[4]
    steps = res[0].n_steps - 1
    out["variants"][cmd] = {"add": c[0], "mul": c[1], "fma": c[2], "div": c[3], "sqrt": c[4], "flops_per_rollout": flops,
                            "steps": steps, "flops_per_step": flops / steps}
    tot += flops
    print("%-30s %.3f MFLOP per rollout (%d steps, %.1f kflop per step): add %.0f mul %.0f fma %.0f div %.0f sqrt %.0f" %
          (cmd, flops * 1e-6, steps, flops / steps * 1e-3, *c))
out["flops_per_rollout"] = tot / len(VARIANTS)
with open(os.path.join(ROOT, "profiles", "flops_per_rollout.json"), "w") as f:
    json.dump(out, f, indent=1)
print("mean %.3f MFLOP per rollout -> profiles/flops_per_rollout.json" % (out["flops_per_rollout"] * 1e-6))
