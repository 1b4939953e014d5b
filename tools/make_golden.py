#!/usr/bin/env python3
"""Regenerates tests/golden/predict_get.json: the CPU oracle's verdict records and
joint-trajectory samples for the reference's pizza scene commands.

The reference pins no numeric output of this path and cannot be built here
(SURVEY 8(c)), so these are NOT outputs of the real AffAction binary; they pin
the restated predictor (oracle/oracle.c) against regressions, and the GPU must
reproduce them exactly."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import affaction_b200 as ab
import oracle_binding as ob

COMMANDS = ["get tomato_sauce_bottle", "get fresh_tomatoes", "get italian_seasoning_bowl", "get anchovies_bowl", "get bbq_sauce_bottle"]
ROWS = [0, 1, 5, 6, 100, 335, 336, 500, 665, 666, 700, 1000, 1055, 1555]


def main():
    scene = ab.HostScene()
    opts = ab.default_opts(log_q=True)
    out = {"scene": "pizza.affscene", "dt": 0.01, "rows": ROWS, "fields": list(ab.Result.COMPARE_FIELDS), "commands": {}}
    for cmd in COMMANDS:
        b = ab.HostBatch(scene)
        n = b.add_action(cmd)
        cands = []
        for i in range(n):
            r, q = ob.predict(scene, b, i, opts, log_q=True)
            cands.append({"result": [getattr(r, f) if not isinstance(getattr(r, f), float) else float(getattr(r, f)).hex()
                                     for f in ab.Result.COMPARE_FIELDS],
                          "q": {str(k): [float(v).hex() for v in q[k]] for k in ROWS if k < len(q)},
                          "message": b.format_message(r)})
        out["commands"][cmd] = cands
    path = os.path.join(ROOT, "tests", "golden", "predict_get.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
