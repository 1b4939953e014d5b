#!/usr/bin/env python3
"""Regenerates tests/golden/predict_get.json and predict_actions.json: the CPU oracle's verdict
records and joint-trajectory samples for the reference's pizza scene commands.

The reference pins no numeric output of this path and cannot be built here
(SURVEY 8(c)), so these are NOT outputs of the real AffAction binary; they pin
the restated predictor (oracle/oracle.c) against regressions, and the GPU must
reproduce them exactly."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import affaction_b200 as ab
import oracle_binding as ob

COMMANDS = ["get tomato_sauce_bottle", "get fresh_tomatoes", "get italian_seasoning_bowl", "get anchovies_bowl", "get bbq_sauce_bottle"]
ROWS = [0, 1, 5, 6, 100, 335, 336, 500, 665, 666, 700, 1000, 1055, 1555]


# the other actions; "after_get:<object>" = start from the state the get of <object> leaves behind
# (tests/chain_util.py), "" = the scene's model state
ACTIONS = [("", "double_get italian_seasoning_bowl garlic_bowl"), ("", "double_get italian_seasoning_bowl tomato_sauce_bottle"),
           ("", "pose TableCalibration"), ("", "pose Default"), ("", "finger_push oven"),
           ("after_get:italian_seasoning_bowl", "put italian_seasoning_bowl table"),
           ("after_get:italian_seasoning_bowl", "put italian_seasoning_bowl shelf"),
           ("after_get:italian_seasoning_bowl", "drop italian_seasoning_bowl table"),
           ("after_get:tomato_sauce_bottle", "pour tomato_sauce_bottle pizza_dough"),
           ("after_get:tomato_sauce_bottle", "poke oven tomato_sauce_bottle")]


def record(scene, cmd, opts):
    b = ab.HostBatch(scene)
    n = b.add_action(cmd)
    cands = []
    for i in range(n):
        r, q = ob.predict(scene, b, i, opts, log_q=True)
        cands.append({"result": [getattr(r, f) if not isinstance(getattr(r, f), float) else float(getattr(r, f)).hex()
                                 for f in ab.Result.COMPARE_FIELDS],
                      "q": {str(k): [float(v).hex() for v in q[k]] for k in ROWS if k < len(q)},
                      "message": b.format_message(r)})
    return cands


def start_scene(setup):
    if not setup:
        return ab.HostScene()
    from chain_util import scene_after_get
    return scene_after_get(setup.split(":", 1)[1], require_success=False)


def main():
    scene = ab.HostScene()
    opts = ab.default_opts(log_q=True)
    out = {"scene": "pizza.affscene", "dt": 0.01, "rows": ROWS, "fields": list(ab.Result.COMPARE_FIELDS), "commands": {}}
    for cmd in COMMANDS:
        out["commands"][cmd] = record(scene, cmd, opts)
    path = os.path.join(ROOT, "tests", "golden", "predict_get.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)
    out = {"scene": "pizza.affscene", "dt": 0.01, "rows": ROWS, "fields": list(ab.Result.COMPARE_FIELDS), "cases": []}
    for setup, cmd in ACTIONS:
        out["cases"].append({"setup": setup, "command": cmd, "candidates": record(start_scene(setup), cmd, opts)})
    path = os.path.join(ROOT, "tests", "golden", "predict_actions.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
