"""Run the action sequences the reference ships with its pizza scenario through the CPU
oracle (test infrastructure: this tool is not part of the product) and print, per action,
the verdict of the ranked-first candidate.

The only observable the reference pins for the predict path is "the shipped sequences finish
with 0 failed actions" (/root/reference/examples/TestLLMSim.cpp:246, unittest.sh:34-53,
config/xml/pizza/g_scenario_pizza.xml:38-184).  The sequence texts are copied into
tests/golden/pizza_sequences.json so this runs without /root/reference.

usage: python tests/tools/run_sequences.py [--lift-arm-speed-limits] [name ...]      (default: all)

--lift-arm-speed-limits: the discriminating experiment of DESIGN.md section 2: the same sequences on a
copy of the scene whose seven arm joints per side have no speed limit (the fingers keep theirs).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import affaction_b200 as ab  # noqa: E402
import oracle_binding as ob  # noqa: E402


def load_sequences():
    with open(os.path.join(ROOT, "tests", "golden", "pizza_sequences.json")) as f:
        return json.load(f)


def expand(name, seqs, depth=0):
    """`s name` inside a sequence expands to that sequence (ExampleActionsECS.cpp:860-925)."""
    out = []
    for a in [x.strip() for x in seqs[name].split(";")]:
        if not a:
            continue
        w = a.split()
        if w[0] == "s" and len(w) > 1 and w[1] in seqs and depth < 4:
            out += expand(w[1], seqs, depth + 1)
        else:
            out.append(" ".join(w))
    return out


SCENE_PATH = [ab.PIZZA_SCENE]


def lifted_scene():
    out = []
    for line in open(ab.PIZZA_SCENE).read().split("\n"):
        if line.startswith("joint j2s7s300_joint_") and "finger" not in line:
            p = line.split(" ")
            p[8] = repr(1.0e308)       # joint <name> <type> <constrained> q_init q0 q_min q_max SPEED acc ...
            line = " ".join(p)
        out.append(line)
    path = "/tmp/pizza_no_arm_speed_limit.affscene"
    with open(path, "w") as f:
        f.write("\n".join(out))
    return path


def run_actions(actions, verbose=True, n_threads=8, stop_on_fail=True):
    """-> list of (action, status, detail).  status: ok / infeasible / unsupported / reset"""
    scene = ab.HostScene(SCENE_PATH[0])
    rows = []
    for a in actions:
        if a.startswith("reset"):
            scene = ab.HostScene(SCENE_PATH[0])
            rows.append((a, "reset", ""))
            continue
        b = ab.HostBatch(scene)
        try:
            n = b.add_action(a)
        except ab.AffError as e:
            rows.append((a, "unsupported", str(e)[:160]))
            if verbose:
                print("   %-52s UNSUPPORTED %s" % (a, str(e)[:100]))
            if stop_on_fail:
                break
            continue
        opts = ab.default_opts()
        res = ob.predict_batch(scene, b, opts, n_threads=n_threads)
        order = ab.rank(res)
        w = res[order[0]]
        nok = sum(1 for r in res if r.success)
        if w.success:
            r, q = ob.predict(scene, b, order[0], ab.default_opts(log_q=True), log_q=True)
            scene.apply_rollout(b, order[0], q)
            rows.append((a, "ok", "%d/%d feasible, cost %.4f" % (nok, n, w.jl_cost + w.coll_cost)))
        else:
            msg = b.format_message(w)
            detail = "0/%d feasible; best: %s step %d id (%d,%d) err %.4f mind %.4f : %s" % (
                n, ab.CODE_NAMES.get(w.first_code, w.first_code), w.first_fail_step, w.fail_id0, w.fail_id1, w.track_err,
                w.min_dist, msg[:120])
            rows.append((a, "infeasible", detail))
        if verbose:
            print("   %-52s %-10s %s" % (a, rows[-1][1], rows[-1][2]))
        if not w.success and stop_on_fail:
            break
    return rows


def main():
    seqs = load_sequences()
    args = sys.argv[1:]
    if "--lift-arm-speed-limits" in args:
        args.remove("--lift-arm-speed-limits")
        SCENE_PATH[0] = lifted_scene()
        print("arm joint speed limits lifted (%s)" % SCENE_PATH[0])
    names = args or list(seqs)
    summary = {}
    for nm in names:
        actions = expand(nm, seqs)
        print("== sequence %s (%d actions)" % (nm, len(actions)))
        rows = run_actions(actions)
        done = sum(1 for r in rows if r[1] in ("ok", "reset"))
        summary[nm] = (done, len(actions), rows[-1][1] if rows else "")
    print()
    for nm, (d, n, last) in summary.items():
        print("%-14s %3d / %3d actions  %s" % (nm, d, n, "PASS" if d == n else "stops: " + last))


if __name__ == "__main__":
    main()
