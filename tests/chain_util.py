"""get -> put chaining used by several tests: predict the get with the oracle, take the
ranked-first candidate's final joint state, re-parent the object to the winning hand frame
(what ConnectBodyConstraint leaves behind) and return a scene ready for `put`."""
import affaction_b200 as ab
import oracle_binding as ob


def scene_after_get(obj, require_success=True):
    """require_success=False: chain on the final state of candidate 0 whatever its verdict
    (`get tomato_sauce_bottle` misses the tracking limit by 1e-4 in the restated dynamics)."""
    scene = ab.HostScene()
    b = ab.HostBatch(scene)
    b.add_action("get " + obj)
    opts = ab.default_opts(log_q=True)
    r, q = ob.predict(scene, b, 0, opts, log_q=True)
    assert r.success == 1 or not require_success
    cons = list((ab.Constraint * b.n_constraints).from_address(b.constraints_ptr))
    c0 = list((ab.Candidate * b.n_candidates).from_address(b.candidates_ptr))[0]
    conn = [c for c in cons[c0.first_constraint:c0.first_constraint + c0.n_constraints] if c.kind == 2][0]
    scene.set_ik_state(q[-1])
    scene.connect(scene.body_name(conn.body_a), scene.body_name(conn.body_b))
    return scene
