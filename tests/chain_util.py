"""get -> put chaining used by several tests: predict the get with the oracle, and bring the
scene into the state the rollout of candidate 0 leaves behind (object re-parented to the hand
frame at the grasp step, joints at their final values): a scene ready for `put` / `pour` / `drop`."""
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
    # the object is attached with the relative pose of the step the ConnectBodyConstraint fired in
    # and then travels with the hand to the final joint state (Graph::applyRollout)
    scene.apply_rollout(b, 0, q)
    return scene
