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


MANY_SCENE = "unittest_predict_many.affscene"


def scene_many_after_get():
    """The reference's 120-Supportable unit-test table (config/xml/unittest/g_scenario_unittest_predict_many.xml,
    resources/g_table_multi.xml; flattened by tools/flatten_scene.py with TwoArmJaco7 on the search path for
    its stale ../g_robo.xml include) after `get block`: the state `put block table` starts from
    (test_multi_2 of that file).  BASELINE.json configs[1] at its stated size: 120 surface candidates."""
    import os
    scene = ab.HostScene(os.path.join(ab.DATA_DIR, MANY_SCENE))
    b = ab.HostBatch(scene)
    b.add_action("get block")
    opts = ab.default_opts(log_q=True)
    res = [ob.predict(scene, b, i, opts, log_q=True) for i in range(b.n_candidates)]
    best = ab.rank([r for r, _ in res])[0]
    scene.apply_rollout(b, best, res[best][1])
    return scene
