"""CPU tier: host shim = scene loading, candidate generation, descriptors."""
import ctypes as C

import numpy as np
import pytest

import affaction_b200 as ab


@pytest.fixture(scope="module")
def scene(built):
    return ab.HostScene()


def _cons(batch):
    n = batch.n_constraints
    return list((ab.Constraint * n).from_address(batch.constraints_ptr))


def _tasks(batch):
    n = batch.n_tasks
    return list((ab.TaskDesc * n).from_address(batch.tasks_ptr))


def _cands(batch):
    n = batch.n_candidates
    return list((ab.Candidate * n).from_address(batch.candidates_ptr))


def test_scene_dimensions(scene):
    # 23 unconstrained joints; "lab" is body 1 (ActionGet.cpp:446 relies on it)
    assert scene.nJ == 23
    assert scene.n_bodies == 455
    assert scene.body_name(1) == "lab"
    assert scene.body_id("tomato_sauce_bottle_grip") > 0
    assert scene.joint_name(0) == "ptu_pan_joint"


@pytest.mark.parametrize("cmd,ntasks,dims", [
    ("get tomato_sauce_bottle", 7, 13),      # PowerGrasp: XYZ X Y Z POLAR POLAR Joints(3)
    ("get fresh_tomatoes", 7, 12),           # BallGrasp: ... Inclination instead of POLAR
    ("get italian_seasoning_bowl", 7, 13),   # TopGrasp
])
def test_get_candidates(scene, cmd, ntasks, dims):
    b = ab.HostBatch(scene)
    assert b.add_action(cmd) == 2            # one grasp type x two free hands
    cands, tasks = _cands(b), _tasks(b)
    for c in cands:
        assert c.n_tasks == ntasks and c.pose_body == -1
        tl = tasks[c.first_task:c.first_task + c.n_tasks]
        d = sum({0: 3, 4: 2, 6: 3}.get(t.kind, 1) for t in tl)
        assert d == dims
    # candidate 0 is the hand closer to the object (sort with wLin=1, wAng=0)
    names = [b.task_name(i, 0) for i in range(2)]
    assert names[0] != names[1]


def test_get_trajectory_times(scene):
    b = ab.HostBatch(scene)
    b.add_action("get tomato_sauce_bottle")
    c0 = _cands(b)[0]
    cons = _cons(b)[c0.first_constraint:c0.first_constraint + c0.n_constraints]
    t0, t1 = 0.05, 10.05                      # delay 5*dt, duration 10 (ActionBase.cpp:74,138)
    t_grasp = t0 + 0.66 * (t1 - t0)
    t_pre = t0 + 0.5 * (t_grasp - t0)
    times = sorted(set(round(c.t, 9) for c in cons))
    for t in (t0, t_pre, t_grasp, t1, t1 + 0.5, t_grasp - 1.0, t_grasp + 1.0):
        assert round(t, 9) in times
    kinds = [c.kind for c in cons]
    assert kinds.count(2) == 1                # one ConnectBodyConstraint
    assert kinds.count(3) == 2                # collision off at pregrasp, on at the end
    conn = [c for c in cons if c.kind == 2][0]
    assert scene.body_name(conn.body_a) == "tomato_sauce_bottle"
    assert scene.body_name(conn.body_b).startswith("hand_")
    assert max(c.t for c in cons) == pytest.approx(10.55)


def test_unknown_object_raises_reference_message(scene):
    b = ab.HostBatch(scene)
    with pytest.raises(ab.AffError, match="is unknown"):
        b.add_action("get apple")            # README example; no such entity in the shipped scenes
    with pytest.raises(ab.AffError, match="not part of the GPU predict path"):
        b.add_action("putx tomato_sauce_bottle")


def test_perturbed_sweep_is_deterministic(scene):
    a, b = ab.HostBatch(scene), ab.HostBatch(scene)
    a.add_perturbed("get fresh_tomatoes", 16, seed0=7)
    b.add_perturbed("get fresh_tomatoes", 16, seed0=7)
    ca, cb = _cands(a), _cands(b)
    assert [tuple(c.pose_q6) for c in ca] == [tuple(c.pose_q6) for c in cb]
    assert len(set(tuple(c.pose_q6) for c in ca)) == 16
    assert all(c.pose_body == scene.body_id("fresh_tomatoes1") for c in ca)


def test_message_strings(scene):
    b = ab.HostBatch(scene)
    b.add_action("get fresh_tomatoes")
    r = ab.Result()
    r.idx, r.code, r.success = 0, 0, 1
    assert b.format_message(r) == "SUCCESS"
    r.code, r.success, r.fail_id0, r.fail_id1, r.fail_step = -4, 0, scene.body_id("hand_left"), scene.body_id("table"), 10
    assert b.format_message(r).startswith("ERROR: REASON: Collision problem, hand_left collides with the table SUGGESTION")
    r.code, r.fail_id0 = -3, 2
    assert "2 joint limit violations" in b.format_message(r)
    r.code = -6
    assert b.format_message(r) == "FATAL_ERROR: Initial state is invalid"
    r.code, r.fail_id0, r.fail_id1, r.fail_step, r.track_err = -5, 0, 1, 100, 0.051
    m = b.format_message(r)
    assert m.startswith("ERROR: Reachability problem REASON: Can't reach the object") and "[dim 1]" in m


def test_rank_is_lesser_with_idx_tiebreak(built):
    rs = []
    for i, (s, jl, co) in enumerate([(0, 0.1, 0.0), (1, 0.5, 0.0), (1, 0.25, 0.25), (1, 0.125, 0.0), (0, 0.0, 0.0)]):
        r = ab.Result()
        r.idx, r.success, r.jl_cost, r.coll_cost = i, s, jl, co
        rs.append(r)
    assert ab.rank(rs) == [3, 1, 2, 4, 0]


def test_put_candidates_after_get(built):
    """SURVEY 8(d) config 2 stand-in: `put <object> table` after a successful get gives the six
    Supportables with extents (g_pizza.xml:56-64), each with a BoxInterval task region."""
    from chain_util import scene_after_get
    scene = scene_after_get("italian_seasoning_bowl")
    b = ab.HostBatch(scene)
    assert b.add_action("put italian_seasoning_bowl table") == 6
    tasks, cands = _tasks(b), _cands(b)
    for c in cands:
        tl = tasks[c.first_task:c.first_task + c.n_tasks]
        assert c.n_tasks == 10                                   # ActionPut::createTasksXML
        assert [t.has_region for t in tl].count(1) == 2          # X and Y of the object on the surface
        assert tl[1].ref_frame >= 0 and tl[1].ref_frame != tl[1].ref_body   # refFrame = surface, refBdy = object
        reg = [t for t in tl if t.has_region][0]
        assert (reg.region_min, reg.region_max, reg.region_dx_scaling) == (-0.05, 0.05, 0.0)
    cons = _cons(b)[cands[0].first_constraint:cands[0].first_constraint + cands[0].n_constraints]
    t_put = 0.05 + 0.66 * 10.0
    assert any(c.kind == 2 and abs(c.t - t_put) < 1e-12 for c in cons)      # ConnectBodyConstraint(t_put, object, surface)
    b2 = ab.HostBatch(scene)
    assert b2.add_action("put italian_seasoning_bowl shelf") == 1           # the free extent-less shelf position
    with pytest.raises(ab.AffError, match="not held in the hand"):
        ab.HostBatch(scene).add_action("put tomato_sauce_bottle table")
    with pytest.raises(ab.AffError, match="already full|holding|cannot be grasped|SUCCESS"):
        ab.HostBatch(scene).add_action("get italian_seasoning_bowl")


def test_pour_after_get(built):
    """SURVEY 8(d) config 4 stand-in: `pour tomato_sauce_bottle pizza_dough` from the state the get
    leaves behind.  One candidate, 20 s, five tasks (src/ActionPour.cpp:247-279), the timing
    ladder t_prep / t_up / t_down / t_end of :303-341."""
    from chain_util import scene_after_get
    scene = scene_after_get("tomato_sauce_bottle", require_success=False)
    b = ab.HostBatch(scene)
    assert b.add_action("pour tomato_sauce_bottle pizza_dough") == 1
    tasks, cands = _tasks(b), _cands(b)
    c = cands[0]
    tl = tasks[c.first_task:c.first_task + c.n_tasks]
    assert [t.kind for t in tl] == [ab.TASK_XYZ, ab.TASK_POLAR, ab.TASK_POLAR, ab.TASK_X, ab.TASK_Z]
    tip, glas, base = scene.body_id("tomato_sauce_bottle_tip"), scene.body_id("pizza_dough_pour"), scene.body_id("base_footprint")
    assert (tl[0].effector, tl[0].ref_body, tl[0].ref_frame) == (tip, glas, base)
    assert b.task_name(0, 0) == "tomato_sauce_bottle_tip-pizza_dough_pour-base_footprint-XYZ"
    assert b.task_name(0, 4) == ""            # the reference never names taskGlasPosZ
    cons = _cons(b)[c.first_constraint:c.first_constraint + c.n_constraints]
    t0, dur = 0.05, 20.0
    t_prep, t_up, t_down, t_end = t0 + 0.15 * dur, t0 + 0.45 * dur, t0 + 0.9 * dur, t0 + dur
    via = [(k.task, k.dim, round(k.t, 9), k.flag) for k in cons if k.kind == ab.CON_VIA]
    assert (0, 1, round(t_prep, 9), 7) in via and (0, 2, round(t_prep, 9), 7) in via
    assert (0, 0, round(t_up, 9), 7) in via and (0, 2, round(t_down, 9), 1) in via and (0, 1, round(t_end, 9), 7) in via
    assert (1, 0, round(t_prep, 9), 1) in via and (1, 0, round(t_up, 9), 7) in via
    phi = {round(k.t, 9): k.x for k in cons if k.kind == ab.CON_VIA and k.task == 1 and k.dim == 0}
    assert phi[round(t_up, 9)] == pytest.approx(np.deg2rad(150.0)) and phi[round(t_end, 9)] == pytest.approx(np.deg2rad(10.0))
    acts = [(k.task, round(k.t, 9), k.flag) for k in cons if k.kind == ab.CON_ACTIVATION]
    assert sorted(a[0] for a in acts) == [0, 0, 1, 1]     # the glass stands on the table: its tasks stay off
    assert not any(k.kind in (ab.CON_CONNECT, ab.CON_COLLISION) for k in cons)
    # the window before t_up holds 3 + 1 + 1 + 3 constrained values: more than a quintic
    import oracle_binding as ob
    assert ob.oracle_lib().oracle_num_steps(b.constraints_ptr, b.candidate_ptr(0), C.byref(ab.default_opts())) == 2555
    for cmd, msg in (("pour tomato_sauce_bottle", "less than two parameters"), ("pour bbq_sauce_bottle pizza_dough", "not held in a hand"),
                     ("pour tomato_sauce_bottle table", "not a container that can be poured into"),
                     ("pour tomato_sauce_bottle pizza_dough 2.0", "overflow"), ("pour tomato_sauce_bottle pizza_dough -1", "negative value")):
        with pytest.raises(ab.AffError, match=msg):
            ab.HostBatch(scene).add_action(cmd)


def test_double_get_is_the_union_of_two_gets(scene):
    """ActionDoubleGet = two ActionGets on the best hand pairing, tasks
    unioned, trajectories merged, one candidate (src/ActionComposite.cpp:86-122,171-255)."""
    b = ab.HostBatch(scene)
    assert b.add_action("double_get italian_seasoning_bowl tomato_sauce_bottle") == 1
    names = [b.task_name(0, k) for k in range(b.n_tasks)]
    assert len(names) == 14 and len(set(names)) == 14
    # the bowl stands right of the robot, the bottle left: closest pairing
    assert names[0] == "italian_seasoning_bowl-hand_right_pincergrasp-XYZ"
    assert names[7] == "tomato_sauce_bottle-hand_left_powergrasp-XYZ"
    single = ab.HostBatch(scene)
    single.add_action("get italian_seasoning_bowl hand_right")
    single.add_action("get tomato_sauce_bottle hand_left")
    n_single = single.n_constraints
    assert b.n_constraints == n_single                      # merged, nothing dropped
    kinds = sorted(k.kind for k in _cons(b))
    assert kinds.count(ab.CON_CONNECT) == 2                 # both objects end up in a hand
    for cmd, msg in (("double_get italian_seasoning_bowl", "1 arguments, but 2 are expected"),
                     ("double_get fresh_tomatoes italian_seasoning_bowl", "Object fresh_tomatoes is unknown"),
                     ("double_get italian_seasoning_bowl nothing", "Object nothing is unknown")):
        with pytest.raises(ab.AffError, match=msg):
            ab.HostBatch(scene).add_action(cmd)


def test_drop_candidates_after_get(built):
    """ActionDrop has one solution per Supportable of the surface
    (src/ActionDrop.cpp:134-150), runs 3 s and places the object onto the surface frame with an
    identity ConnectBodyConstraint (:236-239)."""
    from chain_util import scene_after_get
    scene = scene_after_get("italian_seasoning_bowl")
    b = ab.HostBatch(scene)
    assert b.add_action("drop italian_seasoning_bowl table") == 7
    tasks, cands, cons = _tasks(b), _cands(b), _cons(b)
    surfaces = set()
    for c in cands:
        tl = tasks[c.first_task:c.first_task + c.n_tasks]
        assert [t.kind for t in tl] == [ab.TASK_INCLINATION, ab.TASK_XYZ, ab.TASK_JOINTS]
        assert tl[0].axis == 1 and tl[0].effector == tl[1].effector == scene.body_id("hand_right_pincergrasp")
        cc = cons[c.first_constraint:c.first_constraint + c.n_constraints]
        conn = [k for k in cc if k.kind == ab.CON_CONNECT]
        assert len(conn) == 1 and conn[0].dim == 1 and conn[0].t == pytest.approx(0.05 + 3.0)
        assert conn[0].body_a == scene.body_id("italian_seasoning_bowl")
        surfaces.add(scene.body_name(conn[0].body_b))
        incl = [k for k in cc if k.kind == ab.CON_VIA and k.task == 0]
        assert len(incl) == 1 and incl[0].x == pytest.approx(np.pi / 2) and incl[0].t == pytest.approx(0.05 + 1.5)
    assert len(surfaces) == 7 and "table" in surfaces
    assert b.task_name(0, 0) == "hand_right-InclinationY"
    with pytest.raises(ab.AffError, match="no hand is grasping the object"):
        ab.HostBatch(scene).add_action("drop tomato_sauce_bottle")
    with pytest.raises(ab.AffError, match="not found in scene"):
        ab.HostBatch(scene).add_action("drop nothing")


def test_pose_tasks(scene):
    """ActionPose (src/ActionPose.cpp:54-144): one Joint task per unconstrained joint of the model
    state, a single full constraint at t_end, tasks switched off half a second later."""
    b = ab.HostBatch(scene)
    assert b.add_action("pose TableCalibration") == 1
    tasks, cons = _tasks(b), _cons(b)
    assert len(tasks) == 20 and all(t.kind == ab.TASK_JOINTS and t.n_joints == 1 for t in tasks)
    assert b.task_name(0, 0) == "j2s7s300_joint_1_right" and b.task_name(0, 19) == "j2s7s300_joint_finger_3_left"
    via = [k for k in cons if k.kind == ab.CON_VIA]
    assert len(via) == 20 and all(k.flag == 7 and k.t == pytest.approx(10.05) for k in via)
    assert via[0].x == pytest.approx(np.deg2rad(31.807518)) and via[7].x == pytest.approx(np.deg2rad(10.0))
    acts = [k for k in cons if k.kind == ab.CON_ACTIVATION]
    assert sorted(set(round(k.t, 9) for k in acts)) == [0.05, 10.55]
    assert ab.HostBatch(scene).add_action("pose TableCalibration_140 0") == 1
    with pytest.raises(ab.AffError, match="Model state with name not found: Nowhere"):
        ab.HostBatch(scene).add_action("pose Nowhere")
    with pytest.raises(ab.AffError, match="Model state with name not found"):
        ab.HostBatch(scene).add_action("pose TableCalibration 3")


def test_finger_push_tasks(scene):
    """ActionFingerPush (src/ActionFingerPush.cpp:67-306): closest free finger onto the object's
    PointPushable frame, 16 s, five-point timing ladder."""
    b = ab.HostBatch(scene)
    assert b.add_action("finger_push oven") == 1
    tasks, cons = _tasks(b), _cons(b)
    assert [t.kind for t in tasks] == [ab.TASK_XYZ, ab.TASK_POLAR, ab.TASK_JOINTS]
    assert tasks[0].ref_body == tasks[1].ref_body == scene.body_id("oven_switch")
    assert b.task_name(0, 0).startswith("PushButton-XYZ-hand_") and b.task_name(0, 2).endswith("_fingers")
    t0, dur = 0.05, 16.0
    z = sorted((round(k.t, 9), k.x) for k in cons if k.kind == ab.CON_VIA and k.task == 0 and k.dim == 2)
    assert z == [(round(t0 + 0.5 * dur, 9), 0.05), (round(t0 + 0.7 * dur, 9), 0.0), (round(t0 + 0.8 * dur, 9), -0.03), (round(t0 + dur, 9), 0.1)]
    fingers = [k.x for k in cons if k.kind == ab.CON_VIA and k.task == 2]
    assert fingers == [1.32] * 6
    assert ab.HostBatch(scene).add_action("poke oven") == 1
    for cmd, msg in (("finger_push table", "is not switchable"), ("finger_push nothing", "is unknown"),
                     ("finger_push oven table", "cannot be used to switch something on")):
        with pytest.raises(ab.AffError, match=msg):
            ab.HostBatch(scene).add_action(cmd)


def test_apply_rollout_attaches_at_the_grasp_step(built):
    """Graph::applyRollout: the object is re-parented with the relative pose of the step the
    ConnectBodyConstraint fired in and then follows the hand: after `get` it sits in the hand,
    liftHeight (0.12 m) above where it stood."""
    import oracle_binding as ob
    from chain_util import scene_after_get
    before = ob.fk(ab.HostScene())
    scene = scene_after_get("italian_seasoning_bowl")
    after = ob.fk(scene)
    bowl, hand = scene.body_id("italian_seasoning_bowl"), scene.body_id("hand_right_pincergrasp")
    assert after[bowl, 2] - before[bowl, 2] == pytest.approx(0.12, abs=2e-3)
    assert np.linalg.norm(after[bowl, :3] - after[hand, :3]) < 0.15 < np.linalg.norm(before[bowl, :3] - before[hand, :3])
    # held: a second get of the same object is refused the way the reference words it
    with pytest.raises(ab.AffError, match="already held|SUCCESS"):
        ab.HostBatch(scene).add_action("get italian_seasoning_bowl")


def test_reparenting_onto_a_later_body_keeps_the_scene_valid(built):
    """Graph::restoreParentFirstOrder: after a connect onto a body with a HIGHER index the body array is
    re-sorted (parents first), ids are remapped by name, the world pose is kept, and prediction works."""
    import numpy as np
    import oracle_binding as ob
    import emu_binding as eb
    s = ab.HostScene()
    before = s.body_id("italian_seasoning_bowl")
    assert before < s.body_id("tray_pos2")
    pose = ob.fk(s)[before].copy()
    s.connect("italian_seasoning_bowl", "tray_pos2")
    after = s.body_id("italian_seasoning_bowl")
    assert after > s.body_id("tray_pos2") and after != before
    assert np.abs(ob.fk(s)[after] - pose).max() < 1e-12
    b = ab.HostBatch(s)
    assert b.add_action("get italian_seasoning_bowl") == 2
    opts = ab.default_opts()
    ref, _ = ob.predict(s, b, 0, opts)
    for sweep in (False, True):
        (emu,), _ = eb.predict(s, b, opts, first=0, count=1, sweep=sweep)
        assert emu.as_tuple() == ref.as_tuple()


def test_rollout_transforms_replay(built):
    """aff_host_rollout_transforms: PredictionResult::bodyTransforms of one candidate from its q(t)
    (reference src/TrajectoryPredictor.cpp:111-118,356-363).  Row 0 is the scene's FK, the last row is the
    FK of the state apply_rollout leaves, static bodies never move, and after the grasp the object
    travels with the hand frame it was connected to."""
    import numpy as np
    import oracle_binding as ob
    s = ab.HostScene()
    b = ab.HostBatch(s)
    b.add_action("get fresh_tomatoes")
    r, q = ob.predict(s, b, 0, ab.default_opts(log_q=True), log_q=True)
    assert r.success == 1
    T = s.rollout_transforms(b, 0, q)
    assert T.shape == (q.shape[0], s.n_bodies, 12)
    assert np.abs(T[0] - ob.fk(s)).max() < 1e-12       # host FK (libm) vs the oracle's: same transforms, not the same bits
    table, obj = s.body_id("table"), s.body_id("fresh_tomatoes1")       # the body the get connects to the hand
    assert np.abs(T[:, table] - T[0, table]).max() == 0.0
    rot = T[:, :, 3:].reshape(T.shape[0], s.n_bodies, 3, 3)
    assert np.abs(np.einsum("rbij,rbkj->rbik", rot[::97], rot[::97]) - np.eye(3)).max() < 1e-9      # orthonormal
    # lifted by about liftHeight at the end, untouched before the grasp (t_grasp = 0.05 + 0.66 * 10 s)
    assert np.abs(T[:600, obj] - T[0, obj]).max() == 0.0
    assert 0.08 < T[-1, obj, 2] - T[0, obj, 2] < 0.2
    s2 = ab.HostScene()
    b2 = ab.HostBatch(s2)
    b2.add_action("get fresh_tomatoes")
    s2.apply_rollout(b2, 0, q)
    A_end = ob.fk(s2)
    for name in ("fresh_tomatoes1", "hand_left", "hand_right", "table"):
        assert np.abs(A_end[s2.body_id(name)] - T[-1, s.body_id(name)]).max() < 1e-12


def test_joint_limit_message_lists_the_joints(built):
    """reference src/TrajectoryPredictor.cpp:807-818: with the joint values of the failing step the
    joint-limit message names every violated joint with its value and range."""
    import oracle_binding as ob
    s = ab.HostScene()
    b = ab.HostBatch(s)
    b.add_perturbed("get italian_seasoning_bowl", 64, seed0=0xAFFAC7)
    opts = ab.default_opts(log_q=True)
    for i in range(64):
        r, q = ob.predict(s, b, i, opts, log_q=True)
        if r.code == -3:
            short = b.format_message(r)
            long = b.format_message(r, q_fail=q[r.fail_step + 1])
            assert long.startswith(short) and long.count(" outside [") == r.fail_id0 >= 1
            return
    pytest.skip("no joint-limit verdict among the first 64 perturbed candidates")


def test_graph_from_desc_roundtrip(built):
    """Graph::fromDesc (the reference-tree entry: description from the live RcsGraph via
    integration/rcs_graph_adapter.h) inverts Graph::toDesc on both committed scenes."""
    import os
    for f in ("pizza.affscene", "unittest_predict_many.affscene"):
        s = ab.HostScene(os.path.join(ab.DATA_DIR, f))
        assert ab.lib().aff_host_scene_desc_roundtrip(s.h) == 1


def test_perturbed_candidates_do_not_depend_on_the_host_threads(monkeypatch):
    """aff_host_batch_add_perturbed describes slices of the index range on host threads and appends them in
    index order: the task / constraint / candidate tables are byte-identical to the single-threaded ones."""
    import ctypes as C
    scene = ab.HostScene()

    def tables(threads):
        monkeypatch.setenv("AFF_HOST_THREADS", str(threads))
        b = ab.HostBatch(scene)
        b.add_perturbed("get tomato_sauce_bottle", 7)          # something in front, so that the offsets matter
        assert b.add_perturbed("get fresh_tomatoes", 5000, seed0=77) == 5000
        return (b.n_candidates,
                C.string_at(b.tasks_ptr, b.n_tasks * C.sizeof(ab.TaskDesc)),
                C.string_at(b.constraints_ptr, b.n_constraints * C.sizeof(ab.Constraint)),
                C.string_at(b.candidate_ptr(0), b.n_candidates * C.sizeof(ab.Candidate)),
                [b.command(i) for i in (0, 6, 7, 2600, 5006)] if hasattr(b, "command") else None)

    assert tables(1) == tables(4)


def test_perturbed_candidates_from_prototype_tasks_equal_the_generic_path(monkeypatch):
    """add_perturbed copies the task records of a command and hand from one prototype per solution (they do not
    depend on the perturbation) and converts only the trajectory: byte-identical to the generic append()."""
    import ctypes as C
    scene = ab.HostScene()

    def tables(generic):
        if generic:
            monkeypatch.setenv("AFF_HOST_PERTURBED_GENERIC", "1")
        else:
            monkeypatch.delenv("AFF_HOST_PERTURBED_GENERIC", raising=False)
        b = ab.HostBatch(scene)
        for cmd, n in (("get tomato_sauce_bottle", 301), ("get italian_seasoning_bowl", 3000)):
            assert b.add_perturbed(cmd, n, seed0=11) == n
        return (C.string_at(b.tasks_ptr, b.n_tasks * C.sizeof(ab.TaskDesc)),
                C.string_at(b.constraints_ptr, b.n_constraints * C.sizeof(ab.Constraint)),
                C.string_at(b.candidate_ptr(0), b.n_candidates * C.sizeof(ab.Candidate)))

    assert tables(True) == tables(False)
