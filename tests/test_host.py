"""CPU tier: host shim = scene loading, candidate generation, descriptors."""
import ctypes as C

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
