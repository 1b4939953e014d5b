"""CPU tier: the oracle's building blocks against independent mathematical
properties.  The reference pins no numeric output of this path (SURVEY 8(c):
"parity unpinned"), so the restated Rcs/Tropic primitives are checked against
what they must satisfy by definition: Jacobians against finite differences,
distances against brute-force sampling, polynomials against their constraints,
the right inverse against its defining equations."""
import numpy as np
import pytest

import affaction_b200 as ab
import oracle_binding as ob


@pytest.fixture(scope="module")
def scene(built):
    return ab.HostScene()


def _task(scene, kind, eff, ref=None, frame=None, axis=2, joints=()):
    t = ab.TaskDesc()
    t.kind = kind
    t.effector = scene.body_id(eff) if eff else -1
    t.ref_body = scene.body_id(ref) if ref else -1
    t.ref_frame = scene.body_id(frame) if frame else -1
    t.axis = axis
    t.n_joints = len(joints)
    for i, j in enumerate(joints):
        t.joints[i] = j
    return t


def _q0(scene):
    b = ab.HostBatch(scene)
    b.add_action("get fresh_tomatoes")
    _, q = ob.predict(scene, b, 0, ab.default_opts(), log_q=True)
    return q


def test_fk_default_pose_is_mirror_symmetric(scene):
    """model_state "Default" is left/right symmetric: so must the world poses be."""
    A = ob.fk(scene)
    for name in ("upperarm", "forearm", "hand"):
        l, r = A[scene.body_id(name + "_left")], A[scene.body_id(name + "_right")]
        assert l[0] == pytest.approx(r[0], abs=1e-12) and l[2] == pytest.approx(r[2], abs=1e-12)
        assert l[1] == pytest.approx(-r[1], abs=1e-12)
    for b in range(scene.n_bodies):        # every rotation is orthonormal
        R = A[b, 3:].reshape(3, 3)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
    table = A[scene.body_id("table")]
    assert table[:3] == pytest.approx([-0.15, 0.0, 0.7])


TASKS = [
    (0, "hand_left_powergrasp", None, None, 2),                               # XYZ world
    (0, "tomato_sauce_bottle_grip", "hand_left_powergrasp", None, 2),         # XYZ object in hand frame
    (0, "hand_right_palmgrasp", "tomato_sauce_bottle_grip", "table", 2),      # XYZ with a separate refFrame
    (1, "hand_left", "base_footprint", None, 2), (2, "hand_right", None, None, 2), (3, "forearm_left", None, None, 2),
    (4, "hand_left_powergrasp", "tomato_sauce_bottle_grip", None, 2),         # POLAR relative
    (4, "hand_right_palmgrasp", None, None, 0),                               # POLAR world, axis X
    (5, "hand_left_pincergrasp", None, None, 0),                              # Inclination world
    (5, "hand_right_pincergrasp", "table", None, 0),
]


@pytest.mark.parametrize("kind,eff,ref,frame,axis", TASKS)
def test_task_jacobian_matches_finite_differences(scene, kind, eff, ref, frame, axis):
    q_traj = _q0(scene)
    q = q_traj[300].copy()                  # a generic mid-motion configuration
    task = _task(scene, kind, eff, ref, frame, axis)
    x0, J = ob.task_eval(scene, task, q)
    h = 1e-6
    for c in range(scene.nJ):
        qp, qm = q.copy(), q.copy()
        qp[c] += h
        qm[c] -= h
        if kind == 4:
            # POLAR: J maps dq to the angular displacement in the tangent basis; compare with
            # the task error towards the displaced polar angles, which is what computeDX returns
            xp, _ = ob.task_eval(scene, task, qp)
            xm, _ = ob.task_eval(scene, task, qm)
            num = (ob.task_dx(scene, task, q, xp) - ob.task_dx(scene, task, q, xm)) / (2 * h)
        else:
            xp, _ = ob.task_eval(scene, task, qp)
            xm, _ = ob.task_eval(scene, task, qm)
            num = (xp - xm) / (2 * h)
        assert np.allclose(J[:, c], num, atol=2e-6), (c, J[:, c], num)
    assert np.abs(J).max() > 1e-3            # the task does depend on the joints


def test_joints_task(scene):
    task = _task(scene, 6, None, joints=(scene_joint(scene, "j2s7s300_joint_finger_1_left"),))
    x, J = ob.task_eval(scene, task, None)
    assert x[0] == pytest.approx(np.radians(10.0))
    assert J.sum() == 1.0 and J.max() == 1.0


def scene_joint(scene, name):
    # joint ids are positions in the flattened joint table; find by scanning the .affscene file
    jid = -1
    k = 0
    for line in open(scene.path):
        if line.startswith("joint "):
            if line.split()[1] == name:
                jid = k
            k += 1
    assert jid >= 0
    return jid


def _frame(org, rpy=(0, 0, 0)):
    a, b, c = rpy
    sa, ca, sb, cb, sc, cc = np.sin(a), np.cos(a), np.sin(b), np.cos(b), np.sin(c), np.cos(c)
    R = np.array([[cb * cc, ca * sc + sa * sb * cc, sa * sc - ca * sb * cc],
                  [-cb * sc, ca * cc - sa * sb * sc, sa * cc + ca * sb * sc],
                  [sb, -sa * cb, ca * cb]])
    return np.concatenate([np.asarray(org, float), R.ravel()])


def _sample_shape(kind, ext, A, rng, n):
    """Points of the solid shape (surface-dense enough for a brute-force distance)."""
    R = A[3:].reshape(3, 3)
    if kind in (0, 1):     # capsule / sphere: core points, radius handled analytically
        s = rng.random(n) if kind == 0 else np.zeros(n)
        return A[:3] + np.outer(s * ext[2], R[2]), ext[0]
    if kind == 2:
        u = (rng.random((n, 3)) - 0.5)
        face = rng.integers(0, 3, n)
        u[np.arange(n), face] = np.sign(u[np.arange(n), face]) * 0.5
        return A[:3] + (u * ext) @ R, 0.0
    ang = rng.random(n) * 2 * np.pi
    rad = np.sqrt(rng.random(n)) * ext[0]
    z = (rng.random(n) - 0.5) * ext[2]
    edge = rng.random(n) < 0.5
    rad[edge] = ext[0]
    cap = (~edge) & (rng.random(n) < 0.5)
    z[cap] = np.sign(z[cap]) * 0.5 * ext[2]
    local = np.stack([rad * np.cos(ang), rad * np.sin(ang), z], axis=1)
    return A[:3] + local @ R, 0.0


@pytest.mark.parametrize("t1,t2", [(0, 0), (0, 1), (1, 1), (0, 2), (1, 2), (0, 3), (1, 3)])
def test_shape_distance_against_brute_force(t1, t2):
    rng = np.random.default_rng(100 * t1 + t2)
    exts = {0: [0.05, 0.05, 0.4], 1: [0.07, 0.07, 0.0], 2: [0.3, 0.2, 0.1], 3: [0.12, 0.12, 0.06]}
    for trial in range(12):
        A1 = _frame(rng.normal(0, 0.3, 3), rng.normal(0, 1.5, 3))
        A2 = _frame(rng.normal(0, 0.3, 3) + [0.6, 0, 0], rng.normal(0, 1.5, 3))
        d = ob.shape_distance(t1, exts[t1], A1, t2, exts[t2], A2)
        d_sym = ob.shape_distance(t2, exts[t2], A2, t1, exts[t1], A1)
        assert d == pytest.approx(d_sym, abs=1e-9)
        P1, r1 = _sample_shape(t1, exts[t1], A1, rng, 1500)
        P2, r2 = _sample_shape(t2, exts[t2], A2, rng, 1500)
        brute = np.sqrt(((P1[:, None, :] - P2[None, :, :]) ** 2).sum(-1)).min() - r1 - r2
        if d > 0:      # sampling can only over-estimate the distance of separated shapes
            assert d <= brute + 1e-9
            assert brute - d < 0.03


def test_unsupported_shape_pairs_have_no_distance():
    A = _frame([0, 0, 0])
    B = _frame([1, 0, 0])
    for t1, t2 in [(2, 2), (2, 3), (3, 3)]:
        assert ob.shape_distance(t1, [0.1, 0.1, 0.1], A, t2, [0.1, 0.1, 0.1], B) > 1e300


def test_via_point_trajectory_reaches_its_constraints():
    """(t, x, xd, xdd, flag): a position-only via point then a full one (the pregrasp /
    grasp pattern of ActionGet.cpp:573-576)."""
    dt = 0.01
    vias = np.array([[3.35, 0.2, 0, 0, 1], [6.65, 0.0, 0, 0, 7]])
    st = np.array([0.62, 0.0, 0.0])
    t = 0.0
    hit_mid = None
    xs = []
    while True:
        rel = vias.copy()
        rel[:, 0] -= t
        rel = rel[rel[:, 0] > 1e-8]
        if len(rel) == 0:
            break
        st = ob.via_step(st, rel, 1.0, dt)
        t += dt
        xs.append(st[0])
        if hit_mid is None and t >= 3.35 - 1e-9:
            hit_mid = st[0]
    assert hit_mid == pytest.approx(0.2, abs=1e-9)
    assert st[0] == pytest.approx(0.0, abs=1e-9) and abs(st[1]) < 1e-7 and abs(st[2]) < 1e-5
    xs = np.array(xs)
    assert np.abs(np.diff(xs)).max() < 0.01      # smooth: no jumps between the re-solved windows


def test_via_point_minimum_jerk_profile():
    """A single full constraint gives the quintic 10 s^3 - 15 s^4 + 6 s^5."""
    dt, T = 0.01, 2.0
    st = np.array([0.0, 0.0, 0.0])
    t = 0.0
    for k in range(100):
        st = ob.via_step(st, [[T - t, 1.0, 0, 0, 7]], 1.0, dt)
        t += dt
    s = t / T
    assert st[0] == pytest.approx(10 * s**3 - 15 * s**4 + 6 * s**5, abs=1e-10)


def test_right_inverse_equations():
    rng = np.random.default_rng(3)
    m, n = 8, 23
    J = rng.normal(size=(m, n))
    w = rng.uniform(0.2, 1.0, n)
    dx = rng.normal(size=m) * 0.01
    dH = rng.normal(size=n) * 0.01
    lam = 1e-6
    ok, dq = ob.right_inverse(J, w, dx, dH, lam)
    assert ok == 1
    W = np.diag(w)
    Jp = W @ J.T @ np.linalg.inv(J @ W @ J.T + lam * np.eye(m))
    ref = Jp @ dx - (np.eye(n) - Jp @ J) @ (w * dH)
    assert np.allclose(dq, ref, atol=1e-12)
    ok0, dq0 = ob.right_inverse(J, w, dx, np.zeros(n), lam)
    assert np.allclose(J @ dq0, dx, atol=1e-6)           # task-space part reaches dx
    okn, dqn = ob.right_inverse(J, w, np.zeros(m), dH, lam)
    assert np.abs(J @ dqn).max() < 1e-6                   # null-space part does not move the tasks


def test_libm_build_gives_the_same_verdicts(scene):
    """The deterministic sin/cos/atan2/acos (include/aff_detmath.h) versus glibc:
    verdicts identical, trajectories within the 1e-9 rad tolerance."""
    opts = ab.default_opts()
    for cmd in ("get fresh_tomatoes", "get italian_seasoning_bowl"):
        b = ab.HostBatch(scene)
        n = b.add_action(cmd)
        for i in range(n):
            r1, q1 = ob.predict(scene, b, i, opts, log_q=True)
            r2, q2 = ob.predict(scene, b, i, opts, log_q=True, libm=True)
            assert (r1.success, r1.code, r1.first_fail_step, r1.min_dist_b1, r1.min_dist_b2) == \
                   (r2.success, r2.code, r2.first_fail_step, r2.min_dist_b1, r2.min_dist_b2)
            if r1.success:
                assert np.abs(q1 - q2).max() < 1e-9
            assert r1.jl_cost == pytest.approx(r2.jl_cost, rel=1e-9)
