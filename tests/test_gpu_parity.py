"""GPU parity tests proper: the CUDA path, called through the C ABI, against the
CPU oracle on the same inputs.  Bar: every aff_result field equal (verdict,
failure code/step/ids, closest pair, jMask, costs, min distance — compared as
doubles with ==) and joint trajectories within 1e-9 rad (BASELINE.json)."""
import os

import numpy as np
import pytest

import affaction_b200 as ab
import oracle_binding as ob

pytestmark = pytest.mark.gpu

Q_TOL = 1e-9   # rad, north_star tolerance for joint trajectories

GET_COMMANDS = ["get tomato_sauce_bottle", "get fresh_tomatoes", "get italian_seasoning_bowl", "get anchovies_bowl",
                "get bbq_sauce_bottle"]


@pytest.fixture(scope="module")
def scene(built):
    return ab.HostScene()


@pytest.fixture(autouse=True, params=["warp", "sweep", "cta"])
def kernel(request, monkeypatch):
    """Every parity test runs on the three kernels: one rollout per warp (mid-size batches), one rollout per
    thread (sweeps; AFF_SWEEP_MIN=1 selects it for every batch within its per-thread caps, anything larger
    still runs on the warp kernel) and one rollout per block (a handful of candidates)."""
    monkeypatch.delenv("AFF_CTA_MAX", raising=False)
    if request.param == "warp":
        monkeypatch.setenv("AFF_KERNEL", "warp")
    elif request.param == "cta":
        monkeypatch.setenv("AFF_KERNEL", "cta")
    else:
        monkeypatch.delenv("AFF_KERNEL", raising=False)
        monkeypatch.setenv("AFF_SWEEP_MIN", "1")
        monkeypatch.setenv("AFF_CTA_MAX", "0")
    return request.param


def _compare(scene, batch, opts, dev, idxs, check_q=True):
    gpu = dev.results()
    worst = 0.0
    for i in idxs:
        ref, q_ref = ob.predict(scene, batch, i, opts, log_q=check_q)
        assert gpu[i].as_tuple() == ref.as_tuple(), "candidate %d: GPU %s != oracle %s" % (i, gpu[i].as_tuple(), ref.as_tuple())
        assert gpu[i].idx == i
        if check_q:
            q_gpu = dev.fetch_q(i)
            assert q_gpu.shape == q_ref.shape
            worst = max(worst, float(np.abs(q_gpu - q_ref).max()))
    assert worst <= Q_TOL
    return worst


def test_smoke_entry(built):
    import __graft_entry__ as ge
    ge.smoke()


@pytest.mark.parametrize("cmd", GET_COMMANDS)
def test_get_candidates_match_oracle(scene, cmd):
    batch = ab.HostBatch(scene)
    n = batch.add_action(cmd)
    assert n == 2
    opts = ab.default_opts(log_q=True)
    dev = ab.DeviceBatch(scene, batch, opts)
    assert dev.launch() == 2          # prologue + rollout kernel
    worst = _compare(scene, batch, opts, dev, range(n))
    assert worst == 0.0               # in fact bit-identical


@pytest.mark.parametrize("cmd", GET_COMMANDS)
def test_ranked_first_matches_oracle(scene, cmd):
    """aff_predict_batch (host buffers in / out) + PredictionResult::lesser ranking."""
    batch = ab.HostBatch(scene)
    batch.add_action(cmd)
    opts = ab.default_opts()
    gpu = ab.predict_batch(scene, batch, opts)
    ref = ob.predict_batch(scene, batch, opts, n_threads=os.cpu_count() or 1)
    assert [r.as_tuple() for r in gpu] == [r.as_tuple() for r in ref]
    assert ab.rank(gpu) == ab.rank(ref)


def test_too_many_objects_in_one_device_batch_is_reported(scene):
    """Capacity errors are loud, not silent fallbacks: one device batch tracks AFFK_MAX_OBJ objects."""
    batch = ab.HostBatch(scene)
    for cmd in GET_COMMANDS[:3]:
        batch.add_action(cmd)
    with pytest.raises(ab.AffError, match="AFFK_MAX_OBJ"):
        ab.DeviceBatch(scene, batch, ab.default_opts())


def test_mixed_commands_share_one_predict_call(scene):
    """aff_predict_batch splits a call whose candidates name more objects than one device batch tracks
    (three different gets = three objects) into groups and returns the records in the caller's order."""
    batch = ab.HostBatch(scene)
    for cmd in GET_COMMANDS[:3]:
        batch.add_action(cmd)
    opts = ab.default_opts()
    gpu = ab.predict_batch(scene, batch, opts)
    ref = ob.predict_batch(scene, batch, opts, n_threads=os.cpu_count() or 1)
    assert [r.idx for r in gpu] == list(range(6))
    assert [r.as_tuple() for r in gpu] == [r.as_tuple() for r in ref]


def test_double_get_pairs_extension(scene):
    """BASELINE.json configs[2]: the bimanual candidate set as the Cartesian product of the two gets'
    solutions (hands distinct) - `double_get_pairs`, an extension; the reference's double_get pins one pairing."""
    batch = ab.HostBatch(scene)
    assert batch.add_action("double_get_pairs italian_seasoning_bowl tomato_sauce_bottle") == 2
    opts = ab.default_opts(log_q=True)
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    _compare(scene, batch, opts, dev, range(2))
    hands = [(batch.task_name(i, 0), batch.task_name(i, 7)) for i in range(2)]
    assert "hand_right" in hands[0][0] and "hand_left" in hands[0][1] and "hand_left" in hands[1][0] and "hand_right" in hands[1][1]


def test_perturbed_sweep_matches_oracle(scene):
    """512 synthetic perturbed-goal candidates, all compared."""
    batch = ab.HostBatch(scene)
    n = batch.add_perturbed("get fresh_tomatoes", 512)
    opts = ab.default_opts()
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    gpu = dev.results()
    ref = ob.predict_batch(scene, batch, opts, n_threads=os.cpu_count() or 1)
    bad = [i for i in range(n) if gpu[i].as_tuple() != ref[i].as_tuple()]
    assert not bad, "verdict / cost mismatch for candidates %s" % bad[:10]
    codes = sorted(set(r.code for r in gpu))
    assert len(codes) >= 2, "the sweep should produce a mix of verdicts, got %s" % codes


def test_perturbed_trajectories(scene):
    batch = ab.HostBatch(scene)
    batch.add_perturbed("get italian_seasoning_bowl", 8, seed0=1234)
    opts = ab.default_opts(log_q=True)
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    _compare(scene, batch, opts, dev, range(8))


def test_relaunch_is_idempotent(scene):
    batch = ab.HostBatch(scene)
    batch.add_perturbed("get fresh_tomatoes", 64, seed0=99)
    dev = ab.DeviceBatch(scene, batch, ab.default_opts())
    dev.launch()
    a = [r.as_tuple() for r in dev.results()]
    assert dev.launch() == 1          # prologue runs once per batch
    b = [r.as_tuple() for r in dev.results()]
    assert a == b


def test_host_predict_messages(scene):
    batch = ab.HostBatch(scene)
    batch.add_action("get fresh_tomatoes")
    res, msgs = batch.predict(sort=True)
    assert res[0].success == 1 and msgs[0].startswith("SUCCESS")
    assert res[1].success == 0 and msgs[1].startswith("ERROR")
    assert "command: get fresh_tomatoes" in msgs[1]


def test_put_after_get_matches_oracle(built):
    """BASELINE.json configs[1] stand-in: `put` with its location candidates,
    chained on the final state of the ranked-first `get` candidate."""
    from chain_util import scene_after_get
    scene = scene_after_get("italian_seasoning_bowl")
    for cmd, n_expected in (("put italian_seasoning_bowl table", 6), ("put italian_seasoning_bowl shelf", 1)):
        batch = ab.HostBatch(scene)
        assert batch.add_action(cmd) == n_expected
        opts = ab.default_opts(log_q=True)
        dev = ab.DeviceBatch(scene, batch, opts)
        dev.launch()
        worst = _compare(scene, batch, opts, dev, range(n_expected))
        assert worst == 0.0
    res, msgs = batch.predict(sort=True)
    assert len(res) == 1


def test_pour_after_get_matches_oracle(built):
    """BASELINE.json configs[3] stand-in: the shipped sequence
    `get tomato_sauce_bottle; pour tomato_sauce_bottle pizza_dough; put tomato_sauce_bottle shelf`
    (g_scenario_pizza.xml:39) chained action by action.  The pour is a 20 s rollout (2555 steps) with
    trajectory windows of eight constrained values; the bottle rides in the hand as a plain rigid body."""
    from chain_util import scene_after_get
    scene = scene_after_get("tomato_sauce_bottle", require_success=False)
    for cmd in ("pour tomato_sauce_bottle pizza_dough", "put tomato_sauce_bottle shelf"):
        batch = ab.HostBatch(scene)
        n = batch.add_action(cmd)
        assert n >= 1
        opts = ab.default_opts(log_q=True)
        dev = ab.DeviceBatch(scene, batch, opts)
        dev.launch()
        if cmd.startswith("pour"):
            assert dev.num_steps(0) == 2555
        assert _compare(scene, batch, opts, dev, range(n)) == 0.0
        if cmd.startswith("pour"):
            # next action starts where this one ended
            scene.set_ik_state(dev.fetch_q(0)[-1])


@pytest.mark.parametrize("objects", ["italian_seasoning_bowl garlic_bowl", "italian_seasoning_bowl tomato_sauce_bottle",
                                     "mushrooms_bowl italian_seasoning_bowl"])
def test_double_get_matches_oracle(scene, objects):
    """BASELINE.json configs[2] stand-in: both arms move in one rollout,
    14 tasks / up to 26 task dimensions, two objects re-parented."""
    batch = ab.HostBatch(scene)
    assert batch.add_action("double_get " + objects) == 1
    opts = ab.default_opts(log_q=True)
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    assert _compare(scene, batch, opts, dev, [0]) == 0.0


def test_trajectory_component_recheck(scene):
    """the pre-execution re-check (TrajectoryComponent::checkerThread,
    src/TrajectoryComponent.cpp:346-399) = the same rollout at batch 1, with the events the component
    publishes and the prediction array for the ghost animation."""
    batch = ab.HostBatch(scene)
    assert batch.add_action("get fresh_tomatoes") == 2
    opts = ab.default_opts(log_q=True)
    verdicts = []
    for cand in range(2):
        ref, q_ref = ob.predict(scene, batch, cand, opts, log_q=True)
        verdicts.append(ref.success)
        out = batch.check_trajectory(cand)                    # onCheckAndSetTrajectory
        assert out["raw"].as_tuple() == ref.as_tuple()
        assert np.array_equal(out["q"], q_ref)
        if ref.success:
            assert out["set_trajectory"] and not out["action_result"]
        else:
            assert out["action_result"] and not out["ok"] and not out["set_trajectory"]
            assert out["message"] == batch.format_message(ref)
        sim = batch.check_trajectory(cand, simulate_only=True)   # onSimulateTrajectory
        assert not sim["set_trajectory"] and sim["action_result"] and sim["ok"] == bool(ref.success)
        if ref.success:
            assert sim["message"] == "SUCCESS DEVELOPER: Simulated trajectory is valid"
            assert sim["quality"] == 1.0 / (1.0 + ref.jl_cost)
    assert sorted(verdicts) == [0, 1]
    stop = batch.check_trajectory(0, estop=True)
    assert stop["action_result"] and not stop["ok"] and stop["message"] == "FATAL_ERROR REASON: Emergency-Stop triggered"
    assert len(stop["q"]) == 0
    off = batch.check_trajectory(1, check_enabled=False)
    assert off["set_trajectory"] and not off["action_result"] and len(off["q"]) == 0


def test_drop_after_get_matches_oracle(built):
    """`drop`: seven surface candidates, 805 steps each, the object re-parented
    with an identity connect transform at the end."""
    from chain_util import scene_after_get
    scene = scene_after_get("italian_seasoning_bowl")
    batch = ab.HostBatch(scene)
    n = batch.add_action("drop italian_seasoning_bowl table")
    assert n == 7
    n += batch.add_action("drop italian_seasoning_bowl")          # surface found by the ray cast
    opts = ab.default_opts(log_q=True)
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    assert dev.num_steps(0) == 805
    assert _compare(scene, batch, opts, dev, range(n)) == 0.0


@pytest.mark.parametrize("state", ["TableCalibration", "TableCalibration_140", "Default"])
def test_pose_matches_oracle(scene, state):
    """`pose`: 20 one-joint tasks, a 20 x 20 right inverse every step."""
    batch = ab.HostBatch(scene)
    assert batch.add_action("pose " + state) == 1
    opts = ab.default_opts(log_q=True)
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    assert _compare(scene, batch, opts, dev, [0]) == 0.0


def test_finger_push_matches_oracle(built):
    """`finger_push` with a finger and, after a get, with the held bottle's tip as the tool."""
    from chain_util import scene_after_get
    for scene, cmd in ((ab.HostScene(), "finger_push oven"), (scene_after_get("tomato_sauce_bottle", require_success=False), "poke oven tomato_sauce_bottle")):
        batch = ab.HostBatch(scene)
        assert batch.add_action(cmd) == 1
        opts = ab.default_opts(log_q=True)
        dev = ab.DeviceBatch(scene, batch, opts)
        dev.launch()
        assert dev.num_steps(0) == 2155
        assert _compare(scene, batch, opts, dev, [0]) == 0.0


def test_predict_batch_chunks_equal_single_launch(scene):
    """aff_predict_batch pipelines large batches in chunks of whole waves (plan compile of chunk k+1
    under the kernels of chunk k): same records, same order, idx = position in the caller's array."""
    n = 3 * 2368 + 100
    batch = ab.HostBatch(scene)
    batch.add_perturbed("get fresh_tomatoes", n, seed0=99)
    dev = ab.DeviceBatch(scene, batch, ab.default_opts())
    dev.launch()
    single = dev.results()
    chunked = ab.predict_batch(scene, batch)
    assert [r.idx for r in chunked] == list(range(n))
    assert [r.as_tuple() for r in chunked] == [r.as_tuple() for r in single]


def test_full_size_sweep_properties(scene):
    """The bench workload at its full size (BASELINE.json configs[4]: 3 x 9472 perturbed candidates,
    the seeds bench.py uses on rank 0), checked through size-independent properties: the verdict mix is
    the pinned one, a second run is identical, the ranked order is sorted by (success, cost) with the
    best candidate feasible, and a random sample of candidates predicted alone (batch of one, and by
    the CPU oracle) gives the same records as inside the big batch."""
    import random
    per = 9472
    pinned = {"get tomato_sauce_bottle": None, "get fresh_tomatoes": None, "get italian_seasoning_bowl": None}
    mix = {}
    rng = random.Random(7)
    for v, cmd in enumerate(pinned):
        batch = ab.HostBatch(scene)
        batch.add_perturbed(cmd, per, seed0=0xAFFAC7 + v * per)
        res = ab.predict_batch(scene, batch)
        assert [r.idx for r in res] == list(range(per))
        again = ab.predict_batch(scene, batch)
        assert [r.as_tuple() for r in again] == [r.as_tuple() for r in res]
        for r in res:
            name = ab.CODE_NAMES[r.code]
            mix[name] = mix.get(name, 0) + 1
            assert r.n_steps == 1556 and (r.success == 1) == (r.code == 0)
        order = ab.rank(res)
        keys = [(0 if res[i].success else 1, res[i].jl_cost + res[i].coll_cost) for i in order]
        assert keys == sorted(keys) and sorted(order) == list(range(per))
        assert res[order[0]].success == 1
        opts = ab.default_opts()
        for i in rng.sample(range(per), 4):
            ref, _ = ob.predict(scene, batch, i, opts)
            assert res[i].as_tuple() == ref.as_tuple()
    assert mix == {"tracking": 13934, "ok": 9720, "collision": 4110, "joint_limit": 652}


def test_every_get_in_the_scene_matches_oracle(scene):
    """All 56 graspable entities of the pizza scene (bowls, bottles, plates, ingredients, tray ...),
    two candidates each: every record equal to the oracle's."""
    names = []
    for line in open(os.path.join(os.path.dirname(ab.__file__), "data", "pizza.affscene")):
        if line.startswith("entity ") and line.split()[1] not in names:
            names.append(line.split()[1])
    opts = ab.default_opts()
    checked, verdicts = 0, set()
    for name in names:
        batch = ab.HostBatch(scene)
        try:
            n = batch.add_action("get " + name)
        except ab.AffError:
            continue                      # furniture: "cannot be grasped"
        gpu = ab.predict_batch(scene, batch, opts)
        ref = ob.predict_batch(scene, batch, opts, n_threads=min(n, os.cpu_count() or 1))
        for i in range(n):
            assert gpu[i].as_tuple() == ref[i].as_tuple(), "get %s candidate %d" % (name, i)
            verdicts.add(gpu[i].code)
        checked += n
    assert checked == 112 and len(verdicts) >= 3


def test_ragged_batch_and_edge_sizes(scene):
    """One batch mixing actions of different shapes (2 get candidates with 7 tasks / 1555 steps, a pose
    with 20 one-joint tasks, a 16 s finger_push with 2155 steps): ragged task counts, task dimensions
    and step counts in one launch.  Then the edge sizes: an empty batch and a batch of one."""
    batch = ab.HostBatch(scene)
    n = batch.add_action("get fresh_tomatoes") + batch.add_action("pose TableCalibration") + batch.add_action("finger_push oven")
    assert n == 4
    opts = ab.default_opts(log_q=True)
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    assert [dev.num_steps(i) for i in range(n)] == [1555, 1555, 1555, 2155]
    assert _compare(scene, batch, opts, dev, range(n)) == 0.0
    assert ab.predict_batch(scene, ab.HostBatch(scene)) == []          # nothing to predict is not an error
    one = ab.HostBatch(scene)
    one.add_perturbed("get fresh_tomatoes", 1)
    ref, _ = ob.predict(scene, one, 0, ab.default_opts())
    assert ab.predict_batch(scene, one)[0].as_tuple() == ref.as_tuple()


def test_batches_with_different_footprints_coexist(scene):
    """A batch created earlier with the larger shared-memory footprint must still launch after a
    smaller one has been created (the dynamic shared-memory limit belongs to the kernel)."""
    big = ab.HostBatch(scene)
    big.add_perturbed("get italian_seasoning_bowl", 8)
    small = ab.HostBatch(scene)
    small.add_perturbed("get fresh_tomatoes", 8)
    dev_big = ab.DeviceBatch(scene, big, ab.default_opts())
    dev_small = ab.DeviceBatch(scene, small, ab.default_opts())
    assert dev_big.smem_bytes_per_warp > dev_small.smem_bytes_per_warp
    dev_small.launch()
    dev_big.launch()
    ref = ob.predict_batch(scene, big, ab.default_opts(), n_threads=8)
    assert [r.as_tuple() for r in dev_big.results()] == [r.as_tuple() for r in ref]


def test_predict_sequence_chains_actions(built):
    """BASELINE.json configs[3]: an action sequence predicted action by action, each from the state
    the ranked-first candidate of the previous one leaves (examples/ExampleActionsECS.cpp:924-979)."""
    from chain_util import scene_after_get
    scene = ab.HostScene()
    steps = scene.predict_sequence("get italian_seasoning_bowl; put italian_seasoning_bowl table; pose Default")
    assert [n for _, n, _ in steps] == [2, 6, 1]
    assert all(r.success == 1 for r, _, _ in steps) and all(m.startswith("SUCCESS") for _, _, m in steps)
    # the same chain by hand, with the oracle: same winner records
    manual = ab.HostScene()
    for (rec, n, _), cmd in zip(steps, ("get italian_seasoning_bowl", "put italian_seasoning_bowl table", "pose Default")):
        b = ab.HostBatch(manual)
        assert b.add_action(cmd) == n
        opts = ab.default_opts(log_q=True)
        cands = [ob.predict(manual, b, i, opts, log_q=True) for i in range(n)]
        best = ab.rank([c[0] for c in cands])[0]
        assert rec.idx == best and rec.as_tuple() == cands[best][0].as_tuple()
        manual.apply_rollout(b, best, cands[best][1])
    # the sequence the scene file ships: the restated get of the bottle misses the tracking limit, so
    # the rest of the queue is dropped, as the reference does after a failed action
    shipped = ab.HostScene().predict_sequence("get tomato_sauce_bottle; pour tomato_sauce_bottle pizza_dough; put tomato_sauce_bottle shelf")
    assert len(shipped) == 1 and shipped[0][0].success == 0 and "Reachability problem" in shipped[0][2]


@pytest.mark.parametrize("dt,after,alpha,lam", [(0.02, 100, 0.05, 1.0e-6), (0.005, 0, 0.1, 1.0e-4)])
def test_other_predictor_options_match_oracle(scene, dt, after, alpha, lam):
    """Nothing in the kernel is tied to the default dt / after-steps / alpha / lambda."""
    batch = ab.HostBatch(scene)
    lib = ab.lib()
    assert lib.aff_host_batch_add_action(batch.h, scene.h, b"get fresh_tomatoes", 1.0, dt) == 2
    assert lib.aff_host_batch_add_action(batch.h, scene.h, b"get italian_seasoning_bowl", 1.0, dt) == 2
    opts = ab.default_opts(log_q=True)
    opts.dt, opts.n_after_steps, opts.alpha, opts.lam = dt, after, alpha, lam
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    assert dev.num_steps(0) == int(round((10.0 + 5 * dt + 0.5) / dt)) + after      # tasks switch off 0.5 s after t_end
    assert _compare(scene, batch, opts, dev, range(4)) == 0.0


def test_invalid_initial_state_is_reported_like_the_oracle(built):
    """check() before the first step (src/TrajectoryPredictor.cpp:135-140, 426-441): a start state
    with a joint beyond its limit ends the rollout with zero steps, on the GPU as in the oracle."""
    scene = ab.HostScene()
    b = ab.HostBatch(scene)
    b.add_action("pose Default")
    opts = ab.default_opts(log_q=True)
    _, q = ob.predict(scene, b, 0, opts, log_q=True)
    start = q[0].copy()
    finger = [j for j in range(scene.nJ) if "finger_1_right" in scene.joint_name(j)][0]
    start[finger] = 5.0                                         # far beyond the finger's range
    scene.set_ik_state(start)
    batch = ab.HostBatch(scene)
    n = batch.add_action("get fresh_tomatoes")
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    gpu = dev.results()
    for i in range(n):
        ref, _ = ob.predict(scene, batch, i, opts)
        assert ref.code == -6 and ref.n_steps == 0             # AFF_CODE_INITIAL_STATE
        assert gpu[i].as_tuple() == ref.as_tuple()
    _, msgs = batch.predict()
    assert all(m.startswith("FATAL_ERROR: Initial state is invalid") for m in msgs)


@pytest.mark.parametrize("cmd", ["get fresh_tomatoes duration 0.5", "get fresh_tomatoes duration 2", "get italian_seasoning_bowl duration 1.5",
                                 "get tomato_sauce_bottle duration 30", "get fresh_tomatoes liftHeight 0.3"])
def test_get_options_match_oracle(scene, cmd):
    """Command options change the step count and the trajectory (too short a duration fails the
    tracking limit within a few steps and the rollout still runs to its end; 30 s = 3555 steps)."""
    batch = ab.HostBatch(scene)
    n = batch.add_action(cmd)
    opts = ab.default_opts(log_q=True)
    dev = ab.DeviceBatch(scene, batch, opts)
    dev.launch()
    assert _compare(scene, batch, opts, dev, range(n)) == 0.0


def test_predict_batch_multi_splits_statically(scene):
    """aff_predict_batch_multi: one host process, candidates split [r n/G, (r+1) n/G) over the listed devices
    (here the same GPU three times: three host threads, three private streams), records in candidate order."""
    batch = ab.HostBatch(scene)
    n = batch.add_perturbed("get fresh_tomatoes", 50, seed0=777)
    opts = ab.default_opts()
    one = ab.predict_batch(scene, batch, opts)
    multi = ab.predict_batch_multi(scene, batch, [0, 0, 0], opts)
    assert [r.idx for r in multi] == list(range(n))
    assert [r.as_tuple() for r in multi] == [r.as_tuple() for r in one]
    with pytest.raises(ab.AffError, match="device"):
        ab.predict_batch_multi(scene, batch, [0, 99], opts)


def test_kernel_is_chosen_by_batch_size(scene, kernel, monkeypatch):
    """No AFF_KERNEL in the environment: a handful of candidates run one per block, a few hundred and more one
    per warp, sweeps one per thread."""
    if kernel != "sweep":
        pytest.skip("kernel selection is only exercised once")
    monkeypatch.delenv("AFF_SWEEP_MIN", raising=False)
    monkeypatch.delenv("AFF_CTA_MAX", raising=False)
    opts = ab.default_opts()
    kinds = []
    for n in (2, 600, 30000):
        batch = ab.HostBatch(scene)
        batch.add_perturbed("get fresh_tomatoes", n, seed0=5)
        kinds.append(ab.DeviceBatch(scene, batch, opts).kernel)
    assert kinds == ["cta", "warp", "sweep"]


def test_large_batch_runs_on_the_sweep_kernel_and_matches(scene, kernel, monkeypatch):
    """A batch above AFF_SWEEP_MIN selects the thread-per-candidate kernel by itself; 4 blocks of 512 with a
    ragged tail (2000 candidates), sampled against the oracle."""
    if kernel != "sweep":
        pytest.skip("kernel selection is only exercised once")
    monkeypatch.setenv("AFF_SWEEP_MIN", "1500")
    batch = ab.HostBatch(scene)
    n = batch.add_perturbed("get italian_seasoning_bowl", 2000, seed0=31337)
    opts = ab.default_opts()
    dev = ab.DeviceBatch(scene, batch, opts)
    assert dev.kernel == "sweep"
    dev.launch()
    gpu = dev.results()
    assert [r.idx for r in gpu] == list(range(n))
    for i in list(range(0, n, 97)) + [n - 1]:
        ref, _ = ob.predict(scene, batch, i, opts)
        assert gpu[i].as_tuple() == ref.as_tuple()


def test_early_exit_matches_oracle_on_the_gpu(scene):
    batch = ab.HostBatch(scene)
    n = batch.add_perturbed("get italian_seasoning_bowl", 40, seed0=1234)
    opts = ab.default_opts(early_exit=True)
    gpu = ab.predict_batch(scene, batch, opts)
    ref = ob.predict_batch(scene, batch, opts, n_threads=os.cpu_count() or 1)
    assert [r.as_tuple() for r in gpu] == [r.as_tuple() for r in ref]
    assert any(r.n_steps < 1556 for r in gpu) and any(r.n_steps == 1556 for r in gpu)


def test_put_120_candidates_match_oracle(built):
    """BASELINE.json configs[1] at its stated size: `put block table` on the reference's unit-test table with
    120 Supportables (config/xml/unittest/g_scenario_unittest_predict_many.xml:157-160,
    unittest/resources/g_table_multi.xml): every Capability x Affordance candidate, all fields, and the ranking."""
    from chain_util import scene_many_after_get
    scene = scene_many_after_get()
    batch = ab.HostBatch(scene)
    assert batch.add_action("put block table") == 120
    opts = ab.default_opts()
    gpu = ab.predict_batch(scene, batch, opts)
    ref = ob.predict_batch(scene, batch, opts, n_threads=os.cpu_count() or 1)
    assert [r.as_tuple() for r in gpu] == [r.as_tuple() for r in ref]
    assert ab.rank(gpu) == ab.rank(ref)
    assert len(set(r.first_code for r in ref)) >= 2


def test_compiled_c_client_predicts(scene):
    """The plain C client of tests/c_client (gcc, no Python in the call path) through aff_host_predict."""
    import subprocess
    from test_abi import _build_c_client
    exe = _build_c_client()
    out = subprocess.run([exe, ab.PIZZA_SCENE, "get fresh_tomatoes", "--predict"], check=True, capture_output=True, text=True).stdout
    rows = [l for l in out.splitlines()[1:] if "|" in l]
    batch = ab.HostBatch(scene)
    batch.add_action("get fresh_tomatoes")
    ref = ob.predict_batch(scene, batch, ab.default_opts(), n_threads=2)
    ranked = [ref[i] for i in ab.rank(ref)]
    assert len(rows) == 2
    for line, r in zip(rows, ranked):
        f = line.split("|")[0].split()
        assert (int(f[0]), int(f[1]), int(f[2]), int(f[3])) == (r.idx, r.success, r.code, r.first_fail_step)
        assert float(f[4]) == r.jl_cost + r.coll_cost


def test_cpp_shim_client_predicts(scene):
    """tests/cpp_client/shim_client.cpp (INTEGRATION.md section 2 as a program, next to classes named like the
    reference's) through affgpu::TrajectoryPredictor::predictBatch."""
    import subprocess
    from test_abi import _build_cpp_client
    exe = _build_cpp_client()
    out = subprocess.run([exe, ab.PIZZA_SCENE, "get fresh_tomatoes", "--predict"], check=True, capture_output=True, text=True).stdout
    rows = [l for l in out.splitlines() if "|" in l]
    batch = ab.HostBatch(scene)
    batch.add_action("get fresh_tomatoes")
    ref = ob.predict_batch(scene, batch, ab.default_opts(), n_threads=2)
    ranked = [ref[i] for i in ab.rank(ref)]
    assert len(rows) == 2
    for line, r in zip(rows, ranked):
        f = line.split("|")[0].split()
        assert (int(f[0]), int(f[1])) == (r.idx, r.success) and float(f[2]) == r.jl_cost + r.coll_cost
