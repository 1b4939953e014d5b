"""CPU tier: the kernel source (rollout.cuh) compiled for the host through the
test-only warp emulator must reproduce the oracle bit for bit.  This checks the
plan compiler and the kernel's logic without a GPU; the GPU parity tests proper
are in test_gpu_parity.py."""
import numpy as np
import pytest

import affaction_b200 as ab
import emu_binding as eb
import oracle_binding as ob


@pytest.fixture(scope="module")
def scene(built):
    return ab.HostScene()


@pytest.fixture(params=[False, True, "cta"], ids=["warp", "sweep", "cta"])
def sweep(request):
    """False: rollout.cuh (one rollout per warp); True: sweep.cuh (one rollout per thread); "cta": rollout.cuh
    in the block-per-candidate arrangement (warp 0 walks the steps, three warps evaluate the collision model)."""
    return request.param


@pytest.mark.parametrize("cmd,cand", [("get fresh_tomatoes", 0), ("get italian_seasoning_bowl", 1)])
def test_emulated_kernel_equals_oracle(scene, sweep, cmd, cand):
    b = ab.HostBatch(scene)
    b.add_action(cmd)
    opts = ab.default_opts(log_q=True)
    ref, q_ref = ob.predict(scene, b, cand, opts, log_q=True)
    (emu,), q_emu = eb.predict(scene, b, opts, first=cand, count=1, log_q=True, n_steps=ref.n_steps - 1, sweep=sweep)
    assert emu.as_tuple() == ref.as_tuple()
    assert np.array_equal(q_emu[0], q_ref)


def test_emulated_kernel_perturbed_candidate(scene, sweep):
    b = ab.HostBatch(scene)
    b.add_perturbed("get italian_seasoning_bowl", 2, seed0=1234)
    opts = ab.default_opts()
    ref, _ = ob.predict(scene, b, 1, opts)
    (emu,), _ = eb.predict(scene, b, opts, first=1, count=1, sweep=sweep)
    assert ref.code == -4                     # a colliding candidate with several pairs in contact
    assert emu.as_tuple() == ref.as_tuple()


def test_emulated_kernel_put_candidate(built, sweep):
    """`put` after `get`: object starts in the hand (six rigid-body dofs under a moving parent),
    refFrame tasks, BoxInterval task regions, re-parenting onto a static surface."""
    from chain_util import scene_after_get
    scene = scene_after_get("italian_seasoning_bowl")
    b = ab.HostBatch(scene)
    assert b.add_action("put italian_seasoning_bowl table") == 6
    opts = ab.default_opts(log_q=True)
    ref, q_ref = ob.predict(scene, b, 4, opts, log_q=True)        # the one candidate that collides
    (emu,), q_emu = eb.predict(scene, b, opts, first=4, count=1, log_q=True, n_steps=ref.n_steps - 1, sweep=sweep)
    assert ref.code == -4
    assert emu.as_tuple() == ref.as_tuple()
    assert np.array_equal(q_emu[0], q_ref)


def test_sweep_kernel_source_on_a_mixed_batch(scene):
    """sweep.cuh on 24 perturbed candidates of two hands (two structure classes: the plan compiler
    groups them, results come back in candidate order) against the oracle, record by record."""
    b = ab.HostBatch(scene)
    n = b.add_perturbed("get fresh_tomatoes", 24, seed0=4242)
    opts = ab.default_opts()
    ref = ob.predict_batch(scene, b, opts, n_threads=4)
    emu, _ = eb.predict(scene, b, opts, sweep=True)
    assert [r.as_tuple() for r in emu] == [r.as_tuple() for r in ref]
    assert len(set(r.code for r in ref)) >= 2


def test_early_exit_is_opt_in_and_keeps_verdicts(scene, sweep):
    """opts.early_exit (the reference's commented-out break, src/TrajectoryPredictor.cpp:368-373): a failed
    rollout stops at the end of its failing step; verdict, code and failure step are those of the full run,
    and the kernel source agrees with the oracle on every field."""
    b = ab.HostBatch(scene)
    n = b.add_perturbed("get italian_seasoning_bowl", 4, seed0=1234)
    full, early = ab.default_opts(), ab.default_opts(early_exit=True)
    stopped = 0
    for i in range(n):
        r0, _ = ob.predict(scene, b, i, full)
        r1, _ = ob.predict(scene, b, i, early)
        (e1,), _ = eb.predict(scene, b, early, first=i, count=1, sweep=sweep)
        assert e1.as_tuple() == r1.as_tuple()
        assert (r0.success, r0.first_code, r0.first_fail_step) == (r1.success, r1.first_code, r1.first_fail_step)
        if not r1.success:
            assert r1.n_steps == r1.first_fail_step + 2 < r0.n_steps      # rows = initial row + steps up to the failing one
            stopped += 1
        else:
            assert r1.as_tuple() == r0.as_tuple()
    assert stopped >= 1


def test_put_on_the_120_supportable_table(built, sweep):
    """The reference's 120-Supportable unit-test scene (second flattened scene): candidate generation gives
    120 put candidates; a sample runs through the kernel source and equals the oracle."""
    from chain_util import scene_many_after_get
    scene = scene_many_after_get()
    b = ab.HostBatch(scene)
    assert b.add_action("put block table") == 120
    opts = ab.default_opts()
    for i in (0, 59, 119):
        ref, _ = ob.predict(scene, b, i, opts)
        (emu,), _ = eb.predict(scene, b, opts, first=i, count=1, sweep=sweep)
        assert emu.as_tuple() == ref.as_tuple()


def test_plan_compiler_threads_do_not_change_the_plan(scene, monkeypatch):
    """The per-candidate tables are built by several host threads over candidate ranges and merged in order:
    2 100 candidates (two ranges) give the same records as the single-threaded plan."""
    b = ab.HostBatch(scene)
    b.add_perturbed("get italian_seasoning_bowl", 2100, seed0=99)
    opts = ab.default_opts(early_exit=True)      # early exit keeps the 2 x 2 100 CPU rollouts short
    out = []
    for th in ("1", "2"):
        monkeypatch.setenv("AFF_PLAN_THREADS", th)
        res, _ = eb.predict(scene, b, opts, sweep=True)
        out.append([r.as_tuple() for r in res])
    assert out[0] == out[1]
    ref, _ = ob.predict(scene, b, 2099, opts)
    assert out[1][2099] == ref.as_tuple()


def test_fk_schedule_puts_the_step_loops_frames_first(scene):
    """FK schedule of the warp / block kernels: parents in earlier rounds, at most two nodes per round, and for a
    one-object `get` the frames the step loop reads fill only the first rounds (the block-per-candidate kernel
    leaves the others to its FK warp); the static tree is cut into levels for the prologue."""
    b = ab.HostBatch(scene)
    b.add_action("get fresh_tomatoes")
    info = eb.plan_info(scene, b, ab.default_opts())
    assert info["warp_ok"] == 1 and info["sweep_ok"] == 1
    assert 0 < info["n_rounds_crit"] < info["n_rounds"]
    assert info["n_stat_levels"] >= 2
    rounds = info["round_of"]
    for n, p in enumerate(info["parent_of"]):
        if p >= 0:
            assert rounds[p] < rounds[n]
    for r in range(info["n_rounds"]):
        assert 1 <= sum(1 for x in rounds if x == r) <= 2
    # two objects (double_get): both hang on hands that come last in their chains; whatever the split, it is sane
    b2 = ab.HostBatch(scene)
    b2.add_action("double_get tomato_sauce_bottle bbq_sauce_bottle")
    info2 = eb.plan_info(scene, b2, ab.default_opts())
    assert 0 < info2["n_rounds_crit"] <= info2["n_rounds"]
