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


@pytest.mark.parametrize("cmd,cand", [("get fresh_tomatoes", 0), ("get italian_seasoning_bowl", 1)])
def test_emulated_kernel_equals_oracle(scene, cmd, cand):
    b = ab.HostBatch(scene)
    b.add_action(cmd)
    opts = ab.default_opts(log_q=True)
    ref, q_ref = ob.predict(scene, b, cand, opts, log_q=True)
    (emu,), q_emu = eb.predict(scene, b, opts, first=cand, count=1, log_q=True, n_steps=ref.n_steps - 1)
    assert emu.as_tuple() == ref.as_tuple()
    assert np.array_equal(q_emu[0], q_ref)


def test_emulated_kernel_perturbed_candidate(scene):
    b = ab.HostBatch(scene)
    b.add_perturbed("get italian_seasoning_bowl", 2, seed0=1234)
    opts = ab.default_opts()
    ref, _ = ob.predict(scene, b, 1, opts)
    (emu,), _ = eb.predict(scene, b, opts, first=1, count=1)
    assert ref.code == -4                     # a colliding candidate with several pairs in contact
    assert emu.as_tuple() == ref.as_tuple()


def test_emulated_kernel_put_candidate(built):
    """`put` after `get`: object starts in the hand (six rigid-body dofs under a moving parent),
    refFrame tasks, BoxInterval task regions, re-parenting onto a static surface."""
    from chain_util import scene_after_get
    scene = scene_after_get("italian_seasoning_bowl")
    b = ab.HostBatch(scene)
    assert b.add_action("put italian_seasoning_bowl table") == 6
    opts = ab.default_opts(log_q=True)
    ref, q_ref = ob.predict(scene, b, 4, opts, log_q=True)        # the one candidate that collides
    (emu,), q_emu = eb.predict(scene, b, opts, first=4, count=1, log_q=True, n_steps=ref.n_steps - 1)
    assert ref.code == -4
    assert emu.as_tuple() == ref.as_tuple()
    assert np.array_equal(q_emu[0], q_ref)
