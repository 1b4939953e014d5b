"""Golden fixtures (tests/golden/predict_get.json, made by tests/tools/make_golden.py from the
CPU oracle — the reference itself holds no numeric fixtures for this path): the
oracle must keep reproducing them (CPU tier) and the CUDA path must reproduce
them bit for bit (GPU tier)."""
import json
import os

import pytest

import affaction_b200 as ab
import oracle_binding as ob

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "predict_get.json")))
ACTIONS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "predict_actions.json")))


def _decode(fields, vals):
    return tuple(float.fromhex(v) if isinstance(v, str) else v for v in vals)


@pytest.fixture(scope="module")
def scene(built):
    return ab.HostScene()


@pytest.mark.parametrize("cmd", sorted(GOLDEN["commands"]))
def test_oracle_reproduces_golden(scene, cmd):
    b = ab.HostBatch(scene)
    n = b.add_action(cmd)
    gold = GOLDEN["commands"][cmd]
    assert n == len(gold)
    opts = ab.default_opts(log_q=True)
    for i in range(n):
        r, q = ob.predict(scene, b, i, opts, log_q=True)
        assert r.as_tuple() == _decode(GOLDEN["fields"], gold[i]["result"])
        for k, row in gold[i]["q"].items():
            assert [float.fromhex(v) for v in row] == list(q[int(k)])
        assert b.format_message(r) == gold[i]["message"]


def test_golden_covers_every_verdict_kind():
    codes = set()
    for cands in GOLDEN["commands"].values():
        for c in cands:
            codes.add(c["result"][GOLDEN["fields"].index("code")])
    assert {0, -3, -4, -5} <= codes          # success, joint limit, collision, tracking


@pytest.mark.gpu
@pytest.mark.parametrize("cmd", sorted(GOLDEN["commands"]))
def test_gpu_reproduces_golden(scene, cmd):
    b = ab.HostBatch(scene)
    n = b.add_action(cmd)
    gold = GOLDEN["commands"][cmd]
    opts = ab.default_opts(log_q=True)
    dev = ab.DeviceBatch(scene, b, opts)
    dev.launch()
    res = dev.results()
    for i in range(n):
        assert res[i].as_tuple() == _decode(GOLDEN["fields"], gold[i]["result"])
        q = dev.fetch_q(i)
        for k, row in gold[i]["q"].items():
            assert [float.fromhex(v) for v in row] == list(q[int(k)])


def _start_scene(setup):
    if not setup:
        return ab.HostScene()
    from chain_util import scene_after_get
    return scene_after_get(setup.split(":", 1)[1], require_success=False)


def _check_case(case, predict):
    scene = _start_scene(case["setup"])
    b = ab.HostBatch(scene)
    gold = case["candidates"]
    assert b.add_action(case["command"]) == len(gold)
    for i, g in enumerate(gold):
        r, q = predict(scene, b, i)
        assert r.as_tuple() == _decode(ACTIONS["fields"], g["result"])
        for k, row in g["q"].items():
            assert [float.fromhex(v) for v in row] == list(q[int(k)])
        assert b.format_message(r) == g["message"]


@pytest.mark.parametrize("case", ACTIONS["cases"], ids=[c["command"].replace(" ", "_") for c in ACTIONS["cases"]])
def test_oracle_reproduces_action_golden(built, case):
    """put / drop / pour / poke after a get, double_get, pose, finger_push: pinned records."""
    opts = ab.default_opts(log_q=True)
    _check_case(case, lambda scene, b, i: ob.predict(scene, b, i, opts, log_q=True))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ACTIONS["cases"], ids=[c["command"].replace(" ", "_") for c in ACTIONS["cases"]])
def test_gpu_reproduces_action_golden(built, case):
    opts = ab.default_opts(log_q=True)
    cache = {}

    def predict(scene, b, i):
        if "dev" not in cache:
            cache["dev"] = ab.DeviceBatch(scene, b, opts)
            cache["dev"].launch()
            cache["res"] = cache["dev"].results()
        return cache["res"][i], cache["dev"].fetch_q(i)

    _check_case(case, predict)
