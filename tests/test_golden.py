"""Golden fixtures (tests/golden/predict_get.json, made by tools/make_golden.py from the
CPU oracle — the reference itself holds no numeric fixtures for this path): the
oracle must keep reproducing them (CPU tier) and the CUDA path must reproduce
them bit for bit (GPU tier)."""
import json
import os

import pytest

import affaction_b200 as ab
import oracle_binding as ob

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "predict_get.json")))


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
