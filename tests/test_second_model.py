"""SURVEY.md 8f n2 -- the model compiler is not wired to yoloface: a SECOND int8 .tflite of the same nine operator
types (tools/make_mini_tflite.py: widths 16/32/64, 5x5 and 3x3 pools, other slopes / scales / zero points, 48x64 input)
goes through the same planner, both CUDA paths and the decode, checked against the oracle executing that model."""
import os
import subprocess
import sys

import numpy as np
import pytest

import pkg
from oracle_lib import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINI2 = os.path.join(ROOT, "tests", "golden", "mini2_int8.tflite")


@pytest.fixture(scope="module")
def mini_oracle():
    return Oracle(MINI2)


@pytest.fixture(scope="module")
def yf():
    return pkg.load()


def test_fixture_is_what_the_generator_writes(tmp_path):
    out = str(tmp_path / "m.tflite")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_mini_tflite.py"), out], stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(MINI2, "rb").read()


def test_structure(mini_oracle):
    o = mini_oracle
    counts = {}
    for i in range(o.num_ops):
        c = o.op(i)["opcode"]; counts[c] = counts.get(c, 0) + 1
    assert counts == {34: 3, 3: 15, 98: 15, 4: 6, 17: 2, 0: 2, 114: 4, 2: 2}          # the nine types, other counts than yoloface
    widths = {o.tensor(o.op(i)["output"])["shape"][3] for i in range(o.num_ops) if o.op(i)["opcode"] == 3}
    assert widths == {8, 16, 18, 24, 32, 40, 48, 64}


@pytest.mark.parametrize("hw", [(48, 64), (32, 32), (64, 64), (8, 16)])
def test_planner_on_the_second_model(yf, mini_oracle, hw, monkeypatch):
    """Both lowerings (26-step... here 23-step layered plan, fused shared-memory program incl. image pairs) emulated on
    CPU from the planner's own tables reproduce the oracle on this model."""
    from fused_emulator import run_fused
    from plan_emulator import Emulator
    monkeypatch.setenv("YF_B200_TFLITE", MINI2)
    P, F = yf.plan(*hw), yf.fused_program(*hw)
    assert len(P["steps"]) == 23 and len(F["phases"]) == 23 and F["spec"] == 0        # not the compiled-in program
    rng = np.random.default_rng(hw[0] * 100 + hw[1])
    a, b = rng.integers(-128, 128, (2,) + hw + (3,), dtype=np.int8)
    want_a, want_b = mini_oracle.run(a), mini_oracle.run(b)
    em = Emulator(P)
    assert np.array_equal(em.tensor(em.run(a, observer=False), mini_oracle.lib.yfo_output_tensor(mini_oracle.m)), want_a)
    assert np.array_equal(run_fused(F, a, 1).reshape(want_a.shape), want_a)
    ha, hb = run_fused(F, [a, b], 2)
    assert np.array_equal(ha.reshape(want_a.shape), want_a) and np.array_equal(hb.reshape(want_b.shape), want_b)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "layered"])
def test_second_model_on_gpu(yf, mini_oracle, mode):
    net = yf.Network(tflite_path=MINI2, chunk_images=64, mode=mode)
    try:
        net.set_input_size(48, 64)
        assert net.stats()["fused"] == (1 if mode == "fused" else 0) and net.stats()["steps"] == 23
        rng = np.random.default_rng(7)
        x = rng.integers(-128, 128, (150, 48, 64, 3), dtype=np.int8)
        want = mini_oracle.run_batch(x, threads=os.cpu_count())
        assert np.array_equal(net.run(x), want)
        assert np.array_equal(net.run(x[:1]), want[:1]) and np.array_equal(net.run(x[:33]), want[:33])
        # the decode applies to this model's 6x8x18 head with ITS output quantisation
        to = mini_oracle.tensor(mini_oracle.lib.yfo_output_tensor(mini_oracle.m))
        dets, counts = net.decode(want[:16], 0.6, 0.4, False, max_det=144)
        for i in range(16):
            ref = mini_oracle.decode_nms(want[i], 0.6, 0.4, False, scale=to["scale"][0], zp=int(to["zp"][0]))
            assert counts[i] == len(ref)
            if len(ref):
                got = dets[i, :counts[i]]
                fin = np.isfinite(ref[:, :4])               # random weights: exp() of a large logit overflows to inf on both sides
                assert np.array_equal(np.isfinite(got[:, :4]), fin) and np.array_equal(got[:, :4][~fin], ref[:, :4][~fin])
                sc = np.maximum(1.0, np.abs(ref[:, :4][fin]))
                assert np.all(np.abs(got[:, :4][fin] - ref[:, :4][fin]) <= 1e-3 * sc) and np.all(np.abs(got[:, 4] - ref[:, 4]) <= 1e-5)
        # other resolutions of the same model
        for H, W in ((32, 32), (64, 64)) + (((96, 128),) if mode == "layered" else ()):   # 96x128 does not fit the fused kernel
            net.set_input_size(H, W)
            y = rng.integers(-128, 128, (5, H, W, 3), dtype=np.int8)
            out = net.run(y)
            for i in range(5):
                assert np.array_equal(out[i], mini_oracle.run(y[i])), (H, W, i)
        assert net.get_error() == (0, 0)
    finally:
        net.close()


@pytest.mark.gpu
def test_second_model_per_operator_tensors(yf, mini_oracle):
    """observer mode: every materialised operator output of the second model equals the oracle's"""
    net = yf.Network(tflite_path=MINI2, chunk_images=8, observer=True)
    try:
        net.set_input_size(48, 64)
        x = np.random.default_rng(11).integers(-128, 128, (3, 48, 64, 3), dtype=np.int8)
        net.run(x)
        checked = 0
        for i in range(mini_oracle.num_ops):
            t = mini_oracle.op(i)["output"]
            got = net.get_tensor(t, 3)
            if got is None:                        # PAD outputs / concat views are not materialised
                continue
            for k in range(3):
                _, outs = mini_oracle.run(x[k], dump=True)
                assert np.array_equal(got[k].reshape(outs[i].shape), outs[i]), (i, k)
            checked += 1
        assert checked >= 40
    finally:
        net.close()


@pytest.mark.gpu
def test_two_models_share_one_gpu(yf, mini_oracle, oracle, golden):
    """Per-plan tables (no device-wide __constant__ state on the fused path): alternating two models on one GPU needs
    no device synchronisation between them and keeps both bit-exact."""
    a = yf.Network(chunk_images=64)
    b = yf.Network(tflite_path=MINI2, chunk_images=64)
    try:
        b.set_input_size(48, 64)
        xb = np.random.default_rng(3).integers(-128, 128, (40, 48, 64, 3), dtype=np.int8)
        wb = mini_oracle.run_batch(xb, threads=os.cpu_count())
        for _ in range(4):
            assert np.array_equal(a.run(golden["images"]), golden["heads_images"])
            assert np.array_equal(b.run(xb), wb)
    finally:
        a.close(); b.close()
