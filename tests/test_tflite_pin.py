"""The pin against REAL TensorFlow-Lite (the reference's numeric anchor, tflite_prediction.py:23-41).

No TFLite runtime exists in the build container or on this pool's GPU boxes (probed at the start of round 2:
tensorflow, tflite_runtime, ai_edge_litert absent; profiles/r02_tflite_probe.txt), so these tests are opportunistic:
  * tests/golden/tflite_ref.npz, once produced by tools/dump_tflite_reference.py on any machine with TFLite and
    committed, is asserted against the oracle tensor by tensor (and against the CUDA path on a GPU box);
  * on every run the import is attempted, and when it succeeds the live reference-kernel interpreter is compared.
Until one of the two exists the oracle's network arithmetic stays "parity unpinned"  -- the skips
below print that reason instead of passing silently."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "tests", "golden", "tflite_ref.npz")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import dump_tflite_reference as dt  # noqa: E402

UNPINNED = "parity unpinned: no TFLite runtime importable and no tests/golden/tflite_ref.npz committed"


def test_dump_script_degrades_cleanly():
    make, why = dt.find_interpreter()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dump_tflite_reference.py")], capture_output=True, text=True)
    if make is None:
        assert r.returncode == 3 and "nothing written" in r.stdout, (r.returncode, r.stdout, r.stderr)
    else:
        assert r.returncode == 0, r.stderr


def test_dump_inputs_are_the_oracles_vectors():
    from oracle_lib import vector_a, vector_b
    assert np.array_equal(dt.vector_a(), vector_a()) and np.array_equal(dt.vector_b(), vector_b())


@pytest.mark.skipif(not os.path.exists(DUMP), reason=UNPINNED)
def test_oracle_equals_committed_tflite_dump(oracle):
    d = np.load(DUMP)
    assert str(d["resolver"]) == "BUILTIN_REF"
    heads = oracle.run_batch(d["inputs"], threads=4)
    assert np.array_equal(heads, d["heads_ref"])
    for tag, x in (("A", d["inputs"][0]), ("B", d["inputs"][1])):
        _, outs = oracle.run(x, dump=True)
        for i in range(oracle.num_ops):
            key = "%s_t%d" % (tag, oracle.op(i)["output"])
            if key in d:
                assert np.array_equal(outs[i], d[key]), (tag, i)
    for size in (112, 224):
        rep = size // 56
        big = np.stack([np.tile(x, (rep, rep, 1)) for x in d["inputs"][:4]])
        assert np.array_equal(oracle.run_batch(big, threads=4), d["heads_%d" % size])


def test_oracle_equals_live_tflite(oracle, golden):
    from oracle_lib import vector_a, vector_b
    inputs = np.concatenate([np.stack([vector_a(), vector_b()]), golden["images"]])
    live = dt.live_heads(inputs, reference_kernels=True)
    if live is None:
        pytest.skip(UNPINNED + " (%s)" % dt.find_interpreter()[1])
    assert np.array_equal(oracle.run_batch(inputs, threads=4), live)


@pytest.mark.gpu
def test_gpu_equals_tflite_on_this_box(golden):
    """On the GPU box: try the import again (the box may carry wheels the build container lacks) and compare the CUDA
    path with the live interpreter and/or the committed dump."""
    import pkg
    from oracle_lib import vector_a, vector_b
    inputs = np.concatenate([np.stack([vector_a(), vector_b()]), golden["images"]])
    live = dt.live_heads(inputs, reference_kernels=True)
    ref = live if live is not None else (np.load(DUMP)["heads_ref"] if os.path.exists(DUMP) else None)
    if ref is None:
        pytest.skip(UNPINNED + " on this box (%s)" % dt.find_interpreter()[1])
    yf = pkg.load()
    net = yf.Network(device=0)
    try:
        assert np.array_equal(net.ai_run(inputs), ref)
    finally:
        net.close()
