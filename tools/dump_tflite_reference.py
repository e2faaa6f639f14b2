#!/usr/bin/env python
"""One-shot dump of REAL TensorFlow-Lite outputs for the yoloface int8 model.

The reference's numeric anchor is `tf.lite.Interpreter` (yoloface/tflite/tflite_prediction.py:23-41).  No TFLite
runtime exists in the build container or on the GPU boxes of this pool (probed: tensorflow, tflite_runtime,
ai_edge_litert all absent), so the oracle is pinned to everything else the reference offers.  Run this script on ANY
machine that has one of

    pip install tensorflow==2.10.0        (the version the reference pins, yoloface/tensorflow/requirements.txt:2)
    pip install tflite-runtime            or            pip install ai-edge-litert

and commit the file it writes, tests/golden/tflite_ref.npz: tests/test_tflite_pin.py then asserts the C oracle (and,
on a GPU box, the CUDA path) against it tensor by tensor, and DESIGN.md's "parity unpinned" can be struck.

Contents of tflite_ref.npz (inputs: Appendix-B vectors A and B + the 27 calibration images of tests/golden/images_56.npy):
  resolver                 'BUILTIN_REF' (reference kernels; the anchor) -- heads from the default resolver are stored too
  inputs      [29,56,56,3] int8
  heads_ref   [29,7,7,18]  int8   reference-kernel interpreter
  heads_opt   [29,7,7,18]  int8   default (optimised, XNNPACK off) interpreter
  A_t<idx>, B_t<idx>               every int8 activation tensor of the graph for vectors A and B
                                   (experimental_preserve_all_tensors=True), keyed by TFLite tensor index
  heads_112, heads_224             heads of the first four inputs tiled to 112x112 / 224x224 via resize_tensor_input
  versions                 runtime name + version
Exit codes: 0 written, 3 no TFLite runtime importable (nothing written).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL = os.path.join(ROOT, "stm32h7-yolo_b200", "assets", "yoloface_int8.tflite")
OUT = os.path.join(ROOT, "tests", "golden", "tflite_ref.npz")


def find_interpreter():
    """-> (make(model_path, reference_kernels: bool, preserve: bool) -> interpreter, 'name version') or (None, why)."""
    tried = []
    try:
        import tensorflow as tf

        def make(path, ref, preserve=False):
            kw = {}
            if ref:
                kw["experimental_op_resolver_type"] = tf.lite.experimental.OpResolverType.BUILTIN_REF
            else:
                kw["experimental_op_resolver_type"] = tf.lite.experimental.OpResolverType.BUILTIN_WITHOUT_DEFAULT_DELEGATES
            if preserve:
                kw["experimental_preserve_all_tensors"] = True
            return tf.lite.Interpreter(model_path=path, **kw)
        return make, "tensorflow " + tf.__version__
    except Exception as e:                                   # noqa: BLE001 -- any import failure means "not available"
        tried.append("tensorflow: %s" % type(e).__name__)
    for modname in ("tflite_runtime.interpreter", "ai_edge_litert.interpreter"):
        try:
            mod = __import__(modname, fromlist=["Interpreter"])

            def make(path, ref, preserve=False, mod=mod):
                kw = {}
                rt = getattr(mod, "OpResolverType", None)
                if rt is not None:
                    kw["experimental_op_resolver_type"] = rt.BUILTIN_REF if ref else rt.BUILTIN_WITHOUT_DEFAULT_DELEGATES
                elif ref:
                    raise RuntimeError("this runtime cannot select the reference kernels")
                if preserve:
                    kw["experimental_preserve_all_tensors"] = True
                return mod.Interpreter(model_path=path, **kw)
            return make, modname.split(".")[0] + " " + getattr(__import__(modname.split(".")[0]), "__version__", "?")
        except Exception as e:                               # noqa: BLE001
            tried.append("%s: %s" % (modname, type(e).__name__))
    return None, "; ".join(tried)


def vector_a():
    i = np.arange(56 * 56 * 3, dtype=np.int64)
    return (((37 * i + 11) % 256) - 128).astype(np.int8).reshape(56, 56, 3)


def vector_b():
    s = 12345; x = np.empty(56 * 56 * 3, np.int8)
    for k in range(x.size):
        s = (s * 1103515245 + 12345) & 0x7FFFFFFF
        x[k] = ((s >> 16) & 0xFF) - 128
    return x.reshape(56, 56, 3)


def run_heads(interp, inputs):
    i_idx = interp.get_input_details()[0]["index"]; o_idx = interp.get_output_details()[0]["index"]
    out = []
    for x in inputs:
        interp.set_tensor(i_idx, x[None])                    # tflite_prediction.py:39
        interp.invoke()                                      # :40
        out.append(interp.get_tensor(o_idx)[0].copy())       # :41
    return np.stack(out)


def live_heads(inputs, reference_kernels=True, size=None):
    """Heads of `inputs` from a live interpreter, or None when no runtime is importable (used by the tests and bench)."""
    make, _ = find_interpreter()
    if make is None:
        return None
    it = make(MODEL, reference_kernels)
    if size is not None:
        it.resize_tensor_input(it.get_input_details()[0]["index"], [1, size, size, 3])
    it.allocate_tensors()
    return run_heads(it, inputs)


def main():
    make, ver = find_interpreter()
    if make is None:
        print("no TFLite runtime importable (%s): nothing written" % ver)
        return 3
    imgs = np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))
    inputs = np.concatenate([np.stack([vector_a(), vector_b()]), imgs]).astype(np.int8)
    d = {"inputs": inputs, "versions": np.array(ver), "resolver": np.array("BUILTIN_REF")}
    ref = make(MODEL, True); ref.allocate_tensors()
    d["heads_ref"] = run_heads(ref, inputs)
    opt = make(MODEL, False); opt.allocate_tensors()
    d["heads_opt"] = run_heads(opt, inputs)
    full = make(MODEL, True, preserve=True); full.allocate_tensors()
    i_idx = full.get_input_details()[0]["index"]
    for tag, x in (("A", inputs[0]), ("B", inputs[1])):
        full.set_tensor(i_idx, x[None]); full.invoke()
        for t in full.get_tensor_details():
            if t["dtype"] == np.int8 and len(t["shape"]) == 4 and t["shape"][0] == 1:
                try:
                    d["%s_t%d" % (tag, t["index"])] = full.get_tensor(t["index"])[0].copy()
                except ValueError:
                    pass                                      # constant without a buffer in this runtime
    for size in (112, 224):
        rep = size // 56
        big = np.stack([np.tile(x, (rep, rep, 1)) for x in inputs[:4]])
        d["heads_%d" % size] = live_heads(big, True, size)
    np.savez_compressed(OUT, **d)
    print("wrote %s (%s): %d arrays, heads_ref == heads_opt: %s" % (OUT, ver, len(d), np.array_equal(d["heads_ref"], d["heads_opt"])))
    return 0


if __name__ == "__main__":
    sys.exit(main())
