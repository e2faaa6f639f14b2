#!/usr/bin/env python
"""Pin the oracle's head decode and NMS to the reference's OWN Python (run HERE, where /root/reference exists).

The reference holds three runnable Python statements of the post-processing (none importable as a module:
each file executes a demo / needs TensorFlow at import time), so they are lifted out with `ast`, unmodified:

  lift_tflite_decode()   yoloface/tflite/tflite_prediction.py:5-21,43-57  numpy decode the TFLite script applies to the
                         interpreter's int8 head (dequantise, sigmoid/exp, grid, anchors) + its threshold "NMS"
  lift_yolo_layer()      yoloface/pytorch/yoloface.py:288-366              torch `yolo_layer` (the float model's decode)
  lift_nms()             yoloface/tensorflow/yoloface_test.py:148-201      the only IoU-NMS in source form (greedy, +1 areas)

Outputs tests/golden/refpy_decode.npz:
  heads        [N,7,7,18] int8     the 27 image heads of the oracle + 200 seeded synthetic heads (distinct confidences per
                                   head so that sort ties -- numpy's argsort()[::-1] vs a stable sort -- cannot matter)
  tfl_xywhc    [N,147,5] float32   tflite_prediction.py decode of every candidate: cx, cy, w, h, conf, in ITS candidate
                                   order (anchor-major: a*49 + gy*7 + gx)
  torch_xywhc  [N,147,5] float32   yolo_layer.forward on the dequantised head (same order)
  nms_in       list of [k,5]       boxes handed to non_max_suppression: int-truncated x1,y1,x2,y2 (+ conf) of the
                                   candidates with conf >= 0.7, candidate (cell-major) order
  nms_keep     list of [m]         indices into nms_in the reference keeps (iou_threshold 0.4)
The tests (tests/test_refpy_pin.py) compare the C oracle with these arrays, and -- when /root/reference is present --
re-lift the functions and check them live as well.
"""
import ast
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("YF_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
SCALE, ZP = 0.14218327403068542, -15            # yoloface.c:116 / tflite_prediction.py:44
ANCHORS = [[9, 14], [12, 17], [22, 21]]          # yoloface.c:20 / tflite_prediction.py:47-49


def _source(rel):
    path = os.path.join(REF, rel)
    return open(path, encoding="utf-8").read(), path


def lift_nms(iou_threshold=0.4):
    """YoloFaceDetector.non_max_suppression (yoloface_test.py:148-201) bound to a stub `self`."""
    src, path = _source("yoloface/tensorflow/yoloface_test.py")
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "YoloFaceDetector":
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == "non_max_suppression":
                    mod = ast.Module(body=[f], type_ignores=[])
                    ns = {"np": np}
                    exec(compile(mod, path, "exec"), ns)
                    stub = types.SimpleNamespace(iou_threshold=iou_threshold)
                    return lambda boxes: ns["non_max_suppression"](stub, boxes)
    raise RuntimeError("non_max_suppression not found in " + path)


def lift_tflite_decode():
    """tflite_prediction.py: helper functions (:5-21) + the module-level decode statements (:43-57) as one function
    head_int8 [1,7,7,18] -> (all candidates [147,6] after decode, boxes kept by its threshold-only 'NMS')."""
    src, path = _source("yoloface/tflite/tflite_prediction.py")
    tree = ast.parse(src)
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    stmts, take = [], False
    for n in tree.body:
        seg = ast.get_source_segment(src, n) or ""
        if seg.startswith("output = output_data[0]"):
            take = True
        if take:
            stmts.append(n)
        if seg.startswith("boxes = non_max_suppression"):
            break
    assert stmts and len(funcs) == 3, "tflite_prediction.py layout changed"
    code = compile(ast.Module(body=funcs + stmts, type_ignores=[]), path, "exec")

    def run(head):
        ns = {"np": np, "output_data": np.asarray(head, np.int8).reshape(1, 7, 7, 18)}
        exec(code, ns)
        return np.asarray(ns["output"], np.float32), np.asarray(ns["boxes"], np.float32).reshape(-1, 4)
    return run


def lift_yolo_layer():
    """class yolo_layer (pytorch/yoloface.py:288-366), instantiated with the reference's anchors."""
    import torch
    from itertools import chain
    src, path = _source("yoloface/pytorch/yoloface.py")
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "yolo_layer"]
    assert cls, "yolo_layer not found"
    ns = {"torch": torch, "nn": torch.nn, "chain": chain}
    exec(compile(ast.Module(body=cls, type_ignores=[]), path, "exec"), ns)
    layer = ns["yolo_layer"](ANCHORS)

    def run(head):
        deq = (np.asarray(head, np.float32) - ZP) * np.float32(SCALE)          # [7,7,18] NHWC
        x = torch.from_numpy(deq.transpose(2, 0, 1).copy()).unsqueeze(0)       # the float model's NCHW output
        with torch.no_grad():
            return layer(x, 56).numpy().astype(np.float32)                      # [147, 6], anchor-major
    return run


def synthetic_heads(n, seed):
    """Plausible heads: every confidence of a head distinct, w/h logits bounded so boxes stay finite, x/y uniform.
    About 40 % of the candidates pass 0.7."""
    rng = np.random.default_rng(seed)
    h = rng.integers(-128, 128, (n, 49, 3, 6), dtype=np.int64)
    h[..., 2:4] = rng.integers(-64, 17, (n, 49, 3, 2))
    for i in range(n):
        # 147 distinct bytes from [-128, 80]: above ~80 float32 sigmoid saturates and distinct bytes give EQUAL confidences
        h[i, :, :, 4] = (rng.permutation(209)[:147] - 128).reshape(49, 3)
    # a third of the heads: only a handful of confident candidates (the realistic regime)
    for i in range(0, n, 3):
        conf = h[i, :, :, 4].reshape(-1)
        low = conf.argsort()[:-6]
        conf[low] = np.minimum(conf[low], -20)
        h[i, :, :, 4] = conf.reshape(49, 3)
    return h.reshape(n, 7, 7, 18).astype(np.int8)


def nms_inputs(cands_cellmajor):
    """What yoloface_test.py:129-143 hands to non_max_suppression, for w_scale = h_scale = 1 and without its clamp to
    the image (the clamp is a display step): int-truncated corners + conf of the candidates with conf >= 0.7."""
    out = []
    for x1, y1, x2, y2, c in cands_cellmajor:
        if c < 0.7:
            continue
        out.append([int(x1), int(y1), int(x2), int(y2), float(c)])
    return out


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle                       # only to obtain real image heads as inputs
    imgs_heads = np.load(os.path.join(OUT, "oracle_heads.npz"))["heads_images"]
    heads = np.concatenate([imgs_heads, synthetic_heads(int(os.environ.get("YF_REFPY_HEADS", "200")), 20261018)])
    tfl, yl, nms = lift_tflite_decode(), lift_yolo_layer(), lift_nms()
    o = Oracle()
    tfl_all, torch_all, nms_in, nms_keep = [], [], [], []
    for h in heads:
        cand, _ = tfl(h)
        tfl_all.append(cand[:, :5])
        torch_all.append(yl(h)[:, :5])
        # NMS input: the reference's own decode (tflite_prediction order a*49+cell -> cell-major), corners as in :129-132
        c = cand.reshape(3, 49, 6).transpose(1, 0, 2).reshape(147, 6)
        xyxy = np.stack([c[:, 0] - c[:, 2] / 2, c[:, 1] - c[:, 3] / 2, c[:, 0] + c[:, 2] / 2, c[:, 1] + c[:, 3] / 2, c[:, 4]], 1)
        boxes = nms_inputs(xyxy)
        kept = nms(boxes)
        keep_idx = [boxes.index(k) for k in kept]          # boxes are distinct (distinct confidences)
        nms_in.append(np.array(boxes, np.float64).reshape(-1, 5)); nms_keep.append(np.array(keep_idx, np.int32))
    lens = np.array([len(b) for b in nms_in], np.int32); klens = np.array([len(k) for k in nms_keep], np.int32)
    np.savez_compressed(os.path.join(OUT, "refpy_decode.npz"), heads=heads, tfl_xywhc=np.stack(tfl_all).astype(np.float32),
                        torch_xywhc=np.stack(torch_all).astype(np.float32),
                        nms_in=np.concatenate(nms_in) if len(nms_in) else np.zeros((0, 5)), nms_in_len=lens,
                        nms_keep=np.concatenate(nms_keep), nms_keep_len=klens)
    print("refpy_decode: %d heads, %d NMS inputs, %d kept; survivors per head max %d" % (len(heads), lens.sum(), klens.sum(), lens.max()))
    del o


if __name__ == "__main__":
    main()
