#!/usr/bin/env python
"""Generate tests/golden/* from the reference tree (run HERE, where /root/reference exists).

Fixtures extracted from the reference ("things that pin numbers"):
  st_fixtures.npz
    st_lut_ops      [17]       TFLite op index of each LEAKY_RELU (ST conv2d_N fuses op N+1)
    st_luts         [17,256]   ST's int8->int8 activation tables, network.c:2218..2902
    st_blob         [11304]    weight blob bytes, network_data.c:25-388 (ai_u64 little-endian)
    st_blob_offsets [48]       (weights, bias) offsets per conv in ST node order, network.c:3117-3263
    st_blob_ids     [24]       conv2d_N ids in the same order
    st_intq_names / st_intq_scale / st_intq_zp   per-tensor quantisation, network.c:663-1341
  images_56.npy     [27,56,56,3] int8   the 27 calibration JPEGs (yoloface/small_dataset) resized the
                                        way tflite_prediction.py:31-37 does (RGB, cv2.resize, -128)
Fixtures produced by the oracle on those inputs (regression pins, regenerated only deliberately):
  oracle_heads.npz  heads of vector A/B + the 27 images, per-op CRC32 of vector A.
"""
import glob
import os
import re
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("YF_REFERENCE", "/root/reference")
APP = os.path.join(REF, "stm32/X-CUBE-AI/App")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))


def st_fixtures():
    src = open(os.path.join(APP, "network.c"), encoding="latin-1").read()
    luts, ops = [], []
    for mm in re.finditer(r"conv2d_(\d+)_nl_params_data\[\] = \{([^}]*)\}", src):
        ops.append(int(mm.group(1)) + 1)
        luts.append(np.array([int(v) for v in mm.group(2).split(",")], np.int8))
    order = np.argsort(ops)
    ops = np.array(ops)[order]; luts = np.stack(luts)[order]
    assert luts.shape == (17, 256), luts.shape
    offs = {}
    for mm in re.finditer(r"conv2d_(\d+)_(weights|bias)_array\.data = AI_PTR\(weights_map\[0\] \+ (\d+)\)", src):
        offs[(int(mm.group(1)), mm.group(2))] = int(mm.group(3))
    ids = sorted({k[0] for k in offs})
    off_arr = np.array([[offs[(i, "weights")], offs[(i, "bias")]] for i in ids], np.int32)
    names, scales, zps = [], [], []
    for mm in re.finditer(r"AI_INTQ_INFO_LIST_OBJ_DECLARE\((\w+)_intq,[^;]*?AI_PACK_INTQ_SCALE\(([^)]*)\),\s*AI_PACK_INTQ_ZP\(([^)]*)\)\)\)", src, re.S):
        sc = [float(v.strip().rstrip("f")) for v in mm.group(2).split(",")]
        zp = [int(v) for v in mm.group(3).split(",")]
        if len(sc) == 1:                       # activation tensors only (per-channel weight lists skipped)
            names.append(mm.group(1)); scales.append(np.float32(sc[0])); zps.append(zp[0])
    data = open(os.path.join(APP, "network_data.c"), encoding="latin-1").read()
    body = data[data.index("s_network_weights_array_u64"):]
    body = body[body.index("{") + 1:body.index("}")]
    words = [int(v.strip().rstrip("U"), 16) for v in body.split(",") if v.strip()]
    blob = np.array(words, dtype="<u8").view(np.uint8)
    assert blob.size == 11304, blob.size
    np.savez_compressed(os.path.join(OUT, "st_fixtures.npz"), st_lut_ops=ops, st_luts=luts, st_blob=blob,
                        st_blob_offsets=off_arr, st_blob_ids=np.array(ids, np.int32),
                        st_intq_names=np.array(names), st_intq_scale=np.array(scales, np.float32),
                        st_intq_zp=np.array(zps, np.int32))
    print("st_fixtures: %d LUTs, blob crc %08x, %d convs, %d intq records" % (len(ops), zlib.crc32(blob.tobytes()), len(ids), len(names)))


def images():
    import cv2
    files = sorted(glob.glob(os.path.join(REF, "yoloface/small_dataset/*.jpg")))
    arr = []
    for f in files:
        img = cv2.imread(f)
        x = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)                     # tflite_prediction.py:31
        x = cv2.resize(x, (56, 56)).astype(np.float32)              # :35
        arr.append((x - 128).astype(np.int8))                       # :37-38
    arr = np.stack(arr)
    np.save(os.path.join(OUT, "images_56.npy"), arr)
    print("images_56:", arr.shape, "crc %08x" % zlib.crc32(arr.tobytes()))
    return arr


def oracle_pins(imgs):
    from oracle_lib import Oracle, vector_a, vector_b
    o = Oracle()
    ha, outs = o.run(vector_a(), dump=True)
    hb = o.run(vector_b())
    heads = o.run_batch(imgs, threads=4)
    crcs = np.array([zlib.crc32(t.tobytes()) for t in outs], np.uint32)
    np.savez_compressed(os.path.join(OUT, "oracle_heads.npz"), head_a=ha, head_b=hb, heads_images=heads, op_crc_a=crcs)
    print("oracle pins: head_a crc %08x head_b crc %08x images crc %08x" % (
        zlib.crc32(ha.tobytes()), zlib.crc32(hb.tobytes()), zlib.crc32(heads.tobytes())))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    st_fixtures()
    oracle_pins(images())
