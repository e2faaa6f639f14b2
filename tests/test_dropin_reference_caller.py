"""The reference's own caller, unmodified: stm32/X-CUBE-AI/App/yoloface.c compiled from where it lies
(oracle/Makefile target `ref` -> oracle/_ref/libyoloface_ref.so, against ST's headers and stub board
headers) and linked at load time against libyoloface_b200.so.  Proves the drop-in boundary
 and pins the oracle's pre-processing / decode against the reference's C."""
import ctypes as C
import os

import numpy as np
import pytest

import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libyoloface_ref.so")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def ref():
    yf = pkg.load()
    yf.build()
    C.CDLL(yf.LIB_PATH, mode=C.RTLD_GLOBAL)       # provides ai_network_* to the reference object
    r = C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
    r.aiInit.restype = C.c_int
    r.aiRun.restype = C.c_int
    r.yf_ref_rects.argtypes = [C.POINTER(C.c_int), C.c_int]
    return r


def firmware_rects(dets):
    """yoloface.c:139-147: x/y swapped for the rotated LCD, clamped to 0..55, doubled."""
    out = []
    for x1, y1, x2, y2, _ in dets:
        a, b, c, d = int(y1), int(x2), int(y2), int(x1)       # x1=y-h/2, y1=x+w/2, x2=y+h/2, y2=x-w/2
        a = max(a, 0); b = max(b, 0); c = min(c, 55); d = min(d, 55)
        out.append(tuple((v * 2) & 0xFFFF for v in (a, b, c, d)))
    return sorted(out)


def test_reference_preprocessing_matches_oracle(ref, oracle):
    rng = np.random.default_rng(21)
    for _ in range(4):
        frame = rng.integers(0, 256, 112 * 112 * 2, dtype=np.uint8)
        C.memmove(C.addressof((C.c_uint8 * frame.size).in_dll(ref, "RGB_DATA")), frame.ctypes.data, frame.size)
        ref.resize_rgb565_uint8_112_to_56_direct()
        ref.prepare_yolo_data()
        got = np.frombuffer((C.c_int8 * 9408).in_dll(ref, "in_data"), dtype=np.int8).reshape(56, 56, 3)
        assert np.array_equal(got, oracle.rgb565_to_input(frame))


def test_reference_post_process_matches_oracle_decode(ref, oracle, golden):
    heads = golden["heads_images"]
    total = 0
    for h in heads:
        C.memmove(C.addressof((C.c_int8 * 882).in_dll(ref, "out_data")), np.ascontiguousarray(h).ctypes.data, 882)
        ref.yf_ref_reset()
        ref.post_process()
        buf = (C.c_int * (4 * 512))()
        n = ref.yf_ref_rects(buf, 512)
        got = sorted(tuple(v & 0xFFFF for v in buf[4 * i:4 * i + 4]) for i in range(n))
        dets = oracle.decode_nms(h, 0.7, -1.0)              # the firmware thresholds only (yoloface.c:123)
        assert got == firmware_rects(dets)
        total += n
    assert total >= 30


@pytest.mark.gpu
def test_reference_aiinit_airun_on_b200(ref, oracle, golden):
    """aiInit() / aiRun() exactly as the firmware calls them (yoloface.c:188-240), executing on the GPU."""
    assert ref.aiInit() == 0
    for img in golden["images"][:8]:
        C.memmove(C.addressof((C.c_int8 * 9408).in_dll(ref, "in_data")), np.ascontiguousarray(img).ctypes.data, 9408)
        assert ref.aiRun() == 0
        out = np.frombuffer((C.c_int8 * 882).in_dll(ref, "out_data"), dtype=np.int8).reshape(7, 7, 18)
        assert np.array_equal(out, oracle.run(img))
