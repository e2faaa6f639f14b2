"""SURVEY.md 8f n3 (UART text protocol for the reference's PC monitor) and create-time error paths that need no GPU."""
import ctypes as C
import os

import numpy as np
import pytest

import pkg


def test_uart_frames_parse_with_the_monitor_rules(oracle, golden):
    pkg.load()
    uart = __import__("stm32h7_yolo_b200.uart", fromlist=["format_frame"])
    total = 0
    for f, head in enumerate(golden["heads_images"]):
        dets = oracle.decode_nms(head, 0.7, 0.4)
        lines = uart.format_frame(f, dets)
        assert lines[0] == "=== Frame %d ===" % f and lines[-1] == "[INFO] Total faces detected: %d" % len(dets)
        frame, faces, count = uart.parse_frame(lines)
        assert frame == f and count == len(dets) == len(faces)
        for face, d in zip(faces, dets):
            exp = [min(max(int(v), 0), 55) * 2 for v in d[:4]]
            assert [face["x1"], face["y1"], face["x2"], face["y2"]] == exp
            assert abs(face["confidence"] - d[4]) <= 0.005
        total += count
    assert total > 20
    # a box hanging over the frame edge is clamped (the monitor's regex only accepts non-negative integers)
    lines = uart.format_frame(7, [(-3.5, 10.2, 70.0, 60.0, 0.91)])
    assert uart.parse_frame(lines)[1][0] == {"id": 1, "x1": 0, "y1": 20, "x2": 110, "y2": 110, "width": 110, "height": 90, "confidence": 0.91}


def test_create_reports_model_errors_without_a_gpu(tmp_path):
    yf = pkg.load()
    L = yf.lib()

    def create(path):
        cfg = yf.Config(yf.YF_B200_CONFIG_MAGIC, -1, 0, 0, path.encode())
        buf = yf.AiBuffer(yf.AI_BUFFER_FORMAT_U8, 1, 1, 1, C.sizeof(yf.Config), C.cast(C.pointer(cfg), C.c_void_p), None)
        h = C.c_void_p()
        e = L.ai_network_create(C.byref(h), C.byref(buf))
        return e.type, e.code, h.value
    assert create(str(tmp_path / "missing.tflite")) == (0x33, 0x17, None)           # CREATE_FAILED / INVALID_PTR
    bad = tmp_path / "bad.tflite"; bad.write_bytes(b"\x00" * 64)
    assert create(str(bad)) == (0x33, 0x19, None)                                     # CREATE_FAILED / INVALID_FORMAT
    trunc = tmp_path / "trunc.tflite"
    trunc.write_bytes(open(os.path.join(os.path.dirname(yf.LIB_PATH), "assets", "yoloface_int8.tflite"), "rb").read()[:3000])
    t, c, h = create(str(trunc))
    assert t == 0x33 and h is None                                                    # never a half-parsed model
    wrong = yf.AiBuffer(yf.AI_BUFFER_FORMAT_U8, 1, 1, 1, 8, C.cast(C.create_string_buffer(b"notmagic"), C.c_void_p), None)
    e = L.ai_network_create(C.byref(C.c_void_p()), C.byref(wrong))
    assert (e.type, e.code) == (0x33, 0x19)


@pytest.mark.gpu
def test_model_from_path_matches_embedded(oracle, golden):
    """the model can be given by path (same nine operator types)."""
    yf = pkg.load()
    path = os.path.join(os.path.dirname(yf.LIB_PATH), "assets", "yoloface_int8.tflite")
    n = yf.Network(chunk_images=32, tflite_path=path)
    try:
        assert np.array_equal(n.ai_run(golden["images"]), golden["heads_images"])
    finally:
        n.close()
