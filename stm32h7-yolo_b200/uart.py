"""UART text protocol of the reference firmware, so the reference's PC monitor
(上位机/IAP/main.py:317-369) can consume GPU results unchanged.

Emitted per frame, exactly as stm32/User/main.c:46,53 and stm32/X-CUBE-AI/App/yoloface.c:148 print it:
    === Frame <n> ===
    ----------------------------------------
    [Face <k>] BBox: [<x1>, <y1>, <x2>, <y2>], Conf: <0.xx>
    ----------------------------------------
    [INFO] Total faces detected: <n>
Coordinates are clamped to the frame and doubled like the firmware does (yoloface.c:143-147: model
pixels 0..55 -> LCD pixels 0..110); the firmware's x/y swap for its rotated LCD is NOT replicated."""
import re

SEP = "-" * 40
# the three patterns the monitor applies to every received line (上位机/IAP/main.py:324, 332-333, 361)
RE_FRAME = re.compile(r'=== Frame (\d+) ===')
RE_FACE = re.compile(r'\[Face\s+(\d+)\]\s+BBox:\s*\[(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\],\s*Conf:\s*([\d\.]+)')
RE_TOTAL = re.compile(r'Total faces detected:\s*(\d+)', re.IGNORECASE)


def format_frame(frame, dets, size=56):
    """dets: iterable of (x1, y1, x2, y2, conf) in model pixels -> list of text lines (no line endings)."""
    lines = ["=== Frame %d ===" % frame, SEP]
    n = 0
    for x1, y1, x2, y2, conf in dets:
        n += 1
        c = [min(max(int(v), 0), size - 1) * 2 for v in (x1, y1, x2, y2)]
        lines.append("[Face %d] BBox: [%d, %d, %d, %d], Conf: %.2f" % (n, c[0], c[1], c[2], c[3], conf))
    lines += [SEP, "[INFO] Total faces detected: %d" % n]
    return lines


def parse_frame(lines):
    """What the monitor's parse_frame_data does: -> (frame_num, [face dicts], face_count)."""
    frame, faces, count = 0, [], 0
    for line in lines:
        m = RE_FRAME.search(line)
        if m:
            frame = int(m.group(1))
        m = RE_FACE.search(line)
        if m:
            x1, y1, x2, y2 = (int(m.group(i)) for i in range(2, 6))
            faces.append({"id": int(m.group(1)), "x1": x1, "y1": y1, "x2": x2, "y2": y2, "width": x2 - x1, "height": y2 - y1,
                          "confidence": float(m.group(6))})
        m = RE_TOTAL.search(line)
        if m:
            count = int(m.group(1))
    return frame, faces, count
