#!/usr/bin/env python
"""Per-layer roofline table of the layer-by-layer kernels (north_star: tensor-pipe / HBM fractions per layer).
usage: layer_roofline.py [batch] [reps] [out_prefix]
Times every fused step with CUDA events (library step profiling: one event pair + sync per step) on a
batch large enough to be throughput-bound, and relates the ALGORITHMIC bytes of the step (unpadded tensor
in + out) to the measured HBM copy peak in MEASURED_PEAKS.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prefix = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "layers")
HW = int(sys.argv[4]) if len(sys.argv) > 4 else 56
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); src = "measured"
except Exception:  # noqa: BLE001
    peak, src = 6650.0, "fallback"
yf = pkg.load()
net = yf.Network(chunk_images=B, mode="layered")
if HW != 56:
    net.set_input_size(HW, HW)
x = torch.randint(-128, 128, (B, HW, HW, 3), dtype=torch.int8, device="cuda")
y = torch.empty((B, HW // 8, HW // 8, 18), dtype=torch.int8, device="cuda")
for _ in range(3):
    net.run(x, y, n=B)
net.set_step_profiling(True)
acc = None
for _ in range(reps):
    net.run(x, y, n=B)
    cur = [s["last_ms"] for s in net.steps()]
    acc = cur if acc is None else [a + c for a, c in zip(acc, cur)]
net.set_step_profiling(False)
steps = net.steps()
rows = []
for s, a in zip(steps, acc):
    ms = a / reps
    nbytes = (s["bytes_read"] + s["bytes_written"]) * B
    rows.append({"step": s["name"], "kind": s["kind"], "tflite_ops": s["n_ops"], "ms": ms, "alg_bytes": nbytes, "GBps": nbytes / ms / 1e6,
                 "hbm_frac": nbytes / ms / 1e6 / peak, "macs": s["macs"] * B, "int8_TOPs": 2 * s["macs"] * B / ms / 1e9})
total = sum(r["ms"] for r in rows)
json.dump({"batch": B, "reps": reps, "hbm_peak_GBps": peak, "peak_source": src, "total_ms": total, "images_per_s": B / total * 1e3, "steps": rows},
          open(prefix + ".json", "w"), indent=1)
with open(prefix + ".md", "w") as f:
    f.write("# Layer-by-layer kernels, %dx%d input, batch %d, %d reps (CUDA events per step; HBM peak %.0f GB/s %s)\n\n" % (HW, HW, B, reps, peak, src))
    f.write("sum of steps %.3f ms -> %.2f M img/s\n\n| step | TFLite ops | ms | share | algorithmic MB | GB/s | frac of HBM peak | int8 TOP/s |\n|---|---|---|---|---|---|---|---|\n" % (total, B / total / 1e3))
    for r in rows:
        f.write("| %s | %d | %.4f | %.3f | %.2f | %.0f | %.3f | %s |\n" % (r["step"], r["tflite_ops"], r["ms"], r["ms"] / total, r["alg_bytes"] / 1e6, r["GBps"],
                                                                       r["hbm_frac"], ("%.1f" % r["int8_TOPs"]) if r["macs"] else "-"))
print(open(prefix + ".md").read())
net.close()
