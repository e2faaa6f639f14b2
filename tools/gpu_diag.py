#!/usr/bin/env python
"""First-contact GPU diagnostic: run the layer-by-layer path in observer mode and report, per
TFLite tensor, whether it matches the CPU oracle (mismatch counts + first differing element).
Usage (on a B200 box): python tools/gpu_diag.py [n_images]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402
from oracle_lib import Oracle, vector_a, vector_b  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    yf = pkg.load()
    o = Oracle()
    imgs = np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))
    batch = np.stack([vector_a(), vector_b()] + [imgs[i % 27] for i in range(max(0, n - 2))])[:n]
    t0 = time.time()
    net = yf.Network(observer=True, chunk_images=256)
    print("create+init %.2fs stats=%s" % (time.time() - t0, net.stats()))
    try:
        heads = net.ai_run(batch)
    except Exception as e:  # noqa: BLE001
        print("RUN FAILED:", e)
        return 2
    ref = [o.run(batch[i], dump=True) for i in range(n)]
    bad_total = 0
    P = yf.plan(56, 56)
    t2step = {}
    for s in P["steps"]:
        for op in s["ops"]:
            t2step[o.op(op)["output"]] = s["name"]
    for op in range(o.num_ops):
        info = o.op(op)
        t = info["output"]
        got = net.get_tensor(t, n)
        if got is None:
            print("op %2d tensor %3d: folded away" % (op, t))
            continue
        exp = np.stack([ref[i][1][op] for i in range(n)])
        if got.shape != exp.shape:
            print("op %2d tensor %3d: SHAPE %s vs %s" % (op, t, got.shape, exp.shape)); bad_total += 1; continue
        bad = np.argwhere(got != exp)
        if len(bad) == 0:
            print("op %2d tensor %3d %-14s: OK %s" % (op, t, t2step.get(t, ""), got.shape))
        else:
            bad_total += 1
            b = tuple(bad[0])
            print("op %2d tensor %3d %-14s: MISMATCH %d/%d first@%s got %d exp %d; imgs-bad=%s chans-bad=%s" % (
                op, t, t2step.get(t, ""), len(bad), got.size, b, got[b], exp[b],
                sorted(set(bad[:, 0]))[:8], sorted(set(bad[:, 3]))[:24]))
    exp_heads = np.stack([ref[i][0] for i in range(n)])
    print("heads (observer mode) equal:", np.array_equal(heads, exp_heads))
    net.set_observer(False)
    print("stats after observer off:", net.stats())
    h2 = net.run(batch)
    eq = np.array_equal(h2, exp_heads)
    print("heads (auto mode: fused=%d) equal: %s" % (net.stats()["fused"], eq))
    if not eq:
        bad_total += 1
        bad = np.argwhere(h2 != exp_heads)
        print("  fused mismatches: %d/%d first %s; images %s; cells %s; channels %s" % (
            len(bad), h2.size, bad[:5].tolist(), sorted(set(bad[:, 0])), sorted(set(map(tuple, bad[:, 1:3].tolist())))[:10], sorted(set(bad[:, 3]))))
    lay = yf.Network(chunk_images=256, mode="layered")
    h3 = lay.run(batch)
    print("heads (layered fast mode) equal:", np.array_equal(h3, exp_heads))
    if not np.array_equal(h3, exp_heads):
        bad_total += 1
    import time as _t
    for nimg in (256, 4096, 32768):
        big = np.concatenate([batch] * (nimg // n + 1))[:nimg]
        try:
            import torch
            dbig = torch.from_numpy(big).cuda(); dout = torch.empty((nimg, 7, 7, 18), dtype=torch.int8, device="cuda")
            for who, nn in (("fused", net), ("layered", lay)):
                nn.enqueue(dbig, dout, nimg); nn.sync()
                t0 = _t.perf_counter()
                reps = 20 if nimg <= 4096 else 5
                for _ in range(reps):
                    nn.enqueue(dbig, dout, nimg)
                nn.sync()
                dt = (_t.perf_counter() - t0) / reps
                okk = np.array_equal(dout.cpu().numpy()[:n], exp_heads)
                print("  %-8s n=%6d  %.3f ms  %.2f Mimg/s  ok=%s" % (who, nimg, dt * 1e3, nimg / dt / 1e6, okk))
        except Exception as e:  # noqa: BLE001
            print("  timing failed:", e)
    lay.close()
    # timing per step
    net.set_step_profiling(True)
    big = np.concatenate([batch] * (256 // n + 1))[:256]
    net.run(big); net.run(big)
    for s in net.steps():
        print("  step %-16s kind %d ops %2d  %.3f ms" % (s["name"], s["kind"], s["n_ops"], s["last_ms"]))
    net.set_step_profiling(False)
    print("stats:", net.stats())
    print("DIAG RESULT:", "ALL OK" if bad_total == 0 else "%d tensors differ" % bad_total)
    net.close()
    return 0 if bad_total == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
