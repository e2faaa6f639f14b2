#!/usr/bin/env python
"""bench.py -- yoloface int8 images/sec on N B200s (one process per GPU), next to the CPU baseline.

  python bench.py --gpus N --steps K --warmup W            our arm  (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU arithmetic (oracle port)

A "step" is one pass of the hot path over one batch of 256 synthetic 56x56x3 int8 images per GPU
(BASELINE.json configs[1]).  `value` = device-resident throughput (inputs already in HBM, the K steps
handed to the library in one call, CUDA events on the launching stream); `e2e` = the same batches
through the pipelined host API with pinned HOST buffers (H2D + kernels + D2H of the raw heads inside
the timed region).  The K-step region is only ~1 ms long, so it is REPEATED (each repeat bracketed by
its own events and a synchronise) until >= 0.5 s of device time has been measured; the reported
figures are the median repeat, with the spread next to them.  No collective is involved: the batch
is sharded by image, ranks only meet at barriers.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

NCU_B256 = "r02_fused_v11_b256_ncu_summary.txt"
NCU_B8192 = "r02_fused_v11_b8192_ncu_summary.txt"
BATCH = 256
RING = 64                       # distinct input batches: 64 x 2.4 MB = 154 MB > 126 MB of L2
IN_BYTES, OUT_BYTES = 56 * 56 * 3, 7 * 7 * 18
METRIC = "yoloface int8 images/sec (bit-exact vs TFLite reference kernels)"
CONFIG = {
    "workload": "yoloface int8 batch 256 per GPU, 56x56x3 in -> 7x7x18 raw head out (BASELINE configs[1])",
    "batch_per_gpu": BATCH,
    "input": "synthetic uniform int8 [256,56,56,3]",
    "l2": "inputs larger than L2: %d distinct input batches (%.0f MB) used round-robin" % (RING, RING * BATCH * IN_BYTES / 1e6),
    "sharding": "by image, one process per GPU, no collective",
}


def env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        for k, bit in names.items():
            if r & bit:
                self.reasons.add(k)

    def run(self):
        if not self.nv:
            return
        while not self.stop_flag:
            try:
                self.sample()
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.002)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def ncu_summary_numbers(name):
    """dram bytes / executed warp instructions of the committed `ncu --set full` summary profiles/<name> (written by
    tools/ncu_summary.py from the .ncu-rep of the same kernel): the bench cites the file instead of a literal."""
    out = {}
    try:
        for line in open(os.path.join(ROOT, "profiles", name)):
            t = line.split()
            if len(t) >= 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(t[1], 1)
                out[t[0]] = float(t[-1]) * mult
            elif len(t) >= 3 and t[0] in ("smsp__inst_executed.sum", "launch__grid_size", "gpu__time_duration.sum"):
                out[t[0]] = float(t[-1])
    except OSError:
        return None
    return out or None


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def tflite_baseline_sample(seconds=10.0):
    """The reference's own interpreter (tflite_prediction.py:23-41) when a TFLite runtime is importable on this box:
    default resolver, num_threads = all cores, batch via resize_tensor_input.  None when there is none (this pool)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import numpy as np
        import dump_tflite_reference as dt
        make, ver = dt.find_interpreter()
        if make is None:
            return None
        cores = os.cpu_count() or 1
        it = make(dt.MODEL, False)
        try:
            it = type(it)(model_path=dt.MODEL, num_threads=cores)
        except Exception:  # noqa: BLE001
            pass
        idx = it.get_input_details()[0]["index"]
        it.resize_tensor_input(idx, [BATCH, 56, 56, 3]); it.allocate_tensors()
        x = np.random.default_rng(0).integers(-128, 128, (BATCH, 56, 56, 3), dtype=np.int8)
        it.set_tensor(idx, x); it.invoke()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            it.set_tensor(idx, x); it.invoke(); n += BATCH
        dt_s = time.perf_counter() - t0
        return {"value": n / dt_s, "unit": "images/s", "cores": cores, "kind": "tflite",
                "sample": "%d images in batches of %d through %s (default resolver, %d threads), %.1f s" % (n, BATCH, ver, cores, dt_s)}
    except Exception as e:  # noqa: BLE001
        print("bench: TFLite baseline unavailable (%s)" % e, file=sys.stderr)
        return None


def cpu_baseline_sample(seconds=10.0):
    """The TFLite interpreter when this box has one, else the oracle port of the reference's CPU arithmetic on all
    host threads; bounded sample."""
    t = tflite_baseline_sample(seconds)
    if t is not None:
        return t
    import numpy as np
    from oracle_lib import Oracle
    o = Oracle()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    probe = rng.integers(-128, 128, (max(64, 2 * cores), 56, 56, 3), dtype=np.int8)
    t = time.perf_counter(); o.run_batch(probe, threads=cores); dt = time.perf_counter() - t
    rate = len(probe) / dt
    n = int(min(32768, max(BATCH, rate * seconds)))
    n -= n % BATCH
    x = rng.integers(-128, 128, (n, 56, 56, 3), dtype=np.int8)
    t = time.perf_counter(); o.run_batch(x, threads=cores); dt = time.perf_counter() - t
    return {"value": n / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d images (%d batches of %d) once through the C oracle (TFLite reference-kernel restatement), %d threads, %.1f s"
                      % (n, n // BATCH, BATCH, cores, dt)}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path.  Its arithmetic lives in TensorFlow-Lite
    (absent here and not installable) and in ST's closed Cortex-M7 library, so this arm times the
    oracle port on the box's host cores (DESIGN.md 'Reference arm')."""
    if rank != 0:
        return
    import numpy as np
    from oracle_lib import Oracle
    o = Oracle()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    # Steps go to the oracle in groups of up to 32 batches per call (its thread pool is per call): one 256-image call per
    # step would charge the CPU arm a thread start-up and a 16-images-per-thread tail per step (5.0 k instead of 7.7 k img/s)
    group = 32
    big = rng.integers(-128, 128, (group * BATCH, 56, 56, 3), dtype=np.int8)
    for k in range(0, args.warmup, group):
        o.run_batch(big[:min(group, args.warmup - k) * BATCH], threads=cores)
    t = time.perf_counter()
    for k in range(0, args.steps, group):
        o.run_batch(big[:min(group, args.steps - k) * BATCH], threads=cores)
    dt = time.perf_counter() - t
    value = BATCH * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "K steps of one 256-image batch each through the C oracle on %d host threads, up to 32 steps per call" % cores},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import pkg
    yf = pkg.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    try:                                    # run this rank (and first-touch its pinned buffers) on the GPU's NUMA node
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception as e:  # noqa: BLE001
        print("bench: no NVML CPU affinity (%s)" % e, file=sys.stderr)
    dist, host_group = None, None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")      # detections travel host to host (no GPU collective on the data path)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    net = yf.Network(device=local_rank, chunk_images=BATCH)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1234 + rank)
    d_in = [torch.randint(-128, 128, (BATCH, 56, 56, 3), dtype=torch.int8, device="cuda", generator=gen) for _ in range(RING)]
    d_out = [torch.empty((BATCH, 7, 7, 18), dtype=torch.int8, device="cuda") for _ in range(RING)]
    stream = torch.cuda.Stream()                 # an explicit stream: torch events must see OUR launches
    assert stream.cuda_stream != 0
    torch.cuda.synchronize()
    net.set_stream(stream.cuda_stream)

    # ---------------- device-resident throughput (`value`) ----------------
    # The K steps are K independent batches: they go to the stream in ONE yf_b200_enqueue_batches call, which
    # spreads them over the library's four kernel lanes (one 256-image launch fills 256 of the 444 CTA slots, so
    # the head of step k+1 runs in the slots step k leaves idle).  `serial` below is the same K steps queued one
    # yf_b200_enqueue at a time on a single stream (no overlap between steps).
    def step_lists(k0, k):
        idx = [(k0 + i) % RING for i in range(k)]
        return [d_in[i] for i in idx], [d_out[i] for i in idx], [BATCH] * k

    def agree_min(v):                            # the same repeat count on every rank
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    def spread(ms_list):
        q = sorted(ms_list)
        return {"repeats": len(q), "median_ms": statistics.median(q), "p10_ms": q[len(q) // 10], "p90_ms": q[(9 * len(q)) // 10],
                "min_ms": q[0], "max_ms": q[-1]}

    for k in range(max(args.warmup, 3)):
        net.enqueue(d_in[k % RING], d_out[k % RING], BATCH)
    net.enqueue_batches(*step_lists(0, max(args.warmup, 3)))
    net.sync()

    def region_serial(k0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(args.steps):
            net.enqueue(d_in[(k0 + k) % RING], d_out[(k0 + k) % RING], BATCH)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def region_value(k0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        net.enqueue_batches(*step_lists(k0, args.steps))
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # EXACTLY K steps per region; the region is repeated until >= 0.5 s of device time (at most 4,000 repeats), every
    # repeat between its own events + synchronise, the whole series between barriers.  The headline is the MEDIAN repeat.
    barrier()
    serial_ms = [region_serial(r * args.steps) for r in range(agree_min(max(5, min(200, int(100.0 / max(region_serial(0), 1e-3))))))]
    first = region_value(0)
    reps = agree_min(max(5, min(4000, int(750.0 / max(first, 1e-3)) + 1)))      # first (cold) region over-estimates the rest
    sampler = ClockSampler(local_rank); sampler.start()
    l0 = net.stats()["kernel_launches"]
    barrier()
    value_ms = [region_value(r * args.steps) for r in range(reps)]
    barrier()
    net.sync()
    clocks = sampler.result()
    launches = (net.stats()["kernel_launches"] - l0) // reps        # per K-step region
    ms = max_over_ranks(statistics.median(value_ms))
    ms_serial = max_over_ranks(statistics.median(serial_ms))
    value = world * BATCH * args.steps / (ms * 1e-3)
    value_spread = spread(value_ms)
    value_spread["device_seconds_measured"] = sum(value_ms) * 1e-3
    serial = {"value": world * BATCH * args.steps / (ms_serial * 1e-3), "ms_per_step": ms_serial / args.steps,
              "api": "yf_b200_enqueue per step on one stream (steps do not overlap)", "repeats": len(serial_ms)}
    fused = bool(net.stats()["fused"])
    # duration of single launches of the dominant kernel (each bracketed by its own events)
    kms = []
    for k in range(60):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream); net.enqueue(d_in[k % RING], d_out[k % RING], BATCH); a1.record(stream)
        a1.synchronize(); kms.append(a0.elapsed_time(a1))
    net.sync()
    kernel_ms = statistics.median(kms[5:])

    # ---------------- end to end through the public API with HOST buffers (`e2e`) ----------------
    net.set_stream(None)
    h_in = [torch.empty((BATCH, 56, 56, 3), dtype=torch.int8).pin_memory() for _ in range(RING)]
    for i, t in enumerate(h_in):
        t.copy_(d_in[i])
    h_out = [torch.empty((BATCH, 7, 7, 18), dtype=torch.int8).pin_memory() for _ in range(8)]
    for k in range(max(args.warmup, 3)):
        net.run(h_in[k % RING], h_out[0], n=BATCH)
    # (a) blocking call per step: H2D + kernel(s) + D2H, returns when the heads are in host memory
    def region_block(k0):
        t0 = time.perf_counter()
        for k in range(args.steps):
            net.run(h_in[(k0 + k) % RING], h_out[0], n=BATCH)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    # (b) the pipelined API: every step still copies its own inputs in and its own heads out, but the copy of
    #     step k+1 overlaps the kernels of step k (copy streams + kernel lanes, ring of six staging slots)
    def region_e2e(k0):
        t0 = time.perf_counter()
        for k in range(args.steps):
            net.submit(h_in[(k0 + k) % RING], h_out[k % 8], BATCH)
        net.wait()
        return time.perf_counter() - t0

    barrier()
    block_s = [region_block(r * args.steps) for r in range(agree_min(max(3, min(50, int(0.1 / max(region_block(0), 1e-6))))))]
    for k in range(3):
        net.submit(h_in[k % RING], h_out[k % 8], BATCH)
    net.wait()
    barrier()
    e2e_reps = agree_min(max(5, min(2000, int(0.6 / max(region_e2e(0), 1e-6)) + 1)))
    barrier()
    e2e_s = [region_e2e(r * args.steps) for r in range(e2e_reps)]
    barrier()
    dt = max_over_ranks(statistics.median(e2e_s))
    dt_block = max_over_ranks(statistics.median(block_s))
    e2e = {"value": world * BATCH * args.steps / dt, "unit": "images/s", "h2d_bytes_per_step": BATCH * IN_BYTES,
           "d2h_bytes_per_step": BATCH * OUT_BYTES, "api": "yf_b200_submit(host pinned in, host pinned out) per step + yf_b200_wait",
           "ms_per_step": 1e3 * dt / args.steps, "spread": spread([1e3 * v for v in e2e_s]),
           "blocking": {"value": world * BATCH * args.steps / dt_block, "ms_per_step": 1e3 * dt_block / args.steps,
                        "api": "yf_b200_run(host pinned in, host pinned out), one blocking call per step", "repeats": len(block_s)}}
    # Sustained aggregate: every rank keeps submitting for ~1 s (one wait at the end) and the per-rank rates are SUMMED.
    # The headline above multiplies the SLOWEST rank's K-step time by the world size (the contract's max over ranks); on a
    # box whose GPUs get unequal shares of the host's copy bandwidth the two differ (tools/e2e_probe.py).
    barrier()
    t0 = time.perf_counter(); ks = 0
    while time.perf_counter() - t0 < 1.0:
        for _ in range(16):
            net.submit(h_in[ks % RING], h_out[ks % 8], BATCH); ks += 1
    net.wait()
    rate = ks * BATCH / (time.perf_counter() - t0)
    if dist:
        t = torch.tensor([rate], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        rate = float(t.item())
    barrier()
    e2e["sustained_aggregate"] = {"value": rate, "unit": "images/s",
                                  "how": "sum over ranks of images / wall time, every rank submitting 256-image steps continuously for 1 s"}
    # the bounds of SURVEY.md 8(d): what this BOX can feed -- pinned host -> device copies by ALL ranks at the same time
    # (tools/h2d_ceiling.py; the per-GPU figure is rank 0's share of that concurrent run) -- and the HBM I/O floor
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import h2d_ceiling
    h2d_mine, h2d_total = h2d_ceiling.measure(1.0, 64, dist)
    if rank == 0:
        e2e["bounds"] = {"box_h2d_GBps": h2d_total, "rank0_h2d_GBps": h2d_mine, "n_gpus_copying": world,
                         "box_fed_images_per_s": h2d_total * 1e9 / IN_BYTES, "e2e_frac_of_box_bound": e2e["value"] / (h2d_total * 1e9 / IN_BYTES),
                         "sustained_frac_of_box_bound": e2e["sustained_aggregate"]["value"] / (h2d_total * 1e9 / IN_BYTES),
                         "how": "every rank copies 64 MiB pinned blocks to its GPU for 1 s, all ranks concurrently (tools/h2d_ceiling.py)"}
    # latency of the reference's own call pattern: one image per blocking ai_network_run-style call, pageable host buffers
    one_in, one_out = np.ascontiguousarray(h_in[0][:1].numpy()).copy(), np.zeros((1, 7, 7, 18), np.int8)
    for _ in range(20):
        net.run(one_in, one_out, n=1)
    calls = []
    for _ in range(300):
        t0 = time.perf_counter(); net.run(one_in, one_out, n=1); calls.append(1e6 * (time.perf_counter() - t0))
    e2e["single_image_call_us"] = statistics.median(calls)
    # the same image device-resident: one launch between two CUDA events on the bench stream
    lat0 = net.stats()["latency_launches"]
    dev1 = []
    net.set_stream(stream.cuda_stream)
    for _ in range(100):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream); net.enqueue(d_in[0][:1], d_out[0][:1], 1); c1.record(stream); c1.synchronize(); dev1.append(1e3 * c0.elapsed_time(c1))
    net.set_stream(None)
    st1 = net.stats()
    e2e["single_image"] = {"blocking_call_us": {"median": statistics.median(calls), "p10": sorted(calls)[30], "p90": sorted(calls)[270], "calls": 300,
                                                "how": "yf_b200_run(1 image, pageable host in/out): staged in mapped pinned memory, the kernel reads / writes it over PCIe, "
                                                       "the host polls the CTA's completion word"},
                           "device_resident_launch_us": {"median": statistics.median(dev1), "min": min(dev1), "launches": 100},
                           "kernel_shape": ("latency: 512-thread CTAs, one per SM" if st1["latency_launches"] - lat0 >= 100 else "throughput: 256-thread CTAs")}
    # sanity: the e2e result of the last step equals the device-resident result for that input
    last = (ks - 1) % RING                       # the last step the sustained loop above submitted
    net.run(d_in[last], d_out[last], n=BATCH)
    assert torch.equal(h_out[(ks - 1) % 8], d_out[last].cpu()), "host-path and device-path heads differ"

    # ---------------- roofline of the dominant kernel ----------------
    roofline, per_step = None, []
    peak, peak_src = measured_peak_hbm()
    roofline_issue = None
    if rank == 0 and fused:
        alg = BATCH * (IN_BYTES + OUT_BYTES)
        # average duration of a launch OVER THE TIMED REGION: its K launches overlap on the kernel lanes, so the region's
        # CUDA-event time / K is what one launch costs there; the duration of a launch that runs alone is kept beside it
        step_ms = ms / args.steps
        gbps = alg / (step_ms * 1e-3) / 1e9
        prof256, prof8k = ncu_summary_numbers(NCU_B256), ncu_summary_numbers(NCU_B8192)
        traffic = (prof256["dram__bytes_read.sum"] + prof256.get("dram__bytes_write.sum", 0.0)) if prof256 else None
        roofline = {"bound": "hbm", "kernel": "yoloface_fused_spec_kernel", "achieved": gbps, "peak": peak, "unit": "GB/s", "frac": gbps / peak,
                    "io_floor_images_per_s": peak * 1e9 / (IN_BYTES + OUT_BYTES),
                    "traffic": traffic, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the `ncu --set full` capture of this kernel at "
                    "256 images, read from profiles/%s" % NCU_B256,
                    "peak_source": peak_src, "launch_ms": step_ms, "isolated_launch_ms": kernel_ms,
                    "achieved_isolated_launch": alg / (kernel_ms * 1e-3) / 1e9,
                    "note": "the single persistent kernel IS the step: algorithmic bytes per launch = 256 x (9,408 B image in + 882 B head out) "
                            "/ average launch duration over the timed region (median region time / K launches, CUDA events on the launching "
                            "stream); isolated_launch_ms = median CUDA-event duration of one launch running alone; the kernel is "
                            "latency/issue-bound, not HBM-bound (DESIGN.md 'Roofline')"}
        # What does bound it: issue slots.  Achieved = executed warp instructions per image (ncu, steady state, read
        # from the committed summary) x the measured images/s; peak = SMs x 4 schedulers x the SM clock sampled above.
        if prof8k and prof8k.get("smsp__inst_executed.sum"):
            ipi = prof8k["smsp__inst_executed.sum"] / 8192.0
            mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
            peak_issue = net.stats()["sm_count"] * 4 * mhz * 1e6
            roofline_issue = {"bound": "issue", "kernel": "yoloface_fused_spec_kernel", "achieved": ipi * value / world, "peak": peak_issue,
                              "unit": "warp-instr/s", "frac": ipi * value / world / peak_issue, "warp_instr_per_image": ipi,
                              "source": "smsp__inst_executed.sum / 8192 images of profiles/%s x measured images/s; peak = %d SMs x 4 x %.0f MHz"
                                        % (NCU_B8192, net.stats()["sm_count"], mhz)}
    # per-step table of the layer-by-layer kernels (the path per-layer ncu evidence is taken on)
    if rank == 0:
        net.set_step_profiling(True)
        acc = None
        reps = 20
        for k in range(reps + 3):
            net.run(d_in[k % RING], d_out[k % RING], n=BATCH)
            if k >= 3:
                cur = [s["last_ms"] for s in net.steps()]
                acc = cur if acc is None else [a + c for a, c in zip(acc, cur)]
        net.set_step_profiling(False)
        steps = net.steps()
        total = sum(acc)
        for s, a in zip(steps, acc):
            avg_ms = a / reps
            bytes_launch = (s["bytes_read"] + s["bytes_written"]) * BATCH
            per_step.append({"name": s["name"], "ms": round(avg_ms, 5), "share": round(a / total, 4),
                             "alg_bytes": bytes_launch, "GBps": round(bytes_launch / (avg_ms * 1e-3) / 1e9, 2),
                             "macs": s["macs"] * BATCH})
        if roofline is None:
            dom = max(per_step, key=lambda d: d["ms"])
            roofline = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                        "frac": dom["GBps"] / peak, "traffic": None, "peak_source": peak_src,
                        "note": "algorithmic bytes per launch (unpadded in+out of the step x 256 images) / mean CUDA-event "
                                "duration of that kernel over %d launches; step share of the summed per-kernel time %.3f" % (reps, dom["share"])}

    cpu = cpu_baseline_sample() if (rank == 0 and world == 1 and not args.no_cpu) else None
    net.close()
    # ---------------- informational: larger launches and the decode+NMS path (not the headline) ----------------
    extra = None
    if rank == 0 and not args.no_extra:
        big = 8192
        net2 = yf.Network(device=local_rank, chunk_images=big)
        xb = torch.randint(-128, 128, (big, 56, 56, 3), dtype=torch.int8, device="cuda", generator=gen)
        xb[::2] = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))[np.arange(big // 2) % 27]).cuda()
        yb = torch.empty((big, 7, 7, 18), dtype=torch.int8, device="cuda")
        net2.set_stream(stream.cuda_stream)
        for _ in range(3):
            net2.enqueue(xb, yb, big)
        net2.sync()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        for _ in range(20):
            net2.enqueue(xb, yb, big)
        b1.record(stream); b1.synchronize(); net2.sync()
        ms_big = b0.elapsed_time(b1) / 20
        net2.set_stream(None)
        net2.detect(xb, 0.7, 0.4, max_det=8, n=big)
        t0 = time.perf_counter()
        for _ in range(5):
            dets, counts = net2.detect(xb, 0.7, 0.4, max_det=8, n=big)
        dt_det = (time.perf_counter() - t0) / 5
        extra = {"batch": big, "device_resident_images_per_s": big / (ms_big * 1e-3), "ms_per_launch": ms_big,
                 "detect_images_per_s": big / dt_det, "detect_api": "yf_b200_detect(device in) -> decode + NMS on device, detections D2H",
                 "detections_per_image": float(counts.mean())}
        net2.close()
    # ---------------- BASELINE configs[3]: 4x upscaled input (224x224 -> 28x28x18 head), the layer-by-layer kernels ----------------
    config4 = None
    if rank == 0 and not args.no_extra:
        config4 = {"input": "224x224x3 int8 (4x per side; head 28x28x18, 2,352 candidates)", "path": "layer-by-layer kernels (26 launches per batch)",
                   "alg_bytes_per_image": {"io_floor": 224 * 224 * 3 + 28 * 28 * 18, "layer_by_layer": 6697466}}
        net4 = yf.Network(device=local_rank, chunk_images=512)
        net4.set_input_size(224, 224)
        net4.set_stream(stream.cuda_stream)
        for b4, reps4 in ((16, 40), (4096, 3)):
            x4 = torch.randint(-128, 128, (b4, 224, 224, 3), dtype=torch.int8, device="cuda", generator=gen)
            y4 = torch.empty((b4, 28, 28, 18), dtype=torch.int8, device="cuda")
            net4.enqueue(x4, y4, b4); net4.sync()
            t4 = []
            for _ in range(reps4):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(stream); net4.enqueue(x4, y4, b4); c1.record(stream); c1.synchronize(); t4.append(c0.elapsed_time(c1))
            ms4 = statistics.median(t4)
            config4["batch_%d" % b4] = {"images_per_s": b4 / (ms4 * 1e-3), "ms_per_batch": ms4, "repeats": reps4,
                                        "layer_by_layer_GBps": 6697466 * b4 / (ms4 * 1e-3) / 1e9,
                                        "frac_of_hbm_peak_on_layer_by_layer_bytes": 6697466 * b4 / (ms4 * 1e-3) / 1e9 / peak}
            if b4 == 4096:                                     # decode + NMS at the 28x28 head (block-per-image kernel)
                net4.set_stream(None)
                d4, c4 = net4.detect(x4[:512], 0.7, 0.4, max_det=16, n=512)
                t0 = time.perf_counter(); d4, c4 = net4.detect(x4[:512], 0.7, 0.4, max_det=16, n=512); dtd = time.perf_counter() - t0
                config4["detect_512_images_per_s"] = 512 / dtd
                net4.set_stream(stream.cuda_stream)
            del x4, y4
        # dominant kernel at 224x224 (CUDA-event time per step, batch 64) against the measured HBM peak
        net4.set_stream(None)
        net4.set_step_profiling(True)
        x4 = torch.randint(-128, 128, (64, 224, 224, 3), dtype=torch.int8, device="cuda", generator=gen)
        y4 = torch.empty((64, 28, 28, 18), dtype=torch.int8, device="cuda")
        acc4 = None
        for k in range(6):
            net4.run(x4, y4, n=64)
            if k >= 1:
                cur = [st["last_ms"] for st in net4.steps()]
                acc4 = cur if acc4 is None else [a + c for a, c in zip(acc4, cur)]
        net4.set_step_profiling(False)
        st4 = net4.steps()
        rows4 = [{"name": st["name"], "ms": a / 5, "GBps": (st["bytes_read"] + st["bytes_written"]) * 64 / (a / 5 * 1e-3) / 1e9} for st, a in zip(st4, acc4)]
        dom4 = max(rows4, key=lambda r: r["ms"])
        config4["dominant_kernel"] = {"name": dom4["name"], "ms_at_batch_64": dom4["ms"], "algorithmic_GBps": dom4["GBps"], "frac_of_hbm_peak": dom4["GBps"] / peak,
                                      "share_of_step": dom4["ms"] / sum(r["ms"] for r in rows4)}
        prof4 = os.path.join(ROOT, "profiles", "r02_layered_224_ncu_metrics.md")
        config4["dram_traffic_vs_algorithmic"] = ("profiles/r02_layered_224_ncu_metrics.md" if os.path.exists(prof4) else None)
        net4.close()
        del x4, y4
    # ---------------- BASELINE configs[4]: float32 PyTorch model vs the int8 GPU path (detection-level tolerance) ----------------
    config5 = None
    if rank == 0 and not args.no_extra:
        try:
            from oracle_lib import Oracle
            from test_float_reference import compare
            imgs5 = np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))
            rng5 = np.random.default_rng(12)
            crops = np.stack([np.roll(imgs5[i % 27], (int(rng5.integers(-6, 7)), int(rng5.integers(-6, 7))), axis=(0, 1)) for i in range(64)])
            batch5 = np.concatenate([imgs5, crops])
            net5 = yf.Network(device=local_rank, chunk_images=128)
            heads5 = net5.run(batch5)
            net5.close()
            mae, mx, both, only_i, only_f = compare(heads5, Oracle(), batch5)
            config5 = {"images": len(batch5), "logit_mae": mae, "logit_max_abs_diff": mx, "head_lsb": 0.14218327403068542,
                       "detections_in_both": both, "only_int8": only_i, "only_float32": only_f,
                       "reference": "float32 torch restatement of yoloface/pytorch/yoloface.py:67-175 built from the de-quantised int8 weights (tests/float_reference.py)"}
        except Exception as e:  # noqa: BLE001
            config5 = {"error": repr(e)}
    # ---------------- informational: BASELINE configs[2] -- 65,536 images sharded by image range over the ranks,
    # decode + NMS on device, only the packed detections gathered on rank 0 (no data-path collective) ----------------
    config3 = None
    if not args.no_extra:
        from stm32h7_yolo_b200 import sharding
        total = 65536
        lo, hi = sharding.shard_bounds(total, world, rank)
        mine = hi - lo
        net3 = yf.Network(device=local_rank, chunk_images=8192)
        faces = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))).cuda()
        x3 = torch.randint(-128, 128, (mine, 56, 56, 3), dtype=torch.int8, device="cuda", generator=gen)
        x3[::2] = faces[(torch.arange(lo, hi, 2, device="cuda") // 2) % len(faces)]
        dets, counts = net3.detect(x3, 0.7, 0.4, max_det=8, n=mine)      # warm-up of the kernels and of the gather path
        sharding.gather_detections(*sharding.pack_detections(dets[:64], counts[:64]), dist if dist else None, group=host_group)
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        dets, counts = net3.detect(x3, 0.7, 0.4, max_det=8, n=mine)
        flat, cnt = sharding.pack_detections(dets, counts)
        gathered = sharding.gather_detections(flat, cnt, dist if dist else None, group=host_group)
        dt3 = max_over_ranks(time.perf_counter() - t0)
        if rank == 0:
            gflat, gcnt = gathered
            assert len(gcnt) == total and len(gflat) == int(gcnt.sum())
            config3 = {"images": total, "images_per_rank": mine, "images_per_s": total / dt3, "seconds": dt3, "detections": int(gcnt.sum()),
                       "api": "yf_b200_detect(device in) per rank -> decode + NMS on device -> detections D2H -> packed lists gathered on rank 0 "
                              "host to host over a gloo group (no NCCL on the data path)"}
        net3.close()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
                "data": "synthetic", "config": CONFIG, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "roofline_issue": roofline_issue, "cpu_baseline": cpu, "serial": serial, "value_spread": value_spread,
                "value_api": "yf_b200_enqueue_batches(K independent device-resident batches) on the caller's stream", "path": "fused single kernel" if fused else "layer-by-layer kernels",
                "extra": extra, "config3": config3, "config4": config4, "config5": config5, "layered_kernels": per_step}
        emit(line)
    if dist:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banners, warnings
    from libraries) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                         # fd 1 -> stderr for the rest of the run (native libraries print there too)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the informational large-batch / decode leg")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
