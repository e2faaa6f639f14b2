mkdir -p gpurun_out/r2k
nvidia-smi topo -m > gpurun_out/r2k/topo.txt 2>&1
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py > gpurun_out/r2k/h2d_n$n.json 2> gpurun_out/r2k/h2d_n$n.err
done
cat gpurun_out/r2k/h2d_n*.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2k/bench_n8.json 2> gpurun_out/r2k/bench_n8.err; echo "bench8 rc $?"
timeout 300 python -m pytest tests/test_multi_gpu_handle.py -m gpu -x -q > gpurun_out/r2k/pytest_mg8.log 2>&1; tail -3 gpurun_out/r2k/pytest_mg8.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2k/bench_n8.json"))
print("N8 value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("bounds"), "config3", d.get("config3"))
PY
# one handle over 8 GPUs, host buffers: throughput of a single blocking call
python - <<'PY' > gpurun_out/r2k/one_handle.log 2>&1
import sys, time, numpy as np, torch
sys.path.insert(0, "tests")
import pkg
yf = pkg.load()
for devs in ([0], [0, 1], [0, 1, 2, 3], list(range(8))):
    net = yf.Network(devices=devs, chunk_images=1024)
    n = 65536
    x = torch.randint(-128, 128, (n, 56, 56, 3), dtype=torch.int8).pin_memory()
    y = torch.empty((n, 7, 7, 18), dtype=torch.int8).pin_memory()
    net.run(x, y, n=n)
    t0 = time.perf_counter()
    for _ in range(3):
        net.run(x, y, n=n)
    dt = (time.perf_counter() - t0) / 3
    print("devices", len(devs), "yf_b200_run(65536 pinned host images): %.2f M img/s (%.1f ms)" % (n / dt / 1e6, dt * 1e3))
    net.close()
PY
cat gpurun_out/r2k/one_handle.log
