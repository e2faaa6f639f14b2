mkdir -p gpurun_out/r2p
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_second_model.py -m gpu -x -q > gpurun_out/r2p/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2p/pytest.log
tail -3 gpurun_out/r2p/pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2p/bench$i.json 2> gpurun_out/r2p/bench$i.err
python -c "
import json; d=json.load(open('gpurun_out/r2p/bench$i.json')); print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d['serial']['value'], 'launch', d['roofline']['launch_ms'], '1img', d['e2e']['single_image_call_us'])"
done
python - <<'PY'
import sys, time, torch
sys.path.insert(0, "tests")
import pkg
yf = pkg.load()
net = yf.Network(chunk_images=8192)
x = torch.randint(-128, 128, (8192, 56, 56, 3), dtype=torch.int8, device="cuda"); y = torch.empty((8192, 7, 7, 18), dtype=torch.int8, device="cuda")
st = torch.cuda.Stream(); net.set_stream(st.cuda_stream)
for _ in range(3): net.enqueue(x, y, 8192)
net.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(20): net.enqueue(x, y, 8192)
e1.record(st); e1.synchronize()
print("8192/launch: %.3f M img/s" % (8192 * 20 / e0.elapsed_time(e1) / 1e3))
PY
