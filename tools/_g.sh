mkdir -p gpurun_out/r2o
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2o/pytest.log
tail -4 gpurun_out/r2o/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o/bench.json 2> gpurun_out/r2o/bench.err; echo "bench rc $?"; tail -2 gpurun_out/r2o/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2o/bench.json"))
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["sustained_aggregate"]["value"], d["e2e"]["bounds"]["box_fed_images_per_s"], "serial", d["serial"]["value"], "1img", d["e2e"]["single_image_call_us"])
print("config4", d["config4"]["batch_16"]["images_per_s"], d["config4"]["batch_4096"]["images_per_s"], d["config4"]["dominant_kernel"])
print("extra", d["extra"])
PY
YF_B200_MODE=layered timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2o/bench_layered.json 2> gpurun_out/r2o/bench_layered.err
python -c "
import json; d=json.load(open('gpurun_out/r2o/bench_layered.json')); print('layered value', d['value'], 'serial', d['serial']['value'])"
YF_B200_LIB=stm32h7-yolo_b200/libyoloface_b200_trace.so timeout 120 python tools/fused_trace.py 8192 > gpurun_out/r2o/trace8192.log 2>&1; grep "img 1 front conv3x3\|img 1 front conv1x1_6\|^total" gpurun_out/r2o/trace8192.log
