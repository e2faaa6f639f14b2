mkdir -p gpurun_out/r3c
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3c/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3c/pytest.log
tail -3 gpurun_out/r3c/pytest.log
YF_B200_LIB=stm32h7-yolo_b200/libyoloface_b200_trace.so timeout 120 python tools/fused_trace.py 1 > gpurun_out/r3c/trace1_lat.log 2>&1
grep "total" gpurun_out/r3c/trace1_lat.log
python tools/lat_probe.py 1 8 2>&1 | tee gpurun_out/r3c/lat_probe.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r3c/bench$i.json 2> gpurun_out/r3c/bench$i.err
python -c "
import json; d=json.load(open('gpurun_out/r3c/bench$i.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'sustained', d['e2e']['sustained_aggregate']['value'], 'serial', d['serial']['value'], d['e2e']['single_image']['blocking_call_us']['median'], d['e2e']['single_image']['device_resident_launch_us'])"
done
