mkdir -p gpurun_out/r3f
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shapes or ragged or single_image or golden or many or resol" > gpurun_out/r3f/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3f/pytest.log
tail -3 gpurun_out/r3f/pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r3f/bench$i.json 2> gpurun_out/r3f/bench$i.err
python -c "
import json; d=json.load(open('gpurun_out/r3f/bench$i.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'sustained', d['e2e']['sustained_aggregate']['value'], 'serial', d['serial']['value'], d['e2e']['single_image']['blocking_call_us']['median'], d['e2e']['single_image']['device_resident_launch_us'])"
done
