mkdir -p gpurun_out/r3j
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > gpurun_out/r3j/bench_n8.json 2> gpurun_out/r3j/bench_n8.err; echo "bench8 rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r3j/bench_n8.json')); print('N=8 value', d['value'], 'e2e', d['e2e']['value'], 'sustained', d['e2e']['sustained_aggregate']['value'], d['e2e']['bounds']['box_fed_images_per_s'], d.get('config3',{}).get('images_per_s'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r3j/bench_n4.json 2> gpurun_out/r3j/bench_n4.err; echo "bench4 rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r3j/bench_n4.json')); print('N=4 value', d['value'], 'e2e', d['e2e']['value'], 'sustained', d['e2e']['sustained_aggregate']['value'], d['e2e']['bounds']['box_fed_images_per_s'])"
timeout 300 python tools/one_handle_probe.py > gpurun_out/r3j/one_handle.log 2>&1; cat gpurun_out/r3j/one_handle.log
timeout 300 python -m pytest tests/test_multi_gpu_handle.py -m gpu -x -q 2>&1 | tail -2
