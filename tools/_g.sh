mkdir -p gpurun_out/r3b
for hs in 1 2; do
YF_B200_H2D_STREAMS=$hs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29513 tools/e2e_probe.py > gpurun_out/r3b/probe_hs$hs.json 2> gpurun_out/r3b/probe_hs$hs.err
tail -1 gpurun_out/r3b/probe_hs$hs.json
YF_B200_H2D_STREAMS=$hs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r3b/bench_hs$hs.json 2> gpurun_out/r3b/bench_hs$hs.err
python -c "
import json; d=json.load(open('gpurun_out/r3b/bench_hs$hs.json')); print('hs$hs value', d['value'], 'e2e', d['e2e']['value'], 'sustained', d['e2e']['sustained_aggregate']['value'], 'blocking', d['e2e']['blocking']['value'], d['e2e']['single_image']['device_resident_launch_us'])"
done
