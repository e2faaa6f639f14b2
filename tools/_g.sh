mkdir -p gpurun_out/r2j
YF_B200_GRAPH=0 timeout 300 python tools/layer_roofline.py 512 5 gpurun_out/r2j/layers_224_b512 224 > gpurun_out/r2j/layers224.log 2>&1
YF_B200_GRAPH=0 timeout 300 python tools/layer_roofline.py 8192 5 gpurun_out/r2j/layers_56_b8192 56 > gpurun_out/r2j/layers56.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2j/layered_224_b256_ncu.csv python tools/run_once.py 256 layered 1 224 > gpurun_out/r2j/ncu224.log 2>&1
python tools/h2d_ceiling.py > gpurun_out/r2j/h2d_n1.json 2>&1
cat gpurun_out/r2j/layers_224_b512.md | head -40; cat gpurun_out/r2j/h2d_n1.json
