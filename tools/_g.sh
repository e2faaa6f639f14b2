mkdir -p gpurun_out/r2q
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yoloface_fused -s 2 -c 1 -o gpurun_out/r2q/fused_v10_b8192 -f python tools/run_once.py 8192 fused 3 > gpurun_out/r2q/ncu8192.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yoloface_fused -s 2 -c 1 -o gpurun_out/r2q/fused_v10_b256 -f python tools/run_once.py 256 fused 3 > gpurun_out/r2q/ncu256.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r2q/ncu_bench.log 2>&1
YF_B200_LIB=stm32h7-yolo_b200/libyoloface_b200_trace.so timeout 120 python tools/fused_trace.py 8192 > gpurun_out/r2q/trace8192.log 2>&1
YF_B200_LIB=stm32h7-yolo_b200/libyoloface_b200_trace.so timeout 120 python tools/fused_trace.py 256 > gpurun_out/r2q/trace256.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q/smoke.log 2>&1; cat gpurun_out/r2q/smoke.log | tail -2
ls -la gpurun_out/r2q
