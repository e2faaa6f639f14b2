mkdir -p gpurun_out/r2w
YF_B200_TRACE_INNER=1,2,3,5,6,7,8,14,15,16,17,25 YF_B200_LIB=stm32h7-yolo_b200/libyoloface_b200_trace.so timeout 120 python tools/fused_trace.py 1 > gpurun_out/r2w/trace1_lat.log 2>&1
grep -A1 "total\|phase " gpurun_out/r2w/trace1_lat.log
python tools/lat_probe.py 1 148 2>&1 | tee gpurun_out/r2w/lat_probe.log
