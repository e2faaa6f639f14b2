mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2d/pytest.log
tail -5 gpurun_out/r2d/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err
YF_B200_LIB=stm32h7-yolo_b200/libyoloface_b200_trace.so timeout 120 python tools/fused_trace.py 8192 > gpurun_out/r2d/trace8192.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2d/bench.json")); print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("serial",{}).get("value"), d.get("extra",{}).get("device_resident_images_per_s"), d["e2e"]["single_image_call_us"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yoloface_fused -s 2 -c 1 -o gpurun_out/r2d/fused_v9_b8192 -f python tools/run_once.py 8192 fused 3 > gpurun_out/r2d/ncu8192.log 2>&1
