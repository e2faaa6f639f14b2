mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2g/pytest.log
tail -4 gpurun_out/r2g/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err; echo "bench rc $?"; tail -3 gpurun_out/r2g/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2g/bench.json"))
print("value", d["value"], d["ms_per_step"], d["value_spread"]); print("e2e", d["e2e"]["value"], d["e2e"]["spread"], d["e2e"]["blocking"]["value"])
print("clocks", d["clocks"]); print("roofline", d["roofline"]); print("issue", d["roofline_issue"]); print("cpu", d["cpu_baseline"])
print("config4", d["config4"]); print("config5", d["config5"]); print("config3", d["config3"]); print("extra", d["extra"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:yoloface_fused -s 2 -c 1 -o gpurun_out/r2g/fused_v9_b256 -f python tools/run_once.py 256 fused 3 > gpurun_out/r2g/ncu256.log 2>&1
