mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2i/pytest.log
tail -6 gpurun_out/r2i/pytest.log
YF_B200_MODE=layered timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2i/bench_layered.json 2> gpurun_out/r2i/bench_layered.err; echo "rc $?"; tail -2 gpurun_out/r2i/bench_layered.err
YF_B200_MODE=layered YF_B200_GRAPH=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2i/bench_layered_nograph.json 2> gpurun_out/r2i/bench_layered_nograph.err
python - <<'PY'
import json
for f in ("bench_layered","bench_layered_nograph"):
    try:
        d=json.load(open("gpurun_out/r2i/%s.json"%f)); print(f, "value", round(d["value"]/1e6,3), "serial", round(d["serial"]["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), "launch_ms", d["roofline"].get("launch_ms"), d.get("path"))
        if d.get("config4"): print(" config4", d["config4"]["batch_16"], d["config4"]["batch_4096"])
        if d.get("extra"): print(" extra", d["extra"])
    except Exception as e: print(f, "ERR", e)
PY
