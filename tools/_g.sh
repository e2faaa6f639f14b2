mkdir -p gpurun_out/r2e
for cfg in "4 -1" "8 -1" "6 -1" "8 0" "4 0" "8 1"; do
  set -- $cfg
  YF_B200_LANES=$1 YF_B200_PAIR=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2e/bench_l$1_p$2.json 2> gpurun_out/r2e/bench_l$1_p$2.err
done
YF_B200_LANES=8 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-extra > gpurun_out/r2e/bench_l8_s200.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e/bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], round(d["value"]/1e6,3), round(d["ms_per_step"]*1e3,1), "e2e", round(d["e2e"]["value"]/1e6,3), "serial", round(d["serial"]["value"]/1e6,3), "launch_ms", round(d["roofline"]["launch_ms"],4), "1img_us", round(d["e2e"]["single_image_call_us"],1))
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
