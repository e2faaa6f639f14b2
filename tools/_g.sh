mkdir -p gpurun_out/r2x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2x/pytest.log
tail -3 gpurun_out/r2x/pytest.log
python tools/lat_probe.py 1 2 8 2>&1 | tee gpurun_out/r2x/lat_probe.log
