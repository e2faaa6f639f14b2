mkdir -p gpurun_out/r3m
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3m/pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3m/pytest.log
tail -3 gpurun_out/r3m/pytest.log
