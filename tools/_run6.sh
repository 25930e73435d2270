mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_results_lean.py tests/test_gpu_wtile.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r6_pytest.log
echo "pytest rc=$?" >> gpurun_out/r6_pytest.log
FORCES=5 timeout 300 python tools/time_wtile.py > gpurun_out/r6_time.log 2>&1
cat gpurun_out/r6_pytest.log gpurun_out/r6_time.log
