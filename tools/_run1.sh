mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_results_lean.py tests/test_gpu_wtile.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/lean_pytest.log
echo "pytest rc=$?" >> gpurun_out/lean_pytest.log
FORCES=5 timeout 300 python tools/time_wtile.py > gpurun_out/lean_time.log 2>&1
SSB_RESULTS_LEAN=0 FORCES=5 timeout 300 python tools/time_wtile.py >> gpurun_out/lean_time.log 2>&1
cat gpurun_out/lean_pytest.log gpurun_out/lean_time.log
