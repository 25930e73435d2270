#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wtile.py -m gpu -q -x > gpurun_out/wt_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/wt_pytest.log
FORCES=5 timeout 300 python tools/time_wtile.py > gpurun_out/wt_time.log 2>&1
tail -3 gpurun_out/wt_pytest.log; cat gpurun_out/wt_time.log
