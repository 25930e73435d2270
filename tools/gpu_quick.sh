#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wtile.py -m gpu -q -x > gpurun_out/wt_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/wt_pytest.log
FORCES=5 timeout 300 python tools/time_wtile.py > gpurun_out/wt_time.log 2>&1
FORCE=5 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_loudness_wtile -s 3 -c 1 \
  -o gpurun_out/prof_wtile_fused -f python tools/prof_cfg2.py > gpurun_out/prof_wtile_fused.log 2>&1
tail -3 gpurun_out/wt_pytest.log; cat gpurun_out/wt_time.log
