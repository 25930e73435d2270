#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/time_oneshot.py > gpurun_out/oneshot.log 2>&1
timeout 600 python -m pytest tests/test_gpu_loudness.py tests/test_gpu_spectrum.py -m gpu -x -q -k "one_shot or calculate_integrated or preanalyze or reference_tests" > gpurun_out/pytest_oneshot.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_oneshot.log
cat gpurun_out/oneshot.log; tail -n 12 gpurun_out/pytest_oneshot.log
