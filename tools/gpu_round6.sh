#!/bin/bash
mkdir -p gpurun_out
{ timeout 200 python tools/time_fft.py; timeout 200 env NFFT=16384 NWIN=8192 python tools/time_fft.py; timeout 200 env NFFT=4096 NWIN=32768 python tools/time_fft.py; } > gpurun_out/fft_packed.log 2>&1
timeout 600 python -m pytest tests/test_gpu_spectrum.py -m gpu -x -q > gpurun_out/pytest_fft.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fft.log
cat gpurun_out/fft_packed.log; tail -n 4 gpurun_out/pytest_fft.log
