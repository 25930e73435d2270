#!/bin/bash
mkdir -p gpurun_out
{
echo "== any 6ch 96k 4096 x 38400"; timeout 200 env CHANNELS=6 RATE=96000 N_STREAMS=4096 FRAMES=38400 python tools/time_cfg2.py
echo "== any 6ch 96k 16384 x 19200"; timeout 300 env CHANNELS=6 RATE=96000 N_STREAMS=16384 FRAMES=19200 python tools/time_cfg2.py
echo "== any 6ch 96k 16384 x 19200 all"; timeout 300 env CHANNELS=6 RATE=96000 N_STREAMS=16384 FRAMES=19200 python tools/time_cfg2.py --all
} > gpurun_out/variants3.log 2>&1
timeout 600 python -m pytest tests/test_gpu_loudness.py tests/test_cpp_header.py -m gpu -x -q -k "channels or tile_kernel or rows_any or cpp" > gpurun_out/pytest_loud3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_loud3.log
cat gpurun_out/variants3.log; tail -n 4 gpurun_out/pytest_loud3.log
