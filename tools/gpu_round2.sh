#!/bin/bash
# variant timing: serial rows kernel F=64 x 1 CTA/SM (base) vs F=32 x 2 CTAs/SM (B); rows_any even split
mkdir -p gpurun_out
B=$PWD/soundscope_b200/libssb_B.so
{
echo "== base rows 32768"; timeout 200 env N_STREAMS=32768 python tools/time_cfg2.py
echo "== B rows 32768"; timeout 200 env SSB_LIB=$B N_STREAMS=32768 python tools/time_cfg2.py
echo "== base rows 125000"; timeout 300 env N_STREAMS=125000 python tools/time_cfg2.py
echo "== B rows 125000"; timeout 300 env SSB_LIB=$B N_STREAMS=125000 python tools/time_cfg2.py
echo "== B rows 32768 all"; timeout 200 env SSB_LIB=$B N_STREAMS=32768 python tools/time_cfg2.py --all
echo "== any 6ch 96k 4096 x 38400"; timeout 200 env CHANNELS=6 RATE=96000 N_STREAMS=4096 FRAMES=38400 python tools/time_cfg2.py
echo "== any 6ch 96k 16384 x 19200"; timeout 300 env CHANNELS=6 RATE=96000 N_STREAMS=16384 FRAMES=19200 python tools/time_cfg2.py
echo "== any 6ch 96k 16384 x 19200 all"; timeout 300 env CHANNELS=6 RATE=96000 N_STREAMS=16384 FRAMES=19200 python tools/time_cfg2.py --all
} > gpurun_out/variants.log 2>&1
timeout 600 python -m pytest tests/test_gpu_loudness.py tests/test_cpp_header.py -m gpu -x -q -k "channels or tile_kernel or rows_any or cpp" > gpurun_out/pytest_loud.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_loud.log
timeout 600 env SSB_LIB=$B python -m pytest tests/test_gpu_loudness.py -m gpu -x -q -k "tile_kernel or cfg2_full" > gpurun_out/pytest_loud_B.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_loud_B.log
grep -v "^$" gpurun_out/variants.log | tail -30; tail -3 gpurun_out/pytest_loud.log gpurun_out/pytest_loud_B.log
