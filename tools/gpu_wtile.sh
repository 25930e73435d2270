#!/bin/bash
# k_loudness_wtile on the GPU: parity tests, A/B timing (+ diagnostic variants), ncu capture (filter only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wtile.py tests/test_gpu_ebu.py -m gpu -q -x > gpurun_out/wt_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/wt_pytest.log
FORCES=3,5,6 timeout 300 python tools/time_wtile.py > gpurun_out/wt_time.log 2>&1
for f in soundscope_b200/_variants/lib_*.so; do
  echo "== $f" >> gpurun_out/wt_time.log
  SSB_LIB=$PWD/$f FORCES=5 timeout 200 python tools/time_wtile.py >> gpurun_out/wt_time.log 2>&1
done
NOFUSE=1 FORCE=5 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_loudness_wtile -s 3 -c 1 \
  -o gpurun_out/prof_wtile -f python tools/prof_cfg2.py > gpurun_out/prof_wtile.log 2>&1
tail -n 6 gpurun_out/wt_pytest.log; cat gpurun_out/wt_time.log
