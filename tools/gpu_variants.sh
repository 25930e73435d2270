#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/variants.log
echo "== default" >> gpurun_out/variants.log
FORCES=5 timeout 200 python tools/time_wtile.py >> gpurun_out/variants.log 2>&1
for f in soundscope_b200/_variants/lib_*.so; do
  echo "== $f" >> gpurun_out/variants.log
  SSB_LIB=$PWD/$f FORCES=5 timeout 200 python tools/time_wtile.py >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log
