#!/bin/bash
# ncu --set full of two fused cfg2 launches with the lean results epilogue: launch 9 (no 3 s entry pending) and launch 10 (one)
mkdir -p gpurun_out
N_LAUNCH=12 FORCE=5 timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_loudness_wtile -s 8 -c 2 \
  -o gpurun_out/prof_wtile_lean2 -f python tools/prof_cfg2.py > gpurun_out/prof_wtile_lean2.log 2>&1
tail -3 gpurun_out/prof_wtile_lean2.log
