#!/bin/bash
# compute-sanitizer over the lean results path (fused per-pair epilogue of k_loudness_wtile with its named barriers and
# shared-memory staging, lra_scan_fast): memcheck, then racecheck, on the 50-stream case of tests/test_gpu_results_lean.py
mkdir -p gpurun_out
K='test_lean_results_match_full_scan_and_oracle and 50-2-48000'
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_results_lean.py -m gpu -x -q -k "$K" > gpurun_out/sanitize_lean_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize_lean_memcheck.log
timeout 110 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_results_lean.py -m gpu -x -q -k "$K" > gpurun_out/sanitize_lean_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitize_lean_racecheck.log
tail -n 6 gpurun_out/sanitize_lean_memcheck.log; tail -n 12 gpurun_out/sanitize_lean_racecheck.log
