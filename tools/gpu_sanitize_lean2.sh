#!/bin/bash
# second pass: k_results_lean as its own kernel (six channels after k_loudness_rows_any; feeds off the 100 ms grid) and the
# reset / lazy-gating-flush sequence, under memcheck
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_results_lean.py -m gpu -x -q -k "130-6-96000 or 257-2-44100" > gpurun_out/sanitize_lean2_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize_lean2_memcheck.log
tail -n 6 gpurun_out/sanitize_lean2_memcheck.log
