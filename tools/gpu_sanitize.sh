#!/bin/bash
# compute-sanitizer memcheck over the kernels added this session (small cases: the tool slows kernels 10-50x)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_capture.py tests/test_gpu_loudness.py -m gpu -x -q \
  -k "golden_capture or unaligned_and_vector_paths_agree and s24le or ring_matches_reference or mic_tick_errors or (one_shot_time_chunked and 1.05) or (one_shot_time_chunked and 12.3) or (rows_any_even_split and 296) or (rows_any_even_split and 700)" \
  > gpurun_out/sanitize.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitize.log
grep -c "Invalid\|Misaligned\|out of bounds" gpurun_out/sanitize.log; tail -n 15 gpurun_out/sanitize.log
