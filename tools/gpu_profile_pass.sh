#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench lines, launch list, ncu captures.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step.csv python tools/prof_step.py > gpurun_out/prof_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_loudness_rows_any|k_pcm_to_f32|k_ring_tick" -c 9 -f -o gpurun_out/prof_extra python tools/prof_extra.py > gpurun_out/prof_extra.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_loudness_tile -s 3 -c 1 -f -o gpurun_out/prof_tile python tools/prof_cfg2.py > gpurun_out/prof_tile.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_n1.json | cut -c1-600
