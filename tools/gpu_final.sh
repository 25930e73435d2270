#!/bin/bash
# end-of-session verification: full parity suite, smoke, bench lines (N=1 with extras, reference arm), launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step.csv python tools/prof_step.py > gpurun_out/prof_step.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
