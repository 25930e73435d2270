#!/bin/bash
# full verification pass: parity suite, smoke, bench lines (N=1 with extras, reference arm)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/bench_ref.err
tail -n 5 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; tail -n 5 gpurun_out/bench_n1.err; cut -c1-1500 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_ref.err; cut -c1-600 gpurun_out/bench_ref.json
