#!/bin/bash
# final pass of the session: A/B of the tile kernel's producer-warp placement, full parity suite, smoke, bench lines,
# launch list of bench-shaped steps, ncu --set full of the multichannel kernel
mkdir -p gpurun_out
P=$PWD/soundscope_b200/libssb_P.so
{
echo "== cfg2 base"; timeout 200 python tools/time_cfg2.py
echo "== cfg2 producer-on-idle-warp"; timeout 200 env SSB_LIB=$P python tools/time_cfg2.py
echo "== cfg2 base (repeat)"; timeout 200 python tools/time_cfg2.py
echo "== cfg2 producer-on-idle-warp (repeat)"; timeout 200 env SSB_LIB=$P python tools/time_cfg2.py
echo "== cfg2 all base"; timeout 200 python tools/time_cfg2.py --all
echo "== cfg2 all producer-on-idle-warp"; timeout 200 env SSB_LIB=$P python tools/time_cfg2.py --all
} > gpurun_out/variants4.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step.csv python tools/prof_step.py > gpurun_out/prof_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_loudness_rows_any" -c 4 -f -o gpurun_out/prof_any python tools/prof_extra.py > gpurun_out/prof_any.log 2>&1
cat gpurun_out/variants4.log; tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_n1.json
