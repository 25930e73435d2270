mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/final_pytest.log
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/final_ncu_bench.log 2>&1
for c in cfg4 cfg5 cfg5_x4; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/final_bench_$c.json 2> gpurun_out/final_bench_$c.err
done
cat gpurun_out/final_pytest.log gpurun_out/final_smoke.log; head -c 3000 gpurun_out/final_bench_n1.json; echo; tail -3 gpurun_out/final_bench_n1.err
