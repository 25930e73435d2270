mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2b_pytest.log
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
FORCES=5 timeout 300 python tools/time_wtile.py > gpurun_out/r2b_time.log 2>&1
for c in cfg4 cfg5 cfg5_x4; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2b_bench_$c.json 2> gpurun_out/r2b_bench_$c.err
done
cat gpurun_out/r2b_pytest.log gpurun_out/r2b_time.log
for c in cfg4 cfg5 cfg5_x4; do python - <<PY
import json
for l in open("gpurun_out/r2b_bench_$c.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$c", d["value"], d["ms_per_step"], json.dumps(d.get("detail",{}))[:700])
PY
done
