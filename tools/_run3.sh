mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loudness.py -x -q -m gpu -k "rows or dispatch or tile_kernel or batch_parity or chunking" 2>&1 | tail -8 > gpurun_out/r3_pytest.log
echo "pytest rc=$?" >> gpurun_out/r3_pytest.log
timeout 300 python bench.py --config cfg4 --steps 10 --warmup 3 > gpurun_out/r3_bench_cfg4.json 2> gpurun_out/r3_bench_cfg4.err
cat gpurun_out/r3_pytest.log
python - <<PY
import json
for l in open("gpurun_out/r3_bench_cfg4.json"):
    if l.startswith("{"):
        d=json.loads(l); print("cfg4", d["value"], d["ms_per_step"], json.dumps(d.get("detail",{}))[150:600])
PY
