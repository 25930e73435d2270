mkdir -p gpurun_out
V=soundscope_b200/_variants
( SSB_LIB=$PWD/$V/lib_TPSPLIT.so timeout 600 python -m pytest tests/test_gpu_wtile.py -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/r4_pytest_tpsplit.log
( SSB_LIB=$PWD/$V/lib_ANYPAIR11.so timeout 600 python -m pytest tests/test_gpu_loudness.py -x -q -m gpu -k "rows_any or dispatch or batch_parity" 2>&1 | tail -4 ) > gpurun_out/r4_pytest_anypair.log
( SSB_LIB=$PWD/$V/lib_F32M2_PAIR.so timeout 600 python -m pytest tests/test_gpu_loudness.py -x -q -m gpu -k "rows or dispatch" 2>&1 | tail -4 ) > gpurun_out/r4_pytest_f32pair.log
echo "== default" > gpurun_out/r4_time.log
MODES=all FORCES=5 timeout 200 python tools/time_wtile.py >> gpurun_out/r4_time.log 2>&1
echo "== TPSPLIT" >> gpurun_out/r4_time.log
SSB_LIB=$PWD/$V/lib_TPSPLIT.so MODES=all FORCES=5 timeout 200 python tools/time_wtile.py >> gpurun_out/r4_time.log 2>&1
for v in F32M2 F32M2_PAIR; do
  SSB_LIB=$PWD/$V/lib_$v.so timeout 300 python bench.py --config cfg4 --steps 10 --warmup 3 > gpurun_out/r4_cfg4_$v.json 2> gpurun_out/r4_cfg4_$v.err
done
for c in cfg5 cfg5_x4; do
  SSB_LIB=$PWD/$V/lib_ANYPAIR11.so timeout 300 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r4_${c}_ANYPAIR11.json 2> gpurun_out/r4_${c}_ANYPAIR11.err
done
tail -n 3 gpurun_out/r4_pytest_*.log; cat gpurun_out/r4_time.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r4_cfg*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l)["detail"]; print(f, "step_ms", round(d["step_ms"],3), "kernel_ms", round(d["kernel_ms"],3), "frac", round(d["frac_of_hbm_peak"],3))
PY
