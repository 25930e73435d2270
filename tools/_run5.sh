mkdir -p gpurun_out
./tools/mb_fma > gpurun_out/mb_fma.txt 2>&1
N_LAUNCH=12 FORCE=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_loudness_wtile -s 8 -c 2 \
  -o gpurun_out/prof_wtile_lean -f python tools/prof_cfg2.py > gpurun_out/prof_wtile_lean.log 2>&1
cat gpurun_out/mb_fma.txt; tail -3 gpurun_out/prof_wtile_lean.log
