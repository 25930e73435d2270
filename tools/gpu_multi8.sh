N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo "bench rc=$?" >> gpurun_out/final_bench_n$N.err
tail -n 2 gpurun_out/final_bench_n$N.err; grep "^{" gpurun_out/final_bench_n$N.json | cut -c1-300
