N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/test_peer_gather.py > gpurun_out/final_peer_gather_n$N.log 2>&1; echo "peer gather rc=$?" >> gpurun_out/final_peer_gather_n$N.log
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo "bench rc=$?" >> gpurun_out/final_bench_n$N.err
grep "peer gather" gpurun_out/final_peer_gather_n$N.log; tail -n 2 gpurun_out/final_bench_n$N.err; grep "^{" gpurun_out/final_bench_n$N.json | cut -c1-400
