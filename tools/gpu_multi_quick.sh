#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/test_peer_gather.py > gpurun_out/peer_gather_n$N.log 2>&1; echo "peer gather rc=$?" >> gpurun_out/peer_gather_n$N.log
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?" >> gpurun_out/bench_n$N.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu > gpurun_out/bench_n1_lean.json 2> gpurun_out/bench_n1_lean.err; echo "bench1 rc=$?" >> gpurun_out/bench_n1_lean.err
grep "peer gather" gpurun_out/peer_gather_n$N.log; tail -n 2 gpurun_out/bench_n$N.err; cut -c1-300 gpurun_out/bench_n$N.json; cut -c1-300 gpurun_out/bench_n1_lean.json
