#!/bin/bash
# multi-GPU pass (run with gpurun --gpus N): peer-memory gather test, bench at N, shard configs at N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/test_peer_gather.py > gpurun_out/peer_gather_n$N.log 2>&1; echo "peer gather rc=$?" >> gpurun_out/peer_gather_n$N.log
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?" >> gpurun_out/bench_n$N.err
for cfg in cfg4 cfg5 cfg5_x4; do
  timeout 600 $TR bench.py --gpus $N --config $cfg --steps 10 --warmup 3 > gpurun_out/bench_${cfg}_n$N.json 2> gpurun_out/bench_${cfg}_n$N.err; echo "$cfg rc=$?" >> gpurun_out/bench_${cfg}_n$N.err
done
tail -n 6 gpurun_out/peer_gather_n$N.log; tail -n 3 gpurun_out/bench_n$N.err; cut -c1-700 gpurun_out/bench_n$N.json
for cfg in cfg4 cfg5 cfg5_x4; do tail -n 2 gpurun_out/bench_${cfg}_n$N.err; cut -c1-400 gpurun_out/bench_${cfg}_n$N.json; done
