#!/bin/bash
# probe of the GPU box: toolchains the oracle pinning would need, host topology, FP64 loop microbenchmark
mkdir -p gpurun_out
{
  echo "== toolchains"; for t in cargo rustc rustup go javac node clang; do printf "%s: " $t; command -v $t || echo ABSENT; done
  ls -d ~/.cargo ~/.rustup /usr/local/cargo /opt/rust* 2>&1 | head
  echo "== host"; nproc; lscpu | head -25; numactl -H 2>/dev/null | head -12
  echo "== gpu"; nvidia-smi -L; nvidia-smi topo -m 2>&1 | head -20
  nvidia-smi --query-gpu=index,pci.bus_id,clocks.sm,clocks.max.sm,power.limit --format=csv
} > gpurun_out/probe.txt 2>&1
timeout 120 tools/mb_filter > gpurun_out/mb_filter.txt 2>&1
cat gpurun_out/probe.txt | head -60; cat gpurun_out/mb_filter.txt
