#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02_n2c; mkdir -p $O
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 1 $O/$name.log | cut -c1-200; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
t bench_driver 300 $TR --nproc-per-node 2 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 5
t ref_arm 400 $TR --nproc-per-node 2 --master-port 29702 bench.py --impl reference --gpus 2 --steps 20 --warmup 5
cat $O/summary.txt
