#!/bin/bash
# Final check of HEAD on eight GPUs: the driver's exact commands (both arms) at N = 8 and N = 4.
cd "$(dirname "$0")/.."
O=gpurun_out/r02_n8c; mkdir -p $O
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 1 $O/$name.log | cut -c1-200; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
t n8_driver 300 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 5
t n4_driver 300 $TR --nproc-per-node 4 --master-port 29803 bench.py --gpus 4 --steps 20 --warmup 5
cat $O/summary.txt
