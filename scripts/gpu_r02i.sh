#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02i; mkdir -p $O
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 1 $O/$name.log | cut -c1-200; }
t ref_arm 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5
t bench_driver 600 python bench.py --gpus 1 --steps 20 --warmup 5
cat $O/summary.txt
