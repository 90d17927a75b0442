#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02_n2b; mkdir -p $O
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 1 $O/$name.log | cut -c1-200; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
t bench_driver 300 $TR --nproc-per-node 2 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 5
t bench_300 300 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --steps 300 --warmup 20
t native2 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k "native and (cuts0 or cuts1 or cuts2)"
cat $O/summary.txt
