#!/bin/bash
# Round-2 GPU call on TWO GPUs: the NCCL / Python-thread path's tests (need one GPU per rank), the native pipeline across
# two real GPUs (cudaIpc peer mapping over NVLink), BASELINE config 2.
cd "$(dirname "$0")/.."
O=gpurun_out/r02_n2; mkdir -p $O
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 2 $O/$name.log | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
t native2 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k "native and (cuts0 or cuts1 or cuts2)"
t bench_driver 300 $TR --nproc-per-node 2 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 5
t bench_300 300 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --steps 300 --warmup 20
t bench_q8 300 $TR --nproc-per-node 2 --master-port 29703 bench.py --gpus 2 --steps 100 --warmup 10 --workload deit-base-q8
t nccl_tests 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k "nccl or adaptive"
cat $O/summary.txt
