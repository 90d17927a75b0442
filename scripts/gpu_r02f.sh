#!/bin/bash
# Round-2 GPU call F (one GPU): overlapped send (two buffer parities) - pipeline tests on one GPU, N=1 bench both ways.
cd "$(dirname "$0")/.."
O=gpurun_out/r02f; mkdir -p $O
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 2 $O/$name.log | cut -c1-300; }
t pipes 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k "native"
t link 600 python -m pytest tests/test_link_gpu.py -q -m gpu
t bench_300 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
PIPEEDGE_OVERLAP_SEND=0 t bench_300_noov 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
t bench_driver 600 python bench.py --gpus 1 --steps 20 --warmup 5
t n8_on_one_gpu 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5
cat $O/summary.txt
