#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 300 --warmup 20 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "== bench N=4 exit $?"; tail -c 1500 gpurun_out/bench_n4.json; grep -v Warning gpurun_out/bench_n4.err | tail -30
