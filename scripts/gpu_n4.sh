#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 300 --warmup 20 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "== bench N=4 exit $?"; tail -c 400 gpurun_out/bench_n4.json; grep -v Warning gpurun_out/bench_n4.err | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/bench_ref_n4.json 2> gpurun_out/bench_ref_n4.err
echo "== reference arm N=4 exit $?"; tail -c 300 gpurun_out/bench_ref_n4.json
