#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "== pipeline tests (2 GPUs)"; timeout 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -30
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json; tail -15 gpurun_out/bench_n2.err
