#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02k
export PIPEEDGE_LINK_TIMEOUT_S=60
timeout 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k "native" > gpurun_out/r02k/pipes.log 2>&1; echo "rc=$? pipes"; tail -2 gpurun_out/r02k/pipes.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02k/bench.log 2>&1; echo "rc=$? bench"; tail -1 gpurun_out/r02k/bench.log | cut -c1-200
