#!/bin/bash
# 2-GPU checks (run under gpurun --gpus 2): pipeline parity + adaptive-policy CLI tests, then the N=2 bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pipeline.log 2>&1
echo "== pipeline tests: exit $?"; grep -v Warning gpurun_out/pipeline.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "== bench N=2 exit $?"; tail -c 1200 gpurun_out/bench_n2.json; grep -v Warning gpurun_out/bench_n2.err | tail -4
