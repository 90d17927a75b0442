#!/bin/bash
# 2-GPU: pipeline tests three times over (teardown flakiness), then the N=2 bench with its exit code
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pipe_$i.log 2>&1
  echo "== pipeline tests run $i: exit $?"; grep -v Warning gpurun_out/pipe_$i.log | tail -5
done
PIPEEDGE_QUEUE_DEPTH=3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "== bench N=2 exit $?"; tail -c 600 gpurun_out/bench_n2.json; grep -v Warning gpurun_out/bench_n2.err | tail -5
