#!/bin/bash
mkdir -p gpurun_out
echo "== pipeline tests (2 GPUs)"; timeout 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | grep -v Warning | tail -6
for d in 1 4; do
echo "== bench N=2 depth $d"; PIPEEDGE_QUEUE_DEPTH=$d timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$d bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/bench_n2_d$d.json 2> gpurun_out/bench_n2_d$d.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2_d$d.json').read().strip().splitlines()[-1]); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3))"
grep -v Warning gpurun_out/bench_n2_d$d.err | tail -4
done
