#!/bin/bash
# 4-GPU run: pipeline tests (2 ranks), then bench at N=4 with queue depth 1 and 3, N=2, N=1
mkdir -p gpurun_out
echo "== pipeline tests"; timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | grep -v Warning | tail -4
bench() { # n depth
  local n=$1 d=$2 tag=n$1_d$2
  PIPEEDGE_QUEUE_DEPTH=$d timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n*10+d)) bench.py --gpus $n --steps 300 --warmup 20 > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/scale_$tag.json').read().strip().splitlines()[-1]); print('N=$n depth $d value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), d.get('stage_stats'))
except Exception as e: print('parse fail', e)"
  grep -v Warning gpurun_out/scale_$tag.err | tail -3
}
bench 4 1
bench 4 3
bench 2 1
