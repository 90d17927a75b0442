#!/bin/bash
# N-GPU pipeline bench (run under gpurun --gpus N): bash scripts/gpu_scale.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$N" -ge 2 ]; then echo "== pipeline tests"; timeout 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4; fi
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    echo "== bench N=$n"
    if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 300 --warmup 20 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
    python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1]); print('N=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3))
except Exception as e: print('parse fail', e)"
    tail -3 gpurun_out/scale_n$n.err
  fi
done
