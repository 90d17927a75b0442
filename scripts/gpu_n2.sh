#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "== bench N=2 exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), json.dumps(d['host_us_per_step']))"; grep -v Warning gpurun_out/bench_n2.err | tail -4
