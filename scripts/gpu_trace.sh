#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "tcgen05 or resid_in_place" > gpurun_out/k.log 2>&1; tail -3 gpurun_out/k.log; grep -E "^(FAILED|E  )" gpurun_out/k.log | head -20
timeout 300 python scripts/gemm_trace.py 2>&1 | tee gpurun_out/gemm_trace4.txt | grep mainloop
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_n1_v3.json 2> gpurun_out/bench_n1_v3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_v3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:(round(v['avg_us'],1), round(v.get('tflops',0))) for k,v in d['roofline']['breakdown'].items()})"
