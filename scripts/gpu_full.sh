#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_check.sh 2>&1 | grep -E "^==|passed|failed|FAILED" 
timeout 300 python scripts/gemm_trace.py 2>&1 > gpurun_out/gemm_trace5.txt; grep -A1 "plan" gpurun_out/gemm_trace5.txt | grep -v "^--" | paste - - | sed 's/start: med 0 max 0//' | cut -c1-330
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_n1_v4.json 2> gpurun_out/bench_n1_v4.err; tail -3 gpurun_out/bench_n1_v4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_v4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:(round(v['avg_us'],1), round(v.get('tflops',0))) for k,v in d['roofline']['breakdown'].items()})"
