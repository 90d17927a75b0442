#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "== $name: exit $?"; tail -n 45 "gpurun_out/$name.log"; }
run pdl_kernels 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x
PE_GEMM_STATIC_W=1 run sweep 600 python scripts/plan_sweep.py fc1
run pdl_shards 900 python -m pytest tests/test_shards_gpu.py -q -m gpu -p no:cacheprovider -x
run bench_pdl 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline
