#!/bin/bash
# Round-2 GPU call D (one GPU): fused projection + residual + LayerNorm (PE_FUSE_LN=1), link kernels after the fence fix.
cd "$(dirname "$0")/.."
O=gpurun_out/r02d; mkdir -p $O
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 3 $O/$name.log; }
t kernels 900 python -m pytest tests/test_kernels_gpu.py tests/test_link_gpu.py -q -m gpu
t bench_300 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
PE_FUSE_LN=1 t bench_300_fuse 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
PE_FUSE_LN=1 t shards_fuse 900 python -m pytest tests/test_shards_gpu.py -q -m gpu -x
PE_FUSE_LN=1 t pipes_fuse 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k "native"
t ncu_link 900 ncu --set full --clock-control none --profile-from-start off -k regex:"link_" -f -o $O/link_kernels python scripts/profile_kernels.py
PE_FUSE_LN=1 t ncu_launches_fuse 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_fuse.csv python bench.py --quick --steps 4 --warmup 3
cat $O/summary.txt
