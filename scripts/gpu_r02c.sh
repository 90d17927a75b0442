#!/bin/bash
# Round-2 GPU call C (one GPU): re-validate after the link-kernel / LayerNorm rewrites; short.
cd "$(dirname "$0")/.."
O=gpurun_out/r02c; mkdir -p $O
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 3 $O/$name.log; }
t link 600 python -m pytest tests/test_link_gpu.py tests/test_kernels_gpu.py -q -m gpu
t pipes 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k "native"
t c5 600 python -m pytest tests/test_shards_gpu.py -q -m gpu -k "C5 or C3"
t bench_300 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
PE_ATTN_TCGEN05=1 t bench_300_attn 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
t attn_cmp 300 python scripts/attention_compare.py
t ncu_full 900 ncu --set full --clock-control none --profile-from-start off -k regex:"link_|layernorm|quant_" -f -o $O/kernels python scripts/profile_kernels.py
t n8_on_one_gpu 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 --workload deit-base-q8 --ubatch 8
cat $O/summary.txt
