#!/bin/bash
# Round-2 GPU call A (one GPU): link kernels, tcgen05 attention, native pipeline with every rank on this GPU, bench lines.
cd "$(dirname "$0")/.."
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu.txt 2>&1
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 3 $O/$name.log; }
t link 600 python -m pytest tests/test_link_gpu.py -x -q -m gpu
t attn 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention_tcgen05
t single 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k single_rank
t native 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k native_pipeline -x
t bench_driver 600 python bench.py --gpus 1 --steps 20 --warmup 5
t bench_300 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
PE_ATTN_TCGEN05=1 t bench_300_attn 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
t attn_cmp 300 python scripts/attention_compare.py
t bench_b1 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline --workload vit-base-b1
t n8_on_one_gpu 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5
t n2_q8_on_one_gpu 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --workload deit-base-q8 --ubatch 8
t smoke 300 python __graft_entry__.py smoke
t all_gpu_tests 1500 python -m pytest tests -q -m gpu -x
cat $O/summary.txt
