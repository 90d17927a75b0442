#!/bin/bash
# Round-2 GPU call on EIGHT GPUs: the scaling curve of the default workload under the driver's exact command, and
# BASELINE configs 3 (ViT-L x4), 4 (BERT-base x8) and 5 (DeiT-B x8, 8-bit hops).
cd "$(dirname "$0")/.."
O=gpurun_out/r02_n8; mkdir -p $O
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 1 $O/$name.log | cut -c1-300; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
t n8_driver 300 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 5
t n8_300 300 $TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 --steps 300 --warmup 20
t n4_driver 300 $TR --nproc-per-node 4 --master-port 29803 bench.py --gpus 4 --steps 20 --warmup 5
t n4_300 300 $TR --nproc-per-node 4 --master-port 29804 bench.py --gpus 4 --steps 300 --warmup 20
t c4_bert_n8 300 $TR --nproc-per-node 8 --master-port 29805 bench.py --gpus 8 --steps 200 --warmup 20 --workload bert-base
t c5_deit_q8_n8 300 $TR --nproc-per-node 8 --master-port 29806 bench.py --gpus 8 --steps 200 --warmup 20 --workload deit-base-q8
t c3_vitl_n4 400 $TR --nproc-per-node 4 --master-port 29807 bench.py --gpus 4 --steps 200 --warmup 20 --workload vit-large
t n2_driver 300 $TR --nproc-per-node 2 --master-port 29808 bench.py --gpus 2 --steps 20 --warmup 5
cat $O/summary.txt
