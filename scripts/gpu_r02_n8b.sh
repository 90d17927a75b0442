#!/bin/bash
# Round-2 GPU call on EIGHT GPUs, after the overlapped send: scaling curve under the driver's command + BASELINE C3 / C4 / C5.
cd "$(dirname "$0")/.."
O=gpurun_out/r02_n8b; mkdir -p $O
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 1 $O/$name.log | cut -c1-200; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
t n8_driver 300 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 5
t n8_300 300 $TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 --steps 300 --warmup 20
t n4_driver 300 $TR --nproc-per-node 4 --master-port 29803 bench.py --gpus 4 --steps 20 --warmup 5
t n4_300 300 $TR --nproc-per-node 4 --master-port 29804 bench.py --gpus 4 --steps 300 --warmup 20
t c4_bert_n8 300 $TR --nproc-per-node 8 --master-port 29805 bench.py --gpus 8 --steps 200 --warmup 20 --workload bert-base
t c5_deit_q8_n8 300 $TR --nproc-per-node 8 --master-port 29806 bench.py --gpus 8 --steps 200 --warmup 20 --workload deit-base-q8
t c3_vitl_n4 400 $TR --nproc-per-node 4 --master-port 29807 bench.py --gpus 4 --steps 200 --warmup 20 --workload vit-large
PIPEEDGE_WIRE_F16=1 t n8_300_wire16 300 $TR --nproc-per-node 8 --master-port 29808 bench.py --gpus 8 --steps 300 --warmup 20
cat $O/summary.txt
