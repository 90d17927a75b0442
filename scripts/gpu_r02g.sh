#!/bin/bash
# Round-2 final GPU call (one GPU): the whole GPU suite at HEAD, the driver's bench command, the evidence pack
# (launch list of graph-mode forwards, `ncu --set full` of one launch of every kernel on the path).
cd "$(dirname "$0")/.."
O=gpurun_out/r02g; mkdir -p $O
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 2 $O/$name.log | cut -c1-300; }
t bench_driver 600 python bench.py --gpus 1 --steps 20 --warmup 5
t ref_arm 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5
t ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv python bench.py --quick --steps 4 --warmup 3
t ncu_full 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/kernels python scripts/profile_kernels.py
t link_bench_default 300 env PE_ONLY_DEFAULT=1 python scripts/link_bench.py
t smoke 300 python __graft_entry__.py smoke
t all_gpu_tests 1500 python -m pytest tests -q -m gpu
cat $O/summary.txt
