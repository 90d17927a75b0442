#!/bin/bash
# Round-2 GPU call B (one GPU): fixed link tests, bench lines after the results-wait fix, attention v2, ncu evidence,
# the 6-rank pipeline test with diagnostics, then the whole GPU suite.
cd "$(dirname "$0")/.."
O=gpurun_out/r02b; mkdir -p $O
export PIPEEDGE_LINK_TIMEOUT_S=60
t() { local name=$1; shift; local lim=$1; shift; echo "== $name"; timeout $lim "$@" > $O/$name.log 2>&1; echo "rc=$? $name" | tee -a $O/summary.txt; tail -n 3 $O/$name.log; }
t link 600 python -m pytest tests/test_link_gpu.py -q -m gpu
t bench_driver 600 python bench.py --gpus 1 --steps 20 --warmup 5
t bench_300 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
PE_ATTN_TCGEN05=1 t bench_300_attn 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline
PE_ATTN_TCGEN05=1 t attn 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention_tcgen05
t attn_cmp 300 python scripts/attention_compare.py
t bench_b1 600 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline --workload vit-base-b1
# ncu: launch list of 4 graph-mode forwards (shares), then the full set for one launch of every kernel
t ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv python bench.py --quick --steps 4 --warmup 3
PE_ATTN_TCGEN05=1 t ncu_launches_attn 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_attn.csv python bench.py --quick --steps 4 --warmup 3
t ncu_full 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/kernels python scripts/profile_kernels.py
PE_ATTN_TCGEN05=1 t ncu_full_attn 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention -f -o $O/attention_tcgen05 python scripts/profile_kernels.py
ls -la $O/*.ncu-rep >> $O/summary.txt 2>&1
t native6 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -k "native_pipeline and cuts4" -x
t all_gpu_tests 1500 python -m pytest tests -q -m gpu --deselect "tests/test_pipeline_gpu.py::test_native_pipeline_is_bit_identical_to_local_shards[test/vit-tiny-cuts4-qbits4]"
cat $O/summary.txt
