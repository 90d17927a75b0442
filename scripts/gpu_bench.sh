#!/bin/bash
# smoke + bench + ncu launch list + one full ncu capture of the dominant kernel; outputs in gpurun_out/
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== graph/api tests"; timeout 300 python -m pytest tests/test_shards_gpu.py -q -m gpu -p no:cacheprovider -k "graph or api" 2>&1 | tail -5
echo "== bench N=1"; timeout 900 python bench.py --steps 300 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
echo "== ncu full (FC1 GEMM)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 100 -c 4 -f -o gpurun_out/prof_gemm python bench.py --quick --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
