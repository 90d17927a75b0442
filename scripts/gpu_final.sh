#!/bin/bash
# Round-end validation on one GPU: every GPU test file, smoke, the bench line, ncu launch list + one full capture
mkdir -p gpurun_out
bash scripts/gpu_check.sh 2>&1 | grep "^== \|passed\|failed\|error" 
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "exit $?"; tail -c 2500 gpurun_out/bench_n1.json; grep -v Warning gpurun_out/bench_n1.err | tail -5
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 6 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
echo "== ncu full (4 consecutive GEMMs: QKV, out, FC1, FC2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 100 -c 4 -f -o gpurun_out/prof_gemm python bench.py --quick --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | head -40
