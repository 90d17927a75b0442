#!/bin/bash
mkdir -p gpurun_out
echo "== tcgen05 tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "tcgen05 or resid_in_place" 2>&1 | tail -8
echo "== forced cluster configs"; for f in "2,1,128" "1,2,128" "2,2,128" "4,1,256" "2,2,192"; do echo "-- $f"; PE_GEMM_FORCE=$f timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "tcgen05_f32 or epilogues" 2>&1 | tail -3; done
echo "== sweep"; timeout 600 python scripts/gemm_sweep.py 2>&1 | tee gpurun_out/gemm_sweep.txt | tail -60
echo "== shards"; timeout 600 python -m pytest tests/test_shards_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5
echo "== bench N=1"; timeout 900 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_n1_v2.json 2> gpurun_out/bench_n1_v2.err; tail -c 2600 gpurun_out/bench_n1_v2.json; tail -5 gpurun_out/bench_n1_v2.err
