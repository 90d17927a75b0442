mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; timeout "$t" python -m pytest "$@" -q -m gpu -p no:cacheprovider > "gpurun_out/$name.log" 2>&1; echo "== $name: exit $?"; tail -n 40 "gpurun_out/$name.log"; }
run kernels_tcgen05 300 tests/test_kernels_gpu.py -k "tcgen05 or resid_in_place or rejects"
run quant 600 tests/test_quant_gpu.py
run shards 900 tests/test_shards_gpu.py
