#!/bin/bash
# Run the GPU test files one process each (a device-side trap in one must not poison the others);
# logs go to gpurun_out/. Usage: scripts/gpu_check.sh [pytest -k expression for the kernels file]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))" >> gpurun_out/gpu.txt 2>&1
rc=0
run() {  # name, timeout, pytest args...
  local name=$1 t=$2; shift 2
  timeout "$t" python -m pytest "$@" -q -m gpu -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  local r=$?
  echo "== $name: exit $r"; tail -n 25 "gpurun_out/$name.log"
  [ $r -ne 0 ] && rc=1
}
run kernels_basic 300 tests/test_kernels_gpu.py -k "layernorm or simt or attention"
run quant 600 tests/test_quant_gpu.py
run kernels_tcgen05 300 tests/test_kernels_gpu.py -k "not (layernorm or simt or attention)"
run shards 900 tests/test_shards_gpu.py
exit $rc
