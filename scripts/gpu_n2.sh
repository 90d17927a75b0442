#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider -x -k "adaptive and CONTROLLER" > gpurun_out/pipe_adaptive.log 2>&1
echo "== adaptive CLI tests: exit $?"; grep -v Warning gpurun_out/pipe_adaptive.log | tail -40
