#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k attention > gpurun_out/hd80_attn.log 2>&1; echo "== attention: exit $?"; tail -n 15 gpurun_out/hd80_attn.log
timeout 300 python -m pytest tests/test_shards_gpu.py -q -m gpu -p no:cacheprovider -k "huge" > gpurun_out/hd80_shards.log 2>&1; echo "== shards: exit $?"; tail -n 25 gpurun_out/hd80_shards.log
