#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python scripts/link_bench.py > $O/link_bench.log 2>&1; echo "rc=$?"; cat $O/link_bench.log
