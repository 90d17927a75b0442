#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02j
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02j/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02j/smoke.log
