#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02h
timeout 200 python scripts/cpu_arm_trajectory.py cuda > gpurun_out/r02h/traj_cuda.log 2>&1; tail -2 gpurun_out/r02h/traj_cuda.log
timeout 200 python scripts/cpu_arm_trajectory.py > gpurun_out/r02h/traj_nocuda.log 2>&1; tail -2 gpurun_out/r02h/traj_nocuda.log
nproc; lscpu | grep -E "Model name|Thread|Core|Socket|NUMA" 
