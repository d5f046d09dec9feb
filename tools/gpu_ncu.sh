#!/bin/bash
# usage: tools/gpu_ncu.sh TAG [kernel-regex]  -- ncu --set full of the particle kernels (one launch each) of the current build
TAG=${1:-dev}; RX=${2:-"k_p2g|k_forces|k_g2p"}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s ${SKIP:-12} -c ${CNT:-3} -o gpurun_out/${TAG}_full -f python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
tail -n 2 gpurun_out/ncu_${TAG}.log
