#!/bin/bash
# usage: tools/gpu_quick.sh TAG   -- GPU parity suite + quick 1-GPU bench line (stage timers) for an in-progress build
TAG=${1:-dev}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -x -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
timeout 600 python bench.py --res 512 --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}.txt 2>&1
tail -n 12 gpurun_out/pytest_${TAG}.txt; cut -c1-700 gpurun_out/bench_${TAG}.txt
