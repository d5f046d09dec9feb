#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -s -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for se in 1 2 4 8; do
timeout 600 python bench.py --res 512 --steps 24 --warmup 8 --quick --sort-every $se > gpurun_out/bench512_se$se.txt 2>&1
done
tail -n 30 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench512_se*.txt | cut -c1-900
