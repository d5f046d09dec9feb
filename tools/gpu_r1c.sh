#!/bin/bash
# v3 kernels (brick order, interior gathers, s-trick): parity tests, sort_every sweep, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for se in 1 2 4 8; do
timeout 600 python bench.py --res 512 --steps 24 --warmup 8 --quick --sort-every $se > gpurun_out/bench512_se$se.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v3.csv python bench.py --res 512 --steps 3 --warmup 3 --quick > gpurun_out/ncu_launch.log 2>&1
tail -n 30 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench512_se*.txt | cut -c1-700
