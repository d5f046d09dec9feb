#!/bin/bash
# v5: smem-tile gathers + adaptive re-sort
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for se in 1 0; do
timeout 600 python bench.py --res 512 --steps 40 --warmup 8 --quick --sort-every $se > gpurun_out/bench_v5_se$se.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v5.csv python bench.py --res 512 --steps 3 --warmup 3 --quick --sort-every 1 > gpurun_out/ncu_launch.log 2>&1
tail -n 12 gpurun_out/pytest_gpu.txt; cut -c1-420 gpurun_out/bench_v5_se*.txt
