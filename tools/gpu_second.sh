#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_report.py > gpurun_out/report.txt 2>&1; echo "report rc=$?" >> gpurun_out/report.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -s -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --res 512 --steps 20 --warmup 5 > gpurun_out/bench512.txt 2>&1; echo "bench rc=$?" >> gpurun_out/bench512.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --res 512 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
for k in k_forces k_g2p k_p2g; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k python bench.py --res 512 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_$k.log 2>&1
done
head -c 1500 gpurun_out/report.txt; echo; tail -n 25 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench512.txt | head -c 4000
