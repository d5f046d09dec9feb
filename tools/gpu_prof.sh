#!/bin/bash
# usage: tools/gpu_prof.sh TAG  -- parity suite, quick bench, ncu --set full of the three particle kernels (one launch each)
TAG=${1:-dev}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -x -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
timeout 600 python bench.py --res 512 --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_p2g|k_forces|k_g2p" -s 12 -c 3 -o gpurun_out/${TAG}_full -f python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
tail -n 12 gpurun_out/pytest_${TAG}.txt; cut -c1-700 gpurun_out/bench_${TAG}.txt; tail -n 2 gpurun_out/ncu_${TAG}.log
