#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --res 512 --steps 16 --warmup 8 --quick --sort-every 1 > gpurun_out/bench_v6_tile.txt 2>&1
AEP_B200_LIB=$PWD/anisotropicelastoplasticity_b200/libaep_b200_notile.so timeout 600 python bench.py --res 512 --steps 16 --warmup 8 --quick --sort-every 1 > gpurun_out/bench_v6_notile.txt 2>&1
tail -n 8 gpurun_out/pytest_gpu.txt; cut -c1-420 gpurun_out/bench_v6_*.txt
