#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for b in 1 0; do
timeout 600 python bench.py --res 512 --steps 12 --warmup 4 --quick --sort-bricks $b > gpurun_out/bench_b${b}_pad.txt 2>&1
AEP_B200_LIB=$PWD/anisotropicelastoplasticity_b200/libaep_b200_nopad.so timeout 600 python bench.py --res 512 --steps 12 --warmup 4 --quick --sort-bricks $b > gpurun_out/bench_b${b}_nopad.txt 2>&1
done
tail -n 12 gpurun_out/pytest_gpu.txt; for f in gpurun_out/bench_b*; do echo $f; cut -c1-420 $f; done
