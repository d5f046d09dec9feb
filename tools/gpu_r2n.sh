#!/bin/bash
# A/B of the F_P prefetch at a pinned dt, collider tests
TAG=${1:-r2n}; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_colliders.py -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
tail -n 12 gpurun_out/pytest_${TAG}.txt | cut -c1-300
for V in default nofp; do
  L=""; [ $V != default ] && L="$PWD/anisotropicelastoplasticity_b200/libaep_b200_$V.so"
  AEP_B200_LIB=$L timeout 600 python bench.py --steps 60 --warmup 5 --quick --pin-dt 1.5e-5 > gpurun_out/bench_${TAG}_pin_$V.txt 2>&1; echo "== $V"; cut -c1-700 gpurun_out/bench_${TAG}_pin_$V.txt
done
AEP_B200_LIB= timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest.txt 2>&1; cut -c1-500 gpurun_out/bench_${TAG}_rest.txt
