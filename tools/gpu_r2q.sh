#!/bin/bash
# un-fused default: full GPU suite; A/B G2P at 4 vs 5 CTAs per SM
TAG=${1:-r2q}; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
tail -n 8 gpurun_out/pytest_${TAG}.txt | cut -c1-300
for V in default ctas5; do
  L=""; [ $V != default ] && L="$PWD/anisotropicelastoplasticity_b200/libaep_b200_$V.so"
  AEP_B200_LIB=$L timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest_$V.txt 2>&1; echo "== $V"; cut -c1-560 gpurun_out/bench_${TAG}_rest_$V.txt
  AEP_B200_LIB=$L timeout 600 python bench.py --steps 60 --warmup 5 --quick --pin-dt 1.5e-5 > gpurun_out/bench_${TAG}_pin_$V.txt 2>&1; cut -c1-560 gpurun_out/bench_${TAG}_pin_$V.txt
done
