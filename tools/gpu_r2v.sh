#!/bin/bash
# CTAs per SM of the gather kernels: A/B at rest and at a pinned dt
TAG=${1:-r2v}; mkdir -p gpurun_out
for V in default f6 f8 c6; do
  L=""; [ $V != default ] && L="$PWD/anisotropicelastoplasticity_b200/libaep_b200_$V.so"
  AEP_B200_LIB=$L timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest_$V.txt 2>&1; echo "== $V"; cut -c1-420 gpurun_out/bench_${TAG}_rest_$V.txt
  AEP_B200_LIB=$L timeout 600 python bench.py --steps 60 --warmup 5 --quick --pin-dt 1.5e-5 > gpurun_out/bench_${TAG}_pin_$V.txt 2>&1; cut -c1-420 gpurun_out/bench_${TAG}_pin_$V.txt
done
