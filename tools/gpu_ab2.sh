#!/bin/bash
# usage: tools/gpu_ab2.sh TAG variant...   -- A/B of library variants (make -C .../csrc variant-NAME VFLAGS=...): rest and pinned-dt flowing bench lines
TAG=${1:-dev}; shift
mkdir -p gpurun_out
for v in "$@"; do
  L=""; [ $v != default ] && L="$PWD/anisotropicelastoplasticity_b200/libaep_b200_$v.so"
  AEP_B200_LIB=$L timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest_$v.txt 2>&1; echo "== $v"; cut -c1-520 gpurun_out/bench_${TAG}_rest_$v.txt
  AEP_B200_LIB=$L timeout 600 python bench.py --steps 60 --warmup 5 --quick --pin-dt 1.5e-5 > gpurun_out/bench_${TAG}_pin_$v.txt 2>&1; cut -c1-520 gpurun_out/bench_${TAG}_pin_$v.txt
done
