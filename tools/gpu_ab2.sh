#!/bin/bash
# usage: tools/gpu_ab2.sh TAG variant...   -- quick bench line per build variant (no tests)
TAG=${1:-dev}; shift
mkdir -p gpurun_out
for v in "$@"; do
  AEP_B200_LIB=$PWD/anisotropicelastoplasticity_b200/libaep_b200_$v.so timeout 600 python bench.py --res 512 --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}_$v.txt 2>&1; echo $v; cut -c1-420 gpurun_out/bench_${TAG}_$v.txt
done
