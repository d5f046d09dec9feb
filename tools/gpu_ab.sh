#!/bin/bash
# usage: tools/gpu_ab.sh TAG variant...   -- parity suite on the default build, then a quick bench line per build variant
TAG=${1:-dev}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -x -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
tail -n 6 gpurun_out/pytest_${TAG}.txt
timeout 600 python bench.py --res 512 --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}_default.txt 2>&1; echo default; cut -c1-420 gpurun_out/bench_${TAG}_default.txt
for v in "$@"; do
  AEP_B200_LIB=$PWD/anisotropicelastoplasticity_b200/libaep_b200_$v.so timeout 600 python bench.py --res 512 --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}_$v.txt 2>&1; echo $v; cut -c1-420 gpurun_out/bench_${TAG}_$v.txt
done
