#!/bin/bash
# round 2, first GPU call: GPU suite, engine-vs-reference table, packed-gather A/B (rest + perturbed state), C1..C4 bench lines
TAG=${1:-r2a}
mkdir -p gpurun_out
PKG=$PWD/anisotropicelastoplasticity_b200
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
timeout 120 python tests/diag/gpu_refpin_report.py > gpurun_out/refpin_${TAG}.txt 2>&1
for v in default gpk; do
  AEP_B200_LIB=$PKG/libaep_b200_$v.so timeout 600 python bench.py --res 512 --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}_$v.txt 2>&1; echo $v; cut -c1-420 gpurun_out/bench_${TAG}_$v.txt
  AEP_B200_LIB=$PKG/libaep_b200_$v.so timeout 600 python bench.py --res 512 --steps 20 --warmup 5 --quick --perturb 2e-3 > gpurun_out/bench_${TAG}_${v}_pert.txt 2>&1; echo $v pert; cut -c1-420 gpurun_out/bench_${TAG}_${v}_pert.txt
done
for c in C1 C2 C3 C4; do timeout 600 python bench.py --config $c > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err; echo "$c rc=$?"; cut -c1-700 gpurun_out/bench_${TAG}_$c.json; tail -n 3 gpurun_out/bench_${TAG}_$c.err; done
tail -n 8 gpurun_out/pytest_${TAG}.txt; cat gpurun_out/refpin_${TAG}.txt | cut -c1-260
