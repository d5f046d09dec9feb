#!/bin/bash
# round 2: correctness of the fused / TMA / peer-memory build + quick bench lines
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
tail -n 40 gpurun_out/pytest_${TAG}.txt
if true; then
  timeout 600 python bench.py --res 512 --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}.txt 2>&1; cut -c1-600 gpurun_out/bench_${TAG}.txt
  timeout 600 python bench.py --res 512 --steps 20 --warmup 5 --quick --perturb 2e-3 > gpurun_out/bench_${TAG}_pert.txt 2>&1; cut -c1-600 gpurun_out/bench_${TAG}_pert.txt
  for c in C1 C2 C3 C4; do timeout 300 python bench.py --config $c > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err; echo "$c rc=$?"; cut -c1-600 gpurun_out/bench_${TAG}_$c.json; tail -n 2 gpurun_out/bench_${TAG}_$c.err; done
fi
