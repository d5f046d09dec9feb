#!/bin/bash
# stand-in fix + new tests (colliders, dt rule at scale) + sort-threshold sweep at a pinned dt
TAG=${1:-r2m}; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_colliders.py tests/test_gpu_dt_rule_at_scale.py tests/test_gpu_peer.py tests/test_gpu_parity.py -m gpu -q --no-header -rf -s -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
grep -E "substeps per frame|passed|failed|Error|assert" gpurun_out/pytest_${TAG}.txt | cut -c1-600 | tail -n 30
timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest.txt 2>&1; cut -c1-700 gpurun_out/bench_${TAG}_rest.txt
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}_flow.txt 2>&1; cut -c1-900 gpurun_out/bench_${TAG}_flow.txt
for TH in 0.05 0.1 0.2 0.5; do
timeout 600 python bench.py --steps 60 --warmup 5 --quick --pin-dt 1.5e-5 --sort-threshold $TH > gpurun_out/bench_${TAG}_pin_th$TH.txt 2>&1; echo "TH $TH"; cut -c1-900 gpurun_out/bench_${TAG}_pin_th$TH.txt
done
