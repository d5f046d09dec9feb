#!/bin/bash
# split forces: parity suites + A/B against the one-kernel forces
TAG=${1:-r2t}; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer.py tests/test_gpu_slabs.py tests/test_reference_pin.py tests/test_zzx_configs_at_size.py tests/test_zzy_c1_exact_gpu.py -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
tail -n 8 gpurun_out/pytest_${TAG}.txt | cut -c1-300
for S in 1 0; do
  AEP_SPLIT_FORCES=$S timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest_split$S.txt 2>&1; echo "== split $S"; cut -c1-560 gpurun_out/bench_${TAG}_rest_split$S.txt
  AEP_SPLIT_FORCES=$S timeout 600 python bench.py --steps 60 --warmup 5 --quick --pin-dt 1.5e-5 > gpurun_out/bench_${TAG}_pin_split$S.txt 2>&1; cut -c1-560 gpurun_out/bench_${TAG}_pin_split$S.txt
done
