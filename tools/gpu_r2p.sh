#!/bin/bash
# fused vs un-fused G2P / P2G (AEP_FUSED), singleton scatter build: GPU suite both ways on the parity files, benches at rest / flowing / pinned
TAG=${1:-r2p}; mkdir -p gpurun_out
for F in 1 0; do
AEP_FUSED=$F timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer.py tests/test_reference_pin.py tests/test_zzx_configs_at_size.py -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}_fused$F.txt 2>&1; echo "pytest fused=$F rc=$?" >> gpurun_out/pytest_${TAG}_fused$F.txt
tail -n 6 gpurun_out/pytest_${TAG}_fused$F.txt | cut -c1-300
AEP_FUSED=$F timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest_fused$F.txt 2>&1; cut -c1-600 gpurun_out/bench_${TAG}_rest_fused$F.txt
AEP_FUSED=$F timeout 600 python bench.py --steps 60 --warmup 5 --quick --pin-dt 1.5e-5 > gpurun_out/bench_${TAG}_pin_fused$F.txt 2>&1; cut -c1-600 gpurun_out/bench_${TAG}_pin_fused$F.txt
AEP_FUSED=$F timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}_flow_fused$F.txt 2>&1; cut -c1-600 gpurun_out/bench_${TAG}_flow_fused$F.txt
done
