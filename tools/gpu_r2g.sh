#!/bin/bash
# bisect the peer-parity failure: same-device (serialised) at the failing size; then the new deferral build: GPU suite + flowing bench
TAG=${1:-r2g}
mkdir -p gpurun_out
for RES in 128 96; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_parity.py --same-device --res $RES --steps 24 --out gpurun_out/peer_parity_${TAG}_samedev_$RES.json > gpurun_out/peer_parity_${TAG}_samedev_$RES.log 2>&1; echo "peer same-device res $RES rc=$?"
cut -c1-900 gpurun_out/peer_parity_${TAG}_samedev_$RES.json
done
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
tail -n 15 gpurun_out/pytest_${TAG}.txt | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}_flow.txt 2>&1; cut -c1-1800 gpurun_out/bench_${TAG}_flow.txt
timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest.txt 2>&1; cut -c1-1200 gpurun_out/bench_${TAG}_rest.txt
