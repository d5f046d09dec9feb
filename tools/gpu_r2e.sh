#!/bin/bash
# round 2: configs at size (C2/C3/C4 parity), cross-process peer exchange on one GPU, flowing-state bench
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zzx_configs_at_size.py -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}_atsize.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}_atsize.txt
tail -n 30 gpurun_out/pytest_${TAG}_atsize.txt | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_parity.py --same-device --res 64 --steps 16 --oracle --out gpurun_out/peer_parity_${TAG}_samedev.json > gpurun_out/peer_parity_${TAG}_samedev.log 2>&1; echo "peer same-device rc=$?"
tail -n 5 gpurun_out/peer_parity_${TAG}_samedev.log | cut -c1-1500
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}_flow.txt 2>&1; cut -c1-1800 gpurun_out/bench_${TAG}_flow.txt
timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest.txt 2>&1; cut -c1-1200 gpurun_out/bench_${TAG}_rest.txt
