#!/bin/bash
# N GPUs of one box: free-running dt rule over peer-memory slabs to one frame boundary, against the whole-domain context (bulk statistics)
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/peer_parity.py --res 128 --steps 40 --adaptive ${PARITY_ARGS} --out gpurun_out/peer_parity_${TAG}_n${N}_adaptive.json > gpurun_out/peer_parity_${TAG}_n${N}_adaptive.log 2>&1; echo "peer parity adaptive rc=$?"
tail -n 1 gpurun_out/peer_parity_${TAG}_n${N}_adaptive.log | cut -c1-1500
