#!/bin/bash
# round 2, N GPUs of one box: cross-process peer-memory exchange (parity on real peers), then the slab bench over peer memory and over NCCL
N=${1:-2}; TAG=${2:-r2f}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${TAG}_n${N}.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/peer_parity.py --res 128 --steps 24 --oracle --out gpurun_out/peer_parity_${TAG}_n${N}.json > gpurun_out/peer_parity_${TAG}_n${N}.log 2>&1; echo "peer parity rc=$?"
tail -n 3 gpurun_out/peer_parity_${TAG}_n${N}.log | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/peer_parity.py --res 128 --steps 24 --exchange nccl --out gpurun_out/peer_parity_${TAG}_n${N}_nccl.json > gpurun_out/peer_parity_${TAG}_n${N}_nccl.log 2>&1; echo "nccl path parity rc=$?"; tail -n 1 gpurun_out/peer_parity_${TAG}_n${N}_nccl.log | cut -c1-800
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/peer_parity.py --res 128 --steps 40 --adaptive --out gpurun_out/peer_parity_${TAG}_n${N}_adaptive.json > gpurun_out/peer_parity_${TAG}_n${N}_adaptive.log 2>&1; echo "peer parity adaptive rc=$?"
tail -n 1 gpurun_out/peer_parity_${TAG}_n${N}_adaptive.log | cut -c1-1200
for X in peer nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --exchange $X ${BENCH_ARGS} > gpurun_out/bench_${TAG}_n${N}_$X.json 2> gpurun_out/bench_${TAG}_n${N}_$X.err; echo "bench $X rc=$?"
  cut -c1-2500 gpurun_out/bench_${TAG}_n${N}_$X.json; tail -n 4 gpurun_out/bench_${TAG}_n${N}_$X.err | cut -c1-400
done
