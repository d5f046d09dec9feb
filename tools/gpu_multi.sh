#!/bin/bash
# usage: tools/gpu_multi.sh N TAG [extra bench args]   -- N-GPU slab bench over peer memory (one process per GPU)
N=$1; TAG=${2:-r2}; shift; shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${TAG}_n${N}.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err; echo "rc=$?"
cut -c1-3500 gpurun_out/bench_${TAG}_n${N}.json; grep -v "^$" gpurun_out/bench_${TAG}_n${N}.err | grep -v OMP_NUM | tail -n 6 | cut -c1-300
