#!/bin/bash
# usage: tools/gpu_multi.sh N [RES] [extra bench args]   -- N-GPU slab bench (one process per GPU, NCCL)
N=$1; RES=${2:-512}; shift; shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n${N}.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --res $RES "$@" > gpurun_out/bench_n${N}_r${RES}.txt 2> gpurun_out/bench_n${N}_r${RES}.err; echo "rc=$?" >> gpurun_out/bench_n${N}_r${RES}.txt
tail -c 3000 gpurun_out/bench_n${N}_r${RES}.txt; tail -n 6 gpurun_out/bench_n${N}_r${RES}.err
