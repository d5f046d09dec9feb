#!/bin/bash
# usage: tools/gpu_r1e.sh N   -- tests + 1-GPU quick bench + N-GPU NCCL-slab bench
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for se in 1 8; do
timeout 600 python bench.py --res 512 --steps 16 --warmup 8 --quick --sort-every $se > gpurun_out/bench_v4_se$se.txt 2>&1
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --res 512 > gpurun_out/bench_n${N}.txt 2> gpurun_out/bench_n${N}.err; echo "rc=$?" >> gpurun_out/bench_n${N}.txt
tail -n 12 gpurun_out/pytest_gpu.txt; cut -c1-420 gpurun_out/bench_v4_se*.txt; tail -c 3500 gpurun_out/bench_n${N}.txt; tail -n 8 gpurun_out/bench_n${N}.err
