#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --res 512 --steps 40 --warmup 8 --quick > gpurun_out/bench_v7_1gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 8 --res 512 > gpurun_out/bench_v7_n${N}.txt 2> gpurun_out/bench_v7_n${N}.err; echo "rc=$?" >> gpurun_out/bench_v7_n${N}.txt
tail -n 8 gpurun_out/pytest_gpu.txt; cut -c1-420 gpurun_out/bench_v7_1gpu.txt; tail -c 2600 gpurun_out/bench_v7_n${N}.txt | cut -c1-3000; tail -n 5 gpurun_out/bench_v7_n${N}.err
