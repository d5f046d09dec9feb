#!/bin/bash
# series (SVD-free sand) build: GPU suite, rest / flowing bench, ncu with source pages of the two main kernels in the flowing state
TAG=${1:-r2l}; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
tail -n 25 gpurun_out/pytest_${TAG}.txt | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_${TAG}_rest.txt 2>&1; cut -c1-900 gpurun_out/bench_${TAG}_rest.txt
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/bench_${TAG}_flow.txt 2>&1; cut -c1-900 gpurun_out/bench_${TAG}_flow.txt
timeout 600 python bench.py --steps 40 --warmup 5 --quick --pin-dt > gpurun_out/bench_${TAG}_flow_pin.txt 2>&1; cut -c1-900 gpurun_out/bench_${TAG}_flow_pin.txt
timeout 600 python bench.py --steps 40 --warmup 5 --quick --pin-dt --sort-threshold 0.1 > gpurun_out/bench_${TAG}_flow_pin_th0.1.txt 2>&1; cut -c1-900 gpurun_out/bench_${TAG}_flow_pin_th0.1.txt
for K in k_forces k_g2p2g; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 16 -c 1 -o gpurun_out/${TAG}_flow_$K -f python bench.py --steps 3 --warmup 5 --quick > gpurun_out/ncu_${TAG}_$K.log 2>&1
ncu -i gpurun_out/${TAG}_flow_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_flow_${K}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_flow_$K.ncu-rep --page source --csv > gpurun_out/${TAG}_flow_${K}_src.csv 2>/dev/null
done
ls -la gpurun_out/${TAG}_*
