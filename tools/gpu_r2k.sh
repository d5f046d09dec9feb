#!/bin/bash
# deferral build: full GPU suite, ncu of rest + flowing, sort-threshold sweep on the flowing state
TAG=${1:-r2k}; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
tail -n 25 gpurun_out/pytest_${TAG}.txt | cut -c1-300
for ST in rest flowing; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_forces|k_g2p2g" -s 16 -c 4 -o gpurun_out/${TAG}_${ST}_full -f python bench.py --steps 3 --warmup 5 --quick --state $ST > gpurun_out/ncu_full_${TAG}_${ST}.log 2>&1
ncu -i gpurun_out/${TAG}_${ST}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_${ST}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_${ST}_full.ncu-rep --page source --csv --kernel-name regex:"k_forces<8" > gpurun_out/${TAG}_${ST}_src_k_forces.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_${ST}_full.ncu-rep --page source --csv --kernel-name regex:"k_g2p2g<8" > gpurun_out/${TAG}_${ST}_src_k_g2p2g.csv 2>/dev/null
rm -f gpurun_out/${TAG}_${ST}_full.ncu-rep
done
for TH in 0.05 0.1 0.2; do
timeout 600 python bench.py --steps 40 --warmup 5 --quick --sort-threshold $TH > gpurun_out/bench_${TAG}_flow_th$TH.txt 2>&1; cut -c1-900 gpurun_out/bench_${TAG}_flow_th$TH.txt
done
timeout 600 python bench.py --steps 40 --warmup 5 --quick > gpurun_out/bench_${TAG}_flow_default.txt 2>&1; cut -c1-900 gpurun_out/bench_${TAG}_flow_default.txt
