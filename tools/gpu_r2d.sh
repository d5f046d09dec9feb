#!/bin/bash
# round 2: ncu evidence of the fused build (launch list + one --set full capture of each particle kernel, with source pages)
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --quick ${BENCH_ARGS} > gpurun_out/ncu_launch_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_forces|k_g2p2g" -s ${SKIP:-8} -c 2 -o gpurun_out/${TAG}_full -f python bench.py --steps 3 --warmup 3 --quick ${BENCH_ARGS} > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv --kernel-name regex:k_forces > gpurun_out/${TAG}_src_k_forces.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_full.ncu-rep --page source --csv --kernel-name regex:k_g2p2g > gpurun_out/${TAG}_src_k_g2p2g.csv 2>/dev/null
tail -n 3 gpurun_out/ncu_full_${TAG}.log; ls -la gpurun_out/${TAG}_*
