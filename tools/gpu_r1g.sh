#!/bin/bash
mkdir -p gpurun_out
AEP_B200_LIB=$PWD/anisotropicelastoplasticity_b200/libaep_b200_notile.so timeout 600 python bench.py --res 512 --steps 16 --warmup 8 --quick --sort-every 1 > gpurun_out/bench_v5_notile.txt 2>&1
for k in k_forces k_g2p; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_v5_$k python bench.py --res 512 --steps 3 --warmup 3 --quick --sort-every 1 > gpurun_out/ncu_$k.log 2>&1
done
cut -c1-420 gpurun_out/bench_v5_notile.txt
