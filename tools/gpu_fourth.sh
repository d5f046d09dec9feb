#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slabs.py -m gpu -q --no-header -rf -s -p no:cacheprovider > gpurun_out/pytest_slabs.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_slabs.txt
for k in k_forces k_g2p k_p2g; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof2_$k python bench.py --res 512 --steps 3 --warmup 3 --quick > gpurun_out/ncu2_$k.log 2>&1
done
tail -n 30 gpurun_out/pytest_slabs.txt
