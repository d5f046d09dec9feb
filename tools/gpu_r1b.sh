#!/bin/bash
# v2 kernels: GPU parity tests, bench line, launch list, full ncu capture of the three particle kernels
mkdir -p gpurun_out
{ nvidia-smi; nproc; lscpu | head -20; } > gpurun_out/box.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -s -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --res 512 --steps 20 --warmup 5 > gpurun_out/bench512.txt 2> gpurun_out/bench512.err; echo "bench rc=$?" >> gpurun_out/bench512.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v2.csv python bench.py --res 512 --steps 3 --warmup 3 --quick > gpurun_out/ncu_launch.log 2>&1
for k in k_forces k_g2p k_p2g; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_v2_$k python bench.py --res 512 --steps 3 --warmup 3 --quick > gpurun_out/ncu_$k.log 2>&1
done
tail -n 5 gpurun_out/smoke.txt; tail -n 25 gpurun_out/pytest_gpu.txt; head -c 4000 gpurun_out/bench512.txt; tail -n 5 gpurun_out/bench512.err; cat gpurun_out/bench_ref.txt | head -c 1500
