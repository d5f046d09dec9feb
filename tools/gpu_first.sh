#!/bin/bash
# first GPU bring-up: everything goes to gpurun_out/
mkdir -p gpurun_out
{ nvidia-smi; nproc; free -g; lscpu | head -20; } > gpurun_out/box.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 900 python tools/gpu_report.py > gpurun_out/report.txt 2>&1; echo "report rc=$?" >> gpurun_out/report.txt
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --res 256 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench256.txt 2>&1; echo "bench rc=$?" >> gpurun_out/bench256.txt
tail -n 30 gpurun_out/smoke.txt gpurun_out/report.txt gpurun_out/bench256.txt
tail -n 40 gpurun_out/pytest_gpu.txt
