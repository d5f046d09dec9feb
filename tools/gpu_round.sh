#!/bin/bash
# usage: tools/gpu_round.sh TAG  -- the round's evidence run on one B200: GPU parity suite, the full bench line, the ncu launch list
# of the bench command and one `ncu --set full` capture of each particle kernel (read here, summarised under profiles/)
TAG=${1:-dev}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
timeout 900 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_launch_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_p2g|k_forces|k_g2p" -s 12 -c 3 -o gpurun_out/${TAG}_full -f python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
tail -n 6 gpurun_out/pytest_${TAG}.txt; cut -c1-1800 gpurun_out/bench_${TAG}_1gpu.json; tail -n 3 gpurun_out/bench_${TAG}_1gpu.err; cut -c1-600 gpurun_out/bench_${TAG}_reference.json; tail -n 2 gpurun_out/ncu_full_${TAG}.log
