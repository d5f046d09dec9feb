#!/bin/bash
# round-1 re-entry call: GPU parity suite, full bench line, ncu launch list, ncu --set full of the three particle kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/bench_v8_1gpu.json 2> gpurun_out/bench_v8_1gpu.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v8_launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_p2g|k_forces|k_g2p" -s 12 -c 3 -o gpurun_out/v8_full -f python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/v8_full.ncu-rep --page raw --csv > gpurun_out/v8_full_raw.csv 2>/dev/null
tail -n 6 gpurun_out/pytest_gpu.txt; cut -c1-1500 gpurun_out/bench_v8_1gpu.json; tail -n 3 gpurun_out/bench_v8_1gpu.err; tail -n 3 gpurun_out/ncu_full.log; ls -la gpurun_out
