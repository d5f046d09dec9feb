#!/bin/bash
# usage: tools/gpu_round2.sh TAG  -- the round's evidence run on one B200: GPU parity suite, the full bench line, the reference arm, the
# ncu launch list of the bench command, one `ncu --set full` capture of each particle kernel (flowing state), C1..C4 bench lines
TAG=${1:-r2}; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
tail -n 6 gpurun_out/pytest_${TAG}.txt | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/bench_${TAG}_1gpu.json; tail -n 3 gpurun_out/bench_${TAG}_1gpu.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "reference rc=$?"
cut -c1-600 gpurun_out/bench_${TAG}_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/ncu_launch_${TAG}.log 2>&1
for KS in k_forces:12 k_g2p2g:12 k_p2g:6 k_force_scatter:6; do
K=${KS%%:*}; SK=${KS##*:}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$|^$K<" -s $SK -c 1 -o gpurun_out/${TAG}_flow_$K -f python bench.py --steps 3 --warmup 5 --quick > gpurun_out/ncu_${TAG}_$K.log 2>&1
ncu -i gpurun_out/${TAG}_flow_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_flow_${K}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_flow_$K.ncu-rep --page source --csv > gpurun_out/${TAG}_flow_${K}_src.csv 2>/dev/null
rm -f gpurun_out/${TAG}_flow_$K.ncu-rep
done
for c in C1 C2 C3 C4; do timeout 300 python bench.py --config $c > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err; echo "$c rc=$?"; cut -c1-500 gpurun_out/bench_${TAG}_$c.json; done
ls -la gpurun_out/${TAG}_*
