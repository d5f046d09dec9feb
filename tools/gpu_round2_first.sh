#!/bin/bash
# usage (round 2, first GPU call):   make -C anisotropicelastoplasticity_b200/csrc variant-default && \
#                                    make -C anisotropicelastoplasticity_b200/csrc variant-gpk VFLAGS=-DAEP_GATHER_PK=1     (here, on the CPU)
#                                    gpurun --timeout 1500 -- 'tools/gpu_round2_first.sh r2a'
# Everything that was written after the round-1 GPU budget ran out, in one call: the whole GPU suite (no -x: every file reports), the
# engine-vs-reference error table, the packed-gather A/B (tests with the variant library, then a quick bench line per build), the
# other BASELINE configurations, and the contract bench lines.
TAG=${1:-r2a}
mkdir -p gpurun_out
PKG=$PWD/anisotropicelastoplasticity_b200
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.txt
timeout 120 python tests/diag/gpu_refpin_report.py > gpurun_out/refpin_${TAG}.txt 2>&1
timeout 300 python tests/diag/gpu_random_report.py > gpurun_out/random_${TAG}.txt 2>&1; tail -n 1 gpurun_out/random_${TAG}.txt | cut -c1-900
if [ -f $PKG/libaep_b200_gpk.so ]; then
  AEP_B200_LIB=$PKG/libaep_b200_gpk.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_pin.py -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_${TAG}_gpk.txt 2>&1; echo "pytest gpk rc=$?" >> gpurun_out/pytest_${TAG}_gpk.txt
  tools/gpu_ab2.sh ${TAG} default gpk
fi
for c in C1 C2 C3 C4; do timeout 600 python bench.py --config $c > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err; echo "$c rc=$?"; cut -c1-500 gpurun_out/bench_${TAG}_$c.json; done
timeout 900 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "reference rc=$?"
tail -n 8 gpurun_out/pytest_${TAG}.txt; tail -n 4 gpurun_out/pytest_${TAG}_gpk.txt 2>/dev/null; cat gpurun_out/refpin_${TAG}.txt | cut -c1-260; cut -c1-1500 gpurun_out/bench_${TAG}_1gpu.json; cut -c1-500 gpurun_out/bench_${TAG}_reference.json
