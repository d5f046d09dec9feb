#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_final.txt | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer.py tests/test_gpu_colliders.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -n 3
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/bench_final_quick.txt 2>&1; cut -c1-700 gpurun_out/bench_final_quick.txt
