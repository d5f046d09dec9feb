#!/bin/bash
# usage: tools/gpu_sanity.sh  -- a short check of a build on one B200: smoke(), the parity / peer / collider suites, one quick bench line per state
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke_final.txt | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer.py tests/test_gpu_colliders.py tests/test_reference_pin.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -n 3
timeout 600 python bench.py --steps 20 --warmup 5 --quick --state rest > gpurun_out/bench_final_quick_rest.txt 2>&1; cut -c1-560 gpurun_out/bench_final_quick_rest.txt
timeout 600 python bench.py --steps 60 --warmup 5 --quick --pin-dt 1.5e-5 > gpurun_out/bench_final_quick_pin.txt 2>&1; cut -c1-560 gpurun_out/bench_final_quick_pin.txt
