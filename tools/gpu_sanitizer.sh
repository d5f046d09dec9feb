#!/bin/bash
# usage: tools/gpu_sanitizer.sh  -- compute-sanitizer memcheck over the hot path on small scenes (one B200): smoke() with NVTX ranges on, one
# full-substep parity test (TMA boxes, deferred strays, graph replay), the peer-slab test (several contexts, migration)
mkdir -p gpurun_out; O=gpurun_out/sanitizer_memcheck.txt; : > $O
AEP_NVTX=1 timeout 120 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" >> $O 2>&1; echo "smoke under memcheck rc=$?" | tee -a $O
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "one_substep_parity and sand" >> $O 2>&1; echo "substep parity under memcheck rc=$?" | tee -a $O
grep -E "ERROR SUMMARY|passed|failed|smoke:" $O | cut -c1-300
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_peer.py -m gpu -q --no-header -p no:cacheprovider -k "pinned_dt and 2 or cloth" >> $O 2>&1; echo "peer slabs (particles + cloth) under memcheck rc=$?" | tee -a $O
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_colliders.py tests/test_zz_checkpoint.py -m gpu -q --no-header -p no:cacheprovider >> $O 2>&1; echo "colliders + checkpoint under memcheck rc=$?" | tee -a $O
grep -E "ERROR SUMMARY|passed|failed" $O | cut -c1-200
