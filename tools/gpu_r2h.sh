#!/bin/bash
TAG=${1:-r2h}; mkdir -p gpurun_out; O=gpurun_out/peer_debug_${TAG}.txt; : > $O
timeout 200 python tools/peer_debug.py --res 128 --tag base >> $O 2>&1
AEP_FORCE_ROUNDS=1 timeout 200 python tools/peer_debug.py --res 128 --tag rounds1 >> $O 2>&1
timeout 200 python tools/peer_debug.py --res 128 --graph 0 --tag nograph >> $O 2>&1
timeout 200 python tools/peer_debug.py --res 128 --sort-every 1000 --tag nosort >> $O 2>&1
timeout 200 python tools/peer_debug.py --res 128 --vy 0 --tag novy >> $O 2>&1
AEP_FORCE_ROUNDS=2 timeout 200 python tools/peer_debug.py --res 64 --tag res64_rounds2 >> $O 2>&1
AEP_FORCE_ROUNDS=8 timeout 200 python tools/peer_debug.py --res 64 --tag res64_rounds8 >> $O 2>&1
grep -v Warning $O | cut -c1-1600
