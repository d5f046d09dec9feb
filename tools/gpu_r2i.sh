#!/bin/bash
TAG=${1:-r2i}; mkdir -p gpurun_out; O=gpurun_out/peer_debug2_${TAG}.txt; : > $O
timeout 400 python tools/peer_debug2.py --res 128 --steps 3 >> $O 2>&1
grep -v Warning $O | cut -c1-3000
