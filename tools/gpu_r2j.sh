#!/bin/bash
TAG=${1:-r2j}; mkdir -p gpurun_out; O=gpurun_out/peer_debug3_${TAG}.txt; : > $O
timeout 300 python tools/peer_debug3.py 128 >> $O 2>&1
timeout 300 python tools/peer_debug3.py 64 >> $O 2>&1
grep -v Warning $O | cut -c1-600
