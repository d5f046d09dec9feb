#!/bin/bash
# usage: tools/gpu_c4_slabs.sh N  -- BASELINE configs[3] (cloth-sand coupling, 4.04 M particles + 256 x 256 cloth, 256^3) on N GPUs over peer
# memory: parity against one whole-domain context, with particles and cloth points changing owner, and the time per substep
N=$1; mkdir -p gpurun_out
for DT in 1e-6; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/peer_parity.py --scene c4 --steps 40 --dt $DT --out gpurun_out/peer_parity_c4_n$N.json > gpurun_out/peer_parity_c4_n$N.log 2>&1; rc=$?; echo "c4 parity N=$N dt=$DT rc=$rc"
grep -v "^$" gpurun_out/peer_parity_c4_n$N.log | grep -E "AepError|\"ok\"" | tail -n 2 | cut -c1-1800
[ $rc = 0 ] && break
done
