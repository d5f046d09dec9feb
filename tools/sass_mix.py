#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp instructions and stall samples per opcode."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
ex = collections.Counter(); sm = collections.Counter(); tot = 0; tots = 0
for r in rows[2:]:
    if len(r) <= ie: continue
    toks = r[ia].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.rstrip(";")
    base = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "RED", "ATOM")) else op.split(".")[0]
    n = int(r[ie] or 0); s = int(r[isamp] or 0)
    ex[base] += n; sm[base] += s; tot += n; tots += s
print(f"total warp-inst {tot:.4g}, samples {tots}")
for k, v in ex.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{k:14s} {v:14d} {100*v/tot:6.2f}%   samples {100*sm[k]/max(tots,1):6.2f}%")
