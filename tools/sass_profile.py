#!/usr/bin/env python
"""Walk an `ncu --page source --csv` dump in address order and print, per bucket of N instructions, the share of executed
warp instructions and of stall samples plus the landmark opcodes inside (to map hot regions back to phases of a kernel)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ins = [(r[ia], int(r[ie] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > ie and r[ia].strip()]
te = sum(i[1] for i in ins); ts = sum(i[2] for i in ins)
print(f"{len(ins)} SASS instructions, {te:.4g} executed, {ts} samples")
for b in range(0, len(ins), N):
    chunk = ins[b:b + N]
    e = sum(c[1] for c in chunk); s = sum(c[2] for c in chunk)
    marks = collections.Counter()
    for src, n, _ in chunk:
        toks = src.split(); op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        for m in ("LDG", "STG", "RED", "ATOM", "LDS", "STS", "MUFU", "BAR", "SHFL", "BRA", "WARPSYNC"):
            if op.startswith(m): marks[m] += 1
    print(f"[{b:5d}-{b+len(chunk):5d}) exec {100*e/te:5.1f}%  samples {100*s/ts:5.1f}%  avg exec/inst {e/len(chunk)/1e6:7.2f}M  {dict(marks)}")
