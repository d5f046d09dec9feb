#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp instructions and stall samples per opcode.
usage: sass_mix.py dump.csv [top_n] [bucket]   (bucket > 0 also prints an address-ordered profile in buckets of that many instructions)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
hdr = rows[hi[0]]; body = rows[hi[0] + 1: hi[1] - 1 if len(hi) > 1 else None]
ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
def opcode(src):
    toks = src.split()
    if not toks: return None
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    return op.rstrip(";")
ins = []
for r in body:
    if len(r) <= ie or not r[ie].strip().isdigit(): continue
    op = opcode(r[ia])
    if op: ins.append((op, int(r[ie]), int(r[isamp] or 0)))
ex = collections.Counter(); sm = collections.Counter()
for op, n, s in ins:
    base = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "RED", "ATOM")) else op.split(".")[0]
    ex[base] += n; sm[base] += s
tot = sum(ex.values()); tots = sum(sm.values())
print(f"{len(ins)} SASS instructions, total warp-inst {tot:.4g}, samples {tots}")
for k, v in ex.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{k:14s} {v:14d} {100*v/tot:6.2f}%   samples {100*sm[k]/max(tots,1):6.2f}%")
N = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if N:
    for b in range(0, len(ins), N):
        chunk = ins[b:b + N]
        e = sum(c[1] for c in chunk); s = sum(c[2] for c in chunk)
        marks = collections.Counter()
        for op, n, _ in chunk:
            for m in ("LDG", "STG", "RED", "ATOM", "LDS", "STS", "MUFU", "BAR", "SHFL", "BRA", "WARPSYNC", "FFMA2", "FFMA", "FMUL", "IMAD", "MOV"):
                if op.startswith(m): marks[m] += 1; break
        print(f"[{b:5d}-{b+len(chunk):5d}) exec {100*e/tot:5.1f}%  samples {100*s/max(tots,1):5.1f}%  avg exec/inst {e/len(chunk)/1e6:7.2f}M  {dict(marks)}")
