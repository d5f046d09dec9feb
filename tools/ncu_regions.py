#!/usr/bin/env python
"""Executed warp instructions of an `ncu --page source --csv` dump split into regions of equal execution count (= straight-line
stretches / loop bodies), with the opcode mix and stall samples of each.  usage: ncu_regions.py src.csv [particles]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1]))); npart = float(sys.argv[2]) if len(sys.argv) > 2 else 64520064.0
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
hdr = rows[hi[0]]; body = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else None)] if len(r) > 10 and r[0] != "Address"]
iS, iE, iT, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
iW = hdr.index("L1 Wavefronts Shared")
tot = sum(int(r[iE] or 0) for r in body); tots = sum(int(r[iN] or 0) for r in body)
print(f"instructions {len(body)}, executed warp-instr {tot:.3e} = {tot/npart:.1f} per particle, samples {tots}")
def op(s):
    t = s.split(); t = t[1] if t[0].startswith("@") else t[0]; return t.split(".")[0]
# regions: consecutive instructions whose execution counts are within 2% of each other
regs = []; cur = None
for i, r in enumerate(body):
    e = int(r[iE] or 0)
    if cur and abs(e - cur["e0"]) <= 0.02 * max(cur["e0"], 1): cur["rows"].append(i)
    else: cur = {"e0": e, "rows": [i]}; regs.append(cur)
for g in regs:
    ex = sum(int(body[i][iE] or 0) for i in g["rows"])
    if ex < 0.004 * tot: continue
    sm = sum(int(body[i][iN] or 0) for i in g["rows"]); wf = sum(int(body[i][iW] or 0) for i in g["rows"])
    mix = collections.Counter(); 
    for i in g["rows"]: mix[op(body[i][iS])] += 1
    print(f"[{g['rows'][0]:5d}-{g['rows'][-1]:5d}] n={len(g['rows']):4d} exec/instr={g['e0']/npart*32:6.2f}/particle-lane  share={100*ex/tot:5.1f}%  warp-instr/particle={ex/npart:6.2f} samples={100*sm/tots:5.1f}% smem-wavefronts/particle={wf/npart:5.2f}  {dict(mix.most_common(8))}")
