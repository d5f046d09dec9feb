#!/usr/bin/env python
"""Top-N SASS instructions of an `ncu --page source --csv` dump by stall samples, in address order, with the stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
hdr = rows[hi[0]]; body = [r for r in rows[hi[0]+1:(hi[1]-1 if len(hi) > 1 else None)] if len(r) > 10]
ia = hdr.index("Source"); isamp = hdr.index("# Samples")
cols = {c: hdr.index(c) for c in hdr if c.startswith("stall_") and "Not Issued" not in c}
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot)
top = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:N]
for i in sorted(top):
    r = body[i]
    st = {k[6:]: int(r[v] or 0) for k, v in cols.items() if int(r[v] or 0) > 0.1 * int(r[isamp])}
    print(f"{i:5d} {r[ia][:64]:64s} {100*int(r[isamp])/tot:5.1f}% {st}")
