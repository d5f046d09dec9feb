#!/usr/bin/env python
"""Offline (no GPU) view of a kernel's SASS loops: for every backward branch print the loop's instruction count and opcode mix.
usage: sass_loops.py lib.so kernel_substring [min_len]"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]; min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 8
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name: continue
    ins = []
    for line in f.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    mix = collections.Counter(s.split()[1].split(".")[0] if s.startswith("@") else s.split()[0].split(".")[0] for _, s in ins)
    print(f"== {name[:90]}: {len(ins)} instructions; mix {dict(mix.most_common(10))}")
    for i, (a, s) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?\S*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", s)
        if not m: continue
        t = int(m.group(1), 16)
        if t in addr_idx and addr_idx[t] <= i and i - addr_idx[t] + 1 >= min_len:
            body = ins[addr_idx[t]: i + 1]
            c = collections.Counter(x.split()[1].split(".")[0] if x.startswith("@") else x.split()[0].split(".")[0] for _, x in body)
            print(f"   loop [{addr_idx[t]:5d}-{i:5d}] {len(body):4d} instrs  {dict(c.most_common(12))}")
