#!/usr/bin/env python
"""profiles/ncu_traffic.json from `ncu --set full ... --page raw --csv` dumps: per kernel, the DRAM bytes of one launch (the longest
launch of each kernel name: the list passes over deferred particles share the name of their main kernel).
usage: tools/ncu_traffic.py particles tag raw1.csv [raw2.csv ...] > profiles/ncu_traffic.json"""
import csv, json, re, sys
out = {"capture": sys.argv[2], "particles": int(sys.argv[1]), "kernels": {}}
for path in sys.argv[3:]:
    rows = list(csv.reader(open(path))); hdr, units, body = rows[0], rows[1], rows[2:]
    def val(r, name):
        i = hdr.index(name); v = float(r[i].replace(",", "")); u = units[i]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u]
    for r in body:
        full = r[hdr.index("Kernel Name")]
        name = re.sub(r"<.*", "", full.replace("void ", "").replace("aep::", "").split("(")[0]).strip()
        ms = val(r, "gpu__time_duration.sum")
        if name in out["kernels"] and out["kernels"][name]["duration_ms_under_ncu"] >= ms:
            continue
        out["kernels"][name] = {"dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                                "duration_ms_under_ncu": ms, "launch": full.split("(")[0].replace("void ", "")}
print(json.dumps(out, indent=1))
