#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full ... --page raw --csv` dump: per kernel, the DRAM bytes of one launch.
usage: tools/ncu_traffic.py raw.csv particles tag > profiles/ncu_traffic.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr, units, body = rows[0], rows[1], rows[2:]
def val(r, name):
    i = hdr.index(name); v = float(r[i].replace(",", "")); u = units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
out = {"capture": sys.argv[3], "particles": int(sys.argv[2]), "kernels": {}}
for r in body:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("aep::", "")
    out["kernels"][name] = {"dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                            "duration_ms_under_ncu": float(r[hdr.index("gpu__time_duration.sum")])}
print(json.dumps(out, indent=1))
