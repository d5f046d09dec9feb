"""Static SASS opcode mix of one kernel in a built library: sass_static_mix.py <lib.so> <kernel-name-substring>"""
import sys, re, collections, subprocess
so=sys.argv[1]; kern=sys.argv[2]
txt=subprocess.run(["cuobjdump","-sass",so],capture_output=True,text=True).stdout
m=re.search(r"Function : (\S*%s\S*)\n(.*?)(?=\n\s*Function :|\Z)"%kern, txt, re.S)
body=m.group(2)
ops=collections.Counter(); n=0
for line in body.split("\n"):
    mm=re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if mm:
        op=mm.group(2); base=op.split(".")[0]
        if base in ("LDS","LDG","STS","STG","RED","ATOMS","LDL","STL"): base=".".join(op.split(".")[:2]) if base in("LDS","LDG") else base
        ops[base]+=1; n+=1
print(m.group(1)[:40], "total", n)
print({k:v for k,v in ops.most_common(22)})
