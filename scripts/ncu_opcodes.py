"""Executed-instruction and stall-sample histogram by SASS opcode from `ncu -i REPORT --page source --csv`.
usage: ncu -i r.ncu-rep --page source --csv > src.csv; python scripts/ncu_opcodes.py src.csv"""
import csv, sys, collections, re
f = sys.argv[1]
rows = list(csv.reader(open(f)))
# find header row
hi = next(i for i,r in enumerate(rows) if r and r[0]=="Address")
h = rows[hi]
ix = {n:i for i,n in enumerate(h)}
ops = collections.Counter(); samples = collections.Counter()
tot=0; tots=0
for r in rows[hi+1:]:
    if len(r) < len(h): continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src[:10]
    base = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    ops[base]+=n; samples[base]+=s; tot+=n; tots+=s
print("total inst", tot, "samples", tots)
for k,v in ops.most_common(40):
    print(f"{k:12s} {v:>14d} {100*v/tot:6.2f}%  samples {100*samples[k]/tots:6.2f}%")
