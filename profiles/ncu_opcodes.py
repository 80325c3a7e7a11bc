#!/usr/bin/env python
"""Executed warp instructions per opcode and per address range of an ncu source-page CSV
(ncu -i X.ncu-rep --page source --csv > file).  usage: ncu_opcodes.py file [nbins]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
c = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
thr = collections.Counter()
total = 0
for r in body:
    n = int(r[c["Instructions Executed"]] or 0)
    src = r[c["Source"]].split()
    op = src[1] if src[0].startswith("@") else src[0]
    op = op.split(".")[0]
    ops[op] += n
    thr[op] += int(r[c["Thread Instructions Executed"]] or 0)
    total += n
print("warp instructions executed:", total)
for k, v in ops.most_common(18):
    print(f"  {k:8s} {v:12d} {100*v/total:5.1f}%  avg threads {thr[k]/max(v,1):5.1f}")
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if nb:
    # contiguous regions between barriers / big branches: print executed counts per block of 200 instructions
    step = max(1, len(body) // nb)
    for i in range(0, len(body), step):
        blk = body[i:i + step]
        n = sum(int(r[c["Instructions Executed"]] or 0) for r in blk)
        s = sum(int(r[c["# Samples"]] or 0) for r in blk)
        print(f"  [{i:5d}..{i+len(blk):5d}) exec {n:12d} {100*n/total:5.1f}%  samples {s}")
