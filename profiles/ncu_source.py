#!/usr/bin/env python
"""Top stalled SASS instructions of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv > file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
# first section only (SASS); header is the row starting with "Address"
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
c = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[c["# Samples"]] or 0) for r in body)
print("instructions:", len(body), "total samples:", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[c[s]] or 0) for r in body) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
ranked = sorted(enumerate(body), key=lambda ir: -int(ir[1][c["# Samples"]] or 0))[:n]
for i, r in sorted(ranked):
    top = sorted(((int(r[c[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{i:4d} {int(r[c['# Samples']]):7d} {r[c['Source']].strip()[:70]:70s} exec={r[c['Instructions Executed']]:>9s} "
          f"smemWF={r[c['L1 Wavefronts Shared']]:>9s}/{r[c['L1 Wavefronts Shared Ideal']]:>9s} {top}")
