#!/usr/bin/env python
"""Split the SASS of one captured kernel at its BAR.SYNC instructions and report, per region, the executed
instructions, the stall samples and their main reasons (input: ncu --page source --csv --print-source sass)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    if len(r) == len(hdr): data.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
regions = []; cur = dict(n=0, inst=0, samp=0, st=collections.Counter(), first=None, ldg=0, top=[])
for r in data:
    src = r[ix["Source"]].strip()
    cur["n"] += 1; cur["inst"] += int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]]); cur["samp"] += s
    if cur["first"] is None: cur["first"] = r[ix["Address"]]
    for h in stalls:
        v = r[ix[h]]
        if v and v != "0": cur["st"][h] += int(v)
    if "LDG" in src: cur["ldg"] += int(r[ix["Instructions Executed"]])
    cur["top"].append((s, src))
    if "BAR.SYNC" in src:
        regions.append(cur); cur = dict(n=0, inst=0, samp=0, st=collections.Counter(), first=None, ldg=0, top=[])
regions.append(cur)
ti = sum(x["inst"] for x in regions); ts = sum(x["samp"] for x in regions)
print(f"total warp instructions {ti}, samples {ts}")
for i, x in enumerate(regions):
    if not x["n"]: continue
    top = ", ".join(f"{k[6:]} {100*v/max(1,x['samp']):.0f}%" for k, v in x["st"].most_common(4))
    print(f"region {i:2d} @{x['first'][-5:]} sass {x['n']:5d}  inst {100*x['inst']/ti:5.1f}%  samples {100*x['samp']/ts:5.1f}%  LDG {x['ldg']:>9d}  | {top}")
    if len(sys.argv) > 2:
        for s, src in sorted(x["top"], reverse=True)[:int(sys.argv[2])]:
            print(f"        {s:6d}  {src[:100]}")
