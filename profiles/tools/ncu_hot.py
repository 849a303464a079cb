#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output:
per-opcode instruction totals, the hottest instructions by issue count and by stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:                       # first captured launch only
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        data.append(r)
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print("instructions executed (warp-level):", tot_inst, " samples:", tot_samp)
byop = collections.Counter(); sop = collections.Counter()
for r in data:
    op = r[ix["Source"]].split()[0]
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    byop[op] += int(r[ix["Instructions Executed"]]); sop[op] += int(r[ix["# Samples"]])
print("\nby opcode (inst share, sample share):")
for op, c in byop.most_common(25):
    print(f"  {op:10s} {100*c/tot_inst:5.1f}%  {100*sop[op]/max(tot_samp,1):5.1f}%")
stall_cols = [h for h in hdr if h.startswith("stall_")] if any(h.startswith("stall_") for h in hdr) else []
print("\ntop 40 instructions by samples:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:40]:
    print(f"  {r[ix['# Samples']]:>7s} {r[ix['Instructions Executed']]:>10s}  {r[ix['Source']].strip()[:90]}")
if len(sys.argv) > 2:
    print("\nall columns:", hdr)
