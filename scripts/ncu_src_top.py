#!/usr/bin/env python
"""Top stall locations of an ncu source-page CSV (SASS view): usage ncu_src_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; data = rows[2:]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
idx = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[isamp]))[:n]:
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{idx[id(r)]:5d} {int(r[isamp]):7d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>10s} {r[isrc].strip()[:70]:70s} {st}")
