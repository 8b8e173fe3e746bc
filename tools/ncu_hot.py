#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i rep --page source --csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[isamp] or 0) for r in data)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total samples", tot)
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][isamp] or 0))[:n]:
    st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
    print(f"{idx:5d} {int(r[isamp]):7d} {100*int(r[isamp])/tot:5.1f}%  exec {r[iexec]:>9s}  {r[isrc].strip()[:70]:70s} {st}")
