#!/usr/bin/env python
"""Warp-state samples of a chain-kernel capture grouped by the mbarrier an instruction waits on
(offsets inside SmemCtl: full +0x00, empty +0x30, acc_full +0x60, act_ready +0x70) plus the rest."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[isamp] or 0) for r in data)
groups = {}
last_wait = None
for r in data:
    s = int(r[isamp] or 0)
    src = r[isrc]
    m = re.search(r"TRYWAIT.*\+0x38(0[0-9a-f]{2})\]", src)
    key = None
    if m:
        off = int(m.group(1), 16)
        key = "wait full[] (weights/enc landed)" if off < 0x30 else "wait empty[] (producer: slot free)" if off < 0x60 else \
              "wait acc_full (epilogue: MMA done)" if off < 0x70 else "wait act_ready (issuer: epilogue done)"
        last_wait = key
    elif last_wait and ("BRA" in src) and s > 0 and data.index(r) and "TRYWAIT" in data[data.index(r) - 1][isrc]:
        key = last_wait
    elif "BAR.SYNC" in src or "BSYNC" in src:
        key = "bar/bsync"
    elif "UTCHMMA" in src or "UTCBAR" in src:
        key = "mma issue"
    elif "LDTM" in src:
        key = "tcgen05.ld"
    elif "STS" in src:
        key = "st.shared"
    else:
        key = "other"
    groups[key] = groups.get(key, 0) + s
print(f"total samples {tot}")
for k, v in sorted(groups.items(), key=lambda kv: -kv[1]):
    print(f"{v:8d} {100*v/tot:5.1f}%  {k}")
