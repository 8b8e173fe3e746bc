#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"# launch list summary: {path}\n")
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms total (ncu-serialised, cold cache: compare shares)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {us:.1f} | {100 * us / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
