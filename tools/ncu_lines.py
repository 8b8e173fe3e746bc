#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of one kernel in an .ncu-rep.
The report gives per-SASS-instruction counts; the line table comes from `nvdisasm --print-line-info` on the
cubin of the matching object file (a -lineinfo build), joined by instruction order.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep ddnerf_b200/build/sampler.o [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    hdr = {h: i for i, h in enumerate(rows[1])}
    data = [(r[hdr["Source"]].strip(), int(r[hdr["Instructions Executed"]] or 0), int(r[hdr["# Samples"]] or 0)) for r in rows[2:]]
    return name, data


def line_table(obj, kernel_key):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    lines, cur, active = [], None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            active = kernel_key(m.group(1))
            continue
        if not active:
            continue
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    return lines


def main():
    rep, obj = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    name, data = sass_rows(rep)
    # mangled-name match: function base name and the integer template arguments in order
    base = re.search(r"(\w+)<", name).group(1) if "<" in name else name.split("(")[0].split("::")[-1]
    targs = re.findall(r"\((?:int|bool)\)(\d+)", name)
    mang = "".join(f"Li{t}E" if True else "" for t in targs)

    def key(sym):
        if base not in sym:
            return False
        got = re.findall(r"L[ib](\d+)E", sym)
        return got == targs
    table = line_table(obj, key)
    if len(table) != len(data):
        print(f"warning: {len(table)} instructions in the object, {len(data)} in the report (different build?)", file=sys.stderr)
    src = {}
    path = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc")
    agg = {}
    for (sass, n, smp), loc in zip(data, table):
        a = agg.setdefault(loc, [0, 0])
        a[0] += n
        a[1] += smp
    tot = sum(a[0] for a in agg.values())
    tots = sum(a[1] for a in agg.values())
    print(f"{name[:100]}\ntotal warp instructions {tot}, samples {tots}")
    for loc, (n, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if loc:
            f = os.path.join(path, loc[0])
            if f not in src and os.path.exists(f):
                src[f] = open(f).read().splitlines()
            if f in src and loc[1] - 1 < len(src[f]):
                text = src[f][loc[1] - 1].strip()
        print(f"{100 * n / max(tot, 1):5.1f}% inst {100 * smp / max(tots, 1):5.1f}% smp  {loc}  {text[:100]}")


if __name__ == "__main__":
    main()
