#!/usr/bin/env python
"""Histogram of the Blackwell-specific SASS opcodes per kernel of ddnerf_b200/libddnerf_b200.so
(`cuobjdump -sass`; runs without a GPU).  Writes profiles/sass_opcodes.md:

    python tools/sass_opcodes.py

UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA engine, no
tensor map), UTMALDG would be cp.async.bulk.tensor (not used: operands are pre-swizzled images moved as plain bytes),
SYNCS = mbarrier operations, MUFU = SFU transcendental, RED/ATOM = global reductions."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "ddnerf_b200", "libddnerf_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "REDUX", "RED", "ATOM", "SHFL", "HMMA", "FFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w in ("RED", "ATOM") and op.startswith(w)):
                    cur[w] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    merged = collections.OrderedDict()       # template instances of the per-ray kernels are merged by kernel name
    for (name, c), dn in zip(kernels.items(), demangle):
        short = dn.replace("(anonymous namespace)::", "").replace("ddnerf::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", short)
        base = short if "mlp_tc" in short or "selftest" in short else re.sub(r"<.*", "<...>", short)
        agg = merged.setdefault(base, [0, collections.Counter()])
        agg[0] += 1
        agg[1].update(c)
    rows = [(f"{b} x{n}" if n > 1 else b, c) for b, (n, c) in merged.items()]
    path = os.path.join(REPO, "profiles", "sass_opcodes.md")
    with open(path, "w") as f:
        f.write("# SASS opcode histogram of `ddnerf_b200/libddnerf_b200.so` (sm_100a)\n\n")
        f.write("`python tools/sass_opcodes.py` = `cuobjdump -sass` of the shipped library, instructions counted per kernel.\n"
                "UTCHMMA = `tcgen05.mma`, UTCBAR = `tcgen05.commit`, LDTM = `tcgen05.ld`, UBLKCP = `cp.async.bulk` (TMA engine, "
                "plain byte ranges of pre-swizzled operand images; the single-CTA kernels and the dW kernel need no tensor map; the CTA-pair chain kernels fetch their ring stages with tiled TMA = UTMALDG.3D.2CTA so that both CTAs signal the leader's barrier), SYNCS = mbarrier, "
                "MUFU = SFU, RED/ATOM = global reductions.\n\n")
        f.write("| kernel | instr | " + " | ".join(WATCH) + " |\n|---|---:|" + "---:|" * len(WATCH) + "\n")
        for short, c in rows:
            f.write(f"| `{short}` | {c['total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |\n")
        tot = collections.Counter()
        for _, c in rows:
            tot.update(c)
        f.write(f"| **all {len(kernels)} kernels** | {tot['total']} | " + " | ".join(str(tot[w]) if tot[w] else "" for w in WATCH) + " |\n")
    print("wrote", path, "-", len(kernels), "kernels;", {w: tot[w] for w in ("UTCHMMA", "LDTM", "UBLKCP", "UTMALDG")})


if __name__ == "__main__":
    sys.exit(main())
