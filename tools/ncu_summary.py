#!/usr/bin/env python
"""Key metrics of one kernel from an .ncu-rep (ncu --set full) as markdown: duration, DRAM traffic,
tensor-pipe activity, issue activity, shared-memory wavefronts, top stall instructions.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, row = raw[0], raw[1], raw[2]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary: {rep}\n")
    print(f"kernel: `{row[col['Kernel Name']]}`\n")
    print("| metric | value | unit |\n|---|---:|---|")
    for h in hdr:
        if any(h == k or h.endswith("." + k) for k in KEYS) and "TriageCompute" not in h or "hmma_cycles_active_realtime.avg" in h:
            print(f"| {h} | {row[col[h]]} | {units[col[h]]} |")
    src = page(rep, "source")
    h2 = src[1]
    isrc, isamp = h2.index("Source"), h2.index("# Samples")
    stall = [(i, h) for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
    data = src[2:]
    tot = sum(int(r[isamp] or 0) for r in data)
    print(f"\n## top instructions by warp-state samples (total {tot})\n")
    print("| samples | share | SASS | top stalls |\n|---:|---:|---|---|")
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:14]:
        st = sorted(((int(r[i] or 0), h) for i, h in stall), reverse=True)[:2]
        print(f"| {r[isamp]} | {100 * int(r[isamp]) / max(tot, 1):.1f}% | `{r[isrc].strip()[:60]}` | {st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]} |")


if __name__ == "__main__":
    main()
