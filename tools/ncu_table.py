#!/usr/bin/env python
"""Markdown table of the per-configuration ncu captures written by tools/gpu_round.sh (tools/profile_cases.py order).
usage: ncu -i gpurun_out/<tag>/prof_cases.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_table.py raw.csv"""
import csv
import sys

CASES = [("NS 8K 444 10b->RGBA10", 7680 * 4320), ("C2 fwd 4K 444 10b+A->RGBA10 premult", 3840 * 2160),
         ("C2 back RGBA10 premult->YUVA", 3840 * 2160), ("C3 8K 420 12b bilinear+unmul", 7680 * 4320),
         ("C3 back RGBA12->8K 420 12b+A premult", 7680 * 4320), ("C5 4K 420 10b->RGB16 bilinear", 3840 * 2160),
         ("C5 4K 420 10b->RGB16 nearest", 3840 * 2160), ("C1 1080p 420 8b->RGBA8 bilinear", 1920 * 1080),
         ("C4 cell fwd (1080p 420 8b+A bilinear)", 1920 * 1080), ("C4 cell back RGBA8->420 8b+A", 1920 * 1080)]
rows = list(csv.reader(open(sys.argv[1])))
ix = {n: i for i, n in enumerate(rows[0])}
data = rows[2:]
stalls = [m for m in rows[0] if "issue_stalled" in m and m.endswith("_per_issue_active.ratio") and "not_issued" not in m]
print("| case | kernel | time us | dram rd MB | dram wr MB | issue active % | achieved occ % | regs | grid | warp inst x32 / px | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for k, (name, px) in enumerate(CASES):
    r = data[2 * k + 1]  # the warm second launch of each pair
    g = lambda m: float(r[ix[m]].replace(",", ""))
    kn = r[ix["Kernel Name"]].replace("abfast::", "").replace("unsigned short", "u16").replace("unsigned char", "u8").replace("(int)", "").replace("(bool)", "")
    kn = kn.replace("void ", "").split("(Ab")[0]
    top = sorted(((g(m), m.split("issue_stalled_")[1].split("_per_")[0]) for m in stalls), reverse=True)[:3]
    print("| %s | %s | %.1f | %.1f | %.1f | %.0f | %.0f | %d | %d | %.1f | %s |" % (
        name, kn, g("gpu__time_duration.sum"), g("dram__bytes_read.sum"), g("dram__bytes_write.sum"),
        g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        g("launch__registers_per_thread"), g("launch__grid_size"), g("smsp__inst_executed.sum") * 32 / px,
        ", ".join("%s %.2f" % (b, a) for a, b in top)))
