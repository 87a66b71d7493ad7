#!/usr/bin/env python
"""Markdown table (+ profiles/traffic.json entries) of the per-configuration ncu captures written by tools/gpu_round.sh.
usage: python tools/ncu_table.py <raw.csv> <cases.json> [--traffic profiles/traffic.json]
raw.csv = `ncu -i prof_cases.ncu-rep --page raw --csv`; cases.json = the launch order tools/profile_cases.py wrote.
Every case must be found (two launches each, the second -- warm -- one is reported): a capture that lost kernels fails loudly
instead of producing a shorter table."""
import csv
import json
import sys

raw, cases_path = sys.argv[1], sys.argv[2]
traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
rows = list(csv.reader(open(raw)))
ix = {n: i for i, n in enumerate(rows[0])}
data = [r for r in rows[2:] if len(r) == len(rows[0])]
cases = json.load(open(cases_path))
stalls = [m for m in rows[0] if "issue_stalled" in m and m.endswith("_per_issue_active.ratio") and "not_issued" not in m]


def is_ours(r):
    n = r[ix["Kernel Name"]]
    return any(k in n for k in ("y2r_", "r2y_", "tone_map_kernel", "scale_"))


def short(n):
    n = n.replace("abfast::", "").replace("ab::", "").replace("unsigned short", "u16").replace("unsigned char", "u8").replace("(int)", "").replace("(bool)", "")
    return n.replace("void ", "").split("(Ab")[0]


ours = [r for r in data if is_ours(r)]
print("| case | kernel | frames / launch | time us | us / frame | dram rd MB | dram wr MB | dram / alg bytes | alg GB/s (under ncu) | issue active % | achieved occ % | regs | grid | warp inst x32 / px | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
pos = 0
traffic = {}
for c in cases:
    match = c.get("match")
    # the case's launches: the next two rows (reformat kernels are captured 1:1); helper cases (tone map, scaling) name their kernel
    if match:
        while pos < len(ours) and match not in ours[pos][ix["Kernel Name"]]:
            pos += 1
        grp = []
        limit = int(c.get("launches", 0))  # helper cases that share a kernel name (the two scaling cases) say how many launches are theirs
        while pos < len(ours) and match in ours[pos][ix["Kernel Name"]] and (not limit or len(grp) < limit):
            grp.append(ours[pos])
            pos += 1
        if not grp:
            sys.exit(f"ncu_table: no kernel matching {match!r} captured for case {c['case']}")
        r = grp[-1]
    else:
        n = int(c.get("launches", 2))  # kernels the case launched (two calls); a case served by one fused launch per call has 2
        if pos + n > len(ours):
            sys.exit(f"ncu_table: capture ends before case {c['case']} ({len(ours)} kernels, needed {pos + n})")
        r = ours[pos + n - 1]  # the last launch of the warm second call
        if n != 2:
            c = dict(c, frames=max(1, c["frames"] * 2 // n))  # unfused path: each launch covers frames * 2 / n frames
        pos += n
    g = lambda m: float(r[ix[m]].replace(",", ""))
    px = c["px_per_frame"] * c["frames"]
    alg = c["alg_bytes_per_px"] * px
    TU = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
    BU = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    t_us = g("gpu__time_duration.sum") * TU[rows[1][ix["gpu__time_duration.sum"]]]
    rd = g("dram__bytes_read.sum") * BU[rows[1][ix["dram__bytes_read.sum"]]]
    wr = g("dram__bytes_write.sum") * BU[rows[1][ix["dram__bytes_write.sum"]]]
    top = sorted(((g(m), m.split("issue_stalled_")[1].split("_per_")[0]) for m in stalls), reverse=True)[:3]
    print("| %s | %s | %d | %.1f | %.1f | %.1f | %.1f | %.2f | %.0f | %.0f | %.0f | %d | %d | %.1f | %s |" % (
        c["case"], short(r[ix["Kernel Name"]]), c["frames"], t_us, t_us / c["frames"], rd / 1e6, wr / 1e6, (rd + wr) / alg, alg / (t_us * 1e-6) / 1e9,
        g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        g("launch__registers_per_thread"), g("launch__grid_size"), g("smsp__inst_executed.sum") * 32 / px,
        ", ".join("%s %.2f" % (b, a) for a, b in top)))
    traffic[c["case"]] = {"dram_bytes_per_frame": (rd + wr) / c["frames"], "dram_read_per_frame": rd / c["frames"], "dram_write_per_frame": wr / c["frames"],
                          "frames_per_launch": c["frames"], "alg_bytes_per_frame": alg / c["frames"],
                          "source": "ncu --set full of the batched launch bench.py times (tools/profile_cases.py), dram__bytes_read.sum + dram__bytes_write.sum / frames"}
if traffic_path:
    json.dump(traffic, open(traffic_path, "w"), indent=1)
