#!/usr/bin/env python
"""Attributes executed instructions of one profiled kernel to CUDA source lines.

  ncu -i rep.ncu-rep --page source --csv --kernel-id :::N > k.csv          (per-SASS-instruction execution counts)
  cuobjdump -xelf all libavif_b200/csrc/<tu>.o ; nvdisasm -g <tu>.sm_100a.cubin > tu.sass   (SASS with //## File/line marks)
  python tools/sass_lines.py k.csv tu.sass <pixels> [top]

The object must be the build that was profiled (same instruction offsets).  Prints warp-instructions per pixel by source
line ("inlined at" chains are kept: the innermost line is used) and by opcode for the heaviest lines.
"""
import collections
import csv
import re
import sys


def main():
    kcsv, sass, px = sys.argv[1], sys.argv[2], float(sys.argv[3])
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(kcsv)))
    kname = rows[0][1]
    hdr = rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    insts, seen = [], set()
    for r in rows[2:]:
        if len(r) <= ix["Instructions Executed"] or not r[ix["Instructions Executed"]].isdigit():
            continue
        addr = int(r[ix["Address"]], 16)
        if addr in seen:  # the page repeats the table once per view: keep the first copy
            continue
        seen.add(addr)
        insts.append((addr, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]])))
    base = insts[0][0]
    # split the disassembly into functions, keep (offset -> line) for the one whose opcodes match
    funcs, cur, line = {}, None, None
    for l in open(sass):
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            line = None
            continue
        if cur is None:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            line = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip(), line))
    want = [re.sub(r"\s+", " ", s.split(";")[0]) for _, s, _ in insts]
    best = None
    for name, body in funcs.items():
        if len(body) != len(insts):
            continue
        same = sum(1 for (o, t, _), w in zip(body, want) if t.split()[0].lstrip("@!P0123456789T ").split(".")[0] in w)
        if best is None or same > best[0]:
            best = (same, name, body)
    if best is None:
        sys.exit("no function with %d instructions in %s" % (len(insts), sass))
    _, name, body = best
    by_line = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    total = 0
    for (off, txt, ln), (addr, src, n) in zip(body, insts):
        by_line[ln] += n
        op = src.split()[1] if src.startswith("@") else src.split()[0]
        ops[ln][op.rstrip(";")] += n
        total += n
    print("kernel", kname[:120])
    print("function", name[:100], "instructions", len(insts), "warp-inst/px %.2f" % (total * 32 / px))
    for ln, n in by_line.most_common(top):
        o = " ".join("%s=%.2f" % (k, v * 32 / px) for k, v in ops[ln].most_common(5))
        print("%-22s %6.2f /px   %s" % ("%s:%d" % ln if ln else "?", n * 32 / px, o))


if __name__ == "__main__":
    main()
