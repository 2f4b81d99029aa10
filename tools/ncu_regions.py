#!/usr/bin/env python
"""Instruction and stall-sample shares of an `ncu --set full --import-source on` capture per REGION of a CUDA source file
(regions = line ranges named on the command line; inlined helpers count where their lines are).

usage: python tools/ncu_regions.py kernel.sass source_page.csv <mangled-name-part> <file.cu> name:lo-hi [name:lo-hi ...]
  kernel.sass      nvdisasm -g -c of the cubin the capture ran (line table)
  source_page.csv  ncu -i report.ncu-rep --page source --csv --launch-skip K --launch-count 1
Lines of other files (headers) are reported as "(other files)".
"""
import collections
import csv
import re
import sys


def main():
    sass, srccsv, func, cufile = sys.argv[1:5]
    regions = []
    for spec in sys.argv[5:]:
        name, rng = spec.rsplit(":", 1)
        lo, hi = rng.split("-")
        regions.append((int(lo), int(hi), name))
    lines = open(sass).read().split("\n")
    start = [i for i, l in enumerate(lines) if l.startswith(".text.") and func in l][0]
    addr2line, cur = {}, None
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            break
        m = re.search(r"line (\d+)", l)
        if "//## File" in l and m:
            cur = int(m.group(1)) if cufile in l else -1
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/", l)
        if m and cur is not None:
            addr2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(srccsv)))
    h = rows[1]
    ai, ii, si = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0]
    base = None
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) < len(h):
            continue
        try:
            a = int(r[ai], 16)
        except ValueError:
            continue
        if base is None:
            base = a
        ln = addr2line.get(a - base, 0)
        name = "(other files)" if ln < 0 else "(unattributed)"
        for lo, hi, nm in regions:
            if lo <= ln <= hi:
                name = nm
                break
        ie, sm = int(r[ii] or 0), int(r[si] or 0)
        agg[name][0] += ie; agg[name][1] += sm; agg[name][2] += 1
        tot[0] += ie; tot[1] += sm
    print(f"# {rows[0][1] if len(rows[0]) > 1 else func}")
    print(f"# {tot[0]} warp instructions executed, {tot[1]} stall samples, {sum(v[2] for v in agg.values())} SASS instructions")
    print(f"{'region':40s} {'instr %':>8s} {'samples %':>10s} {'SASS':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:40s} {100.0 * v[0] / max(tot[0], 1):8.1f} {100.0 * v[1] / max(tot[1], 1):10.1f} {v[2]:6d}")


if __name__ == "__main__":
    main()
