#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).

usage: python tools/launch_summary.py launches.csv [> profiles/rN_launches_summary.txt]
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"^void\s+", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)
        if name.startswith("at::"):
            name = "torch:" + name.split("<")[0].split("::")[-1]
        rows.append((name, ns, r["Grid Size"], r["Block Size"]))
    tot = sum(ns for _, ns, _, _ in rows)
    agg = defaultdict(lambda: [0, 0.0, None, None])
    for name, ns, g, b in rows:
        a = agg[name]
        a[0] += 1
        a[1] += ns
        a[2], a[3] = g, b
    print(f"# {len(rows)} launches, {tot / 1e6:.2f} ms of kernel time")
    for name, (n, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:44s} launches={n:5d} total_ms={ns / 1e6:10.2f} share={ns / tot:.3f} avg_us={ns / n / 1e3:10.1f} last_grid={g} block={b}")


if __name__ == "__main__":
    main(sys.argv[1])
