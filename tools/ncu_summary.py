#!/usr/bin/env python
"""Text summary of an `ncu --set full` report: the metrics DESIGN.md / profiles/README.md quote, one block per kernel launch.

usage: python tools/ncu_summary.py report.ncu-rep|raw_page.csv [--cells-per-launch N] [--json out.json] > profiles/rN_<kernel>_ncu.txt
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/shared throughput %"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("sass__inst_executed_local_loads", "local (spill) loads"), ("sass__inst_executed_local_stores", "local (spill) stores"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "mio_throttle", "math_pipe_throttle", "branch_resolving", "no_instruction",
               "not_selected", "dispatch_stall", "lg_throttle"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--cells-per-launch", type=float, default=0)
    ap.add_argument("--json")
    a = ap.parse_args()
    if a.report.endswith(".csv"):      # the raw page exported on the GPU box (`ncu -i rep --page raw --csv`): gpurun brings back at most 64 MiB
        raw = open(a.report).read()
    else:
        raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    out_json = []
    for v in rows[2:]:
        rec = dict(zip(h, v)); un = dict(zip(h, units))
        print(f"== {rec.get('Kernel Name', '?')[:100]}")
        for k, label in KEYS:
            if k in rec and rec[k] != "":
                print(f"  {label:28s} {rec[k]} {un.get(k, '')}")
        st = [(n, float(rec.get(STALLS % n, 0) or 0)) for n in STALL_NAMES]
        print("  stall cycles per issued instruction: " + ", ".join(f"{n} {x:.2f}" for n, x in sorted(st, key=lambda t: -t[1]) if x >= 0.05))
        def to_bytes(k):
            x = float(rec.get(k, 0) or 0); u = un.get(k, "")
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        d = {"kernel": rec.get("Kernel Name", ""), "dram_bytes": dram}
        if a.cells_per_launch:
            print(f"  dram bytes per cell          {dram / a.cells_per_launch:.2f}  ({a.cells_per_launch:.0f} cells per launch)")
            print(f"  warp instructions per cell   {float(rec.get('smsp__inst_executed.sum', 0)) / a.cells_per_launch:.2f}")
            d["cells_per_launch"] = a.cells_per_launch; d["dram_bytes_per_cell"] = dram / a.cells_per_launch
        out_json.append(d)
    if a.json:
        json.dump(out_json, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
