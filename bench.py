#!/usr/bin/env python
"""bench.py — PHMM forward-backward GCUPS on the 1 Mbp diploid config (BASELINE.json configs[2], "C3").

One step = one pass of the hot path (run_sparse: forward_sparse + backward_sparse + node frequencies, freq.rs:51-55,
245-255) over one batch of synthetic reads per GPU.  A cell = one (read base, graph node) pair evaluated by
f_step / b_step; dense warm-up rows count all N nodes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads-per-gpu R] [--genome-len L]

N > 1 is launched by torchrun (one rank per GPU); reads are sharded over ranks (weak scaling: every rank processes
its own R reads), node frequencies and the summed ln P(R) are combined with one NCCL all-reduce per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_CELL = 48  # read previous M,I,D + write new M,I,D as f64 (our cells add a 4 B exponent: 56 B)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-gpu", type=int, default=1184, help="reads per GPU per step (1184 = 148 SMs x 8 resident sparse jobs: one wave)")
    ap.add_argument("--genome-len", type=int, default=1_000_000)
    ap.add_argument("--read-len", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=40)
    ap.add_argument("--cpu-sample-reads", type=int, default=0, help="reads of the CPU baseline sample (0 = auto)")
    ap.add_argument("--cpu-read-len", type=int, default=8, help="bases kept per read in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end passes (profiling runs under ncu only; the line then has e2e.value = null)")
    return ap.parse_args()


def workload_tag(args):
    """BASELINE.json configs[2] (C3, the bench default) / one GPU's shard of configs[4] (C5: 5 Mbp haplotypes, 20 kbp reads)."""
    if (args.genome_len, args.read_len, args.k) == (1_000_000, 10_000, 40):
        return "C3"
    if (args.genome_len, args.read_len, args.k) == (5_000_000, 20_000, 40):
        return "C5 (one GPU's read shard)"
    return "custom"


def make_inputs(args, rank, n_reads):
    """Same graph on every rank (seed 0); rank-specific reads.  C3: 1 Mbp diploid, 1 % het, 10 kbp HiFi reads, k = 40."""
    from dbgphmm_b200 import graphs, synth
    h0 = synth.random_genome(args.genome_len, 0)
    h1 = synth.mutate_substitutions(h0, 0.01, 1)
    g, _ = graphs.build_dbg([h0.tobytes(), h1.tobytes()], args.k, seed=100)
    cov = n_reads * args.read_len / (2.0 * args.genome_len)
    reads = synth.sample_reads([h0, h1], cov, args.read_len, 0.001, 1000 + rank)[:n_reads]
    while len(reads) < n_reads:
        reads += synth.sample_reads([h0, h1], cov, args.read_len, 0.001, 5000 + rank + len(reads))[:n_reads - len(reads)]
    li, lt = g.to_probs("normal")
    return g, li, lt, reads


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(args, g, li, lt, reads, n_sample):
    """The reference's CPU algorithm (oracle port, OpenMP over reads like rayon over reads, freq.rs:181-191) on a sample."""
    from oracle import oracle as O
    par = O.params_uniform(0.001)
    par.n_warmup = args.k
    o = O.PHMMModel(g.src, g.dst, g.base, li, lt, par)
    threads = min(O.max_threads(), n_sample)
    # Bounded sample: one full 10 kbp read costs ~100 CPU-seconds at N = 1.3 M, so the sample keeps the first
    # `cpu_read_len` bases of each read.  Every row is then a dense warm-up row (len < n_warmup), i.e. the row type that
    # makes up > 98 % of the cells of the full workload, at the same N; GCUPS is per cell, so the rate is comparable.
    sample = [r[:args.cpu_read_len] for r in reads[:n_sample]]
    cells = sum(len(r) for r in sample) * 2 * g.n_nodes
    t0 = time.perf_counter()
    fr, lf, lb = o.run_node_freqs(O.Reads(sample), "sparse", True, None, n_threads=threads, want_freqs=True)
    dt = time.perf_counter() - t0
    # sparse cells of the same sample (cheap second pass is avoided: count from the GPU-side accounting when available)
    return {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
            "sample": f"{len(sample)} reads of the same workload cut to their first {args.cpu_read_len} bases "
                      f"(all rows dense warm-up rows over N={g.n_nodes} nodes; run_sparse + to_node_freqs)",
            "seconds": dt, "logp_fwd": lf.tolist()}


def auto_cpu_sample(n_nodes, k):
    from oracle import oracle as O
    cores = O.max_threads()
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 32 << 30
    per_read = 2 * 16 * n_nodes * 24 * 2.5  # stored dense rows of both directions + temporaries
    return int(max(1, min(cores, avail * 0.5 / per_read)))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        g, li, lt, reads = make_inputs(args, 0, max(args.cpu_sample_reads, 64))
        n_sample = args.cpu_sample_reads or auto_cpu_sample(g.n_nodes, args.k)
        vals = []
        for s in range(args.warmup + args.steps):
            r = cpu_reference(args, g, li, lt, reads, n_sample)
            if s >= args.warmup:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals])); sec = float(np.mean([r["seconds"] for r in vals]))
        cb = {k: vals[-1][k] for k in ("unit", "cores", "kind", "sample")}
        cb["value"] = v
        print(json.dumps({"metric": "PHMM forward-backward GCUPS", "value": v, "unit": "GCUPS", "impl": "reference", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"{workload_tag(args)}: {args.genome_len} bp diploid 1% het, {args.read_len} bp HiFi reads p=0.001, k={args.k}, run_sparse + node freqs",
                                     "n_nodes": int(g.n_nodes), "reads_per_step": n_sample},
                          "cpu_baseline": cb, "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    from dbgphmm_b200 import hmmv2 as H
    from dbgphmm_b200.dist import allreduce_results
    for attempt in range(5):   # a context of a process that has just exited may still be tearing down (exclusive-process boxes)
        try:
            torch.cuda.set_device(local)
            torch.zeros(1, device="cuda")
            break
        except RuntimeError:
            if attempt == 4:
                raise
            time.sleep(2.0)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R = args.reads_per_gpu
    g, li, lt, reads = make_inputs(args, rank, R)
    par = H.params_uniform(0.001)
    par.n_warmup = args.k  # MultiDbg::to_phmm (multi_dbg.rs:1395)
    model = H.PHMMModel(g.src, g.dst, g.base, li, lt, par, device=local)
    N = g.n_nodes
    freqs = torch.zeros(N, dtype=torch.float64, device="cuda")
    logp = torch.zeros(1, dtype=torch.float64, device="cuda")
    lp_dev = torch.zeros(R, dtype=torch.float64, device="cuda")
    rd = H.Reads(reads)
    model.reads_to_device(rd)
    h2d = int(rd.total_bases() + rd.offsets.nbytes); d2h = int(N * 8 + 2 * R * 8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        """inputs resident in HBM; device outputs; one all-reduce of node freqs + summed ln P."""
        freqs.zero_()
        cells = model.run_node_freqs_dev(rd, "sparse", freqs.data_ptr(), logp_fwd_ptr=lp_dev.data_ptr())
        logp[0] = lp_dev.sum()
        ms = H.last_timing()[3]
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        allreduce_results(freqs, logp, dist if world > 1 else None)   # one NCCL all-reduce of [N] freqs + summed ln P
        ev1.record(); torch.cuda.synchronize()
        return sum(cells), ms + ev0.elapsed_time(ev1)

    def step_e2e():
        """host buffers in, host buffers out through the public API (reads handle re-created: validation + H2D inside)."""
        t0 = time.perf_counter()
        r2 = H.Reads(reads)
        fr, lf, lb, cells = model.run_node_freqs(r2, "sparse")
        if world > 1:
            t = torch.from_numpy(fr).cuda(); dist.all_reduce(t); fr = t.cpu().numpy()
        torch.cuda.synchronize()
        return sum(cells), (time.perf_counter() - t0) * 1e3, float(lf.sum())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local); sampler.start()
    H.launch_count(reset=True)
    barrier()
    tot_cells, tot_ms, k_ms, k_launch, k_cells = 0, 0.0, 0.0, 0, 0
    for _ in range(args.steps):
        c, ms = step_resident()
        tot_cells += c; tot_ms += ms
        a, b, cc = H.last_dense_kernel(); k_ms += a; k_launch += b; k_cells += cc
    barrier()
    launches = H.launch_count()
    clocks = sampler.stop()
    # max over ranks of the device time, sum over ranks of the cells
    t = torch.tensor([tot_ms], dtype=torch.float64, device="cuda"); c = torch.tensor([float(tot_cells)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c)
    ms_step = float(t.item()) / args.steps
    value = float(c.item()) / (float(t.item()) * 1e-3) / 1e9
    # end to end
    e_cells, e_ms = 0, 0.0
    for i in range(0 if args.no_e2e else 1 + min(args.steps, 2)):
        cc, ms, _ = step_e2e()
        if i > 0:
            e_cells += cc; e_ms += ms
    te = torch.tensor([e_ms], dtype=torch.float64, device="cuda"); ce = torch.tensor([float(e_cells)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX); dist.all_reduce(ce)
    e2e_val = float(ce.item()) / (float(te.item()) * 1e-3) / 1e9 if te.item() > 0 else None
    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = (k_cells * ALGO_BYTES_PER_CELL / (k_ms * 1e-3) / 1e9) if k_ms > 0 else 0.0
        # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture of this command
        # (profiles/r1_dense_pair_dram.json, tools/ncu_summary.py); scaled per cell when the launch shape differs
        traffic = None
        try:
            caps = json.load(open(os.path.join(ROOT, "profiles", "r1_dense_pair_dram.json")))
            per_cell = sum(c["dram_bytes_per_cell"] for c in caps) / len(caps)
            traffic = per_cell * (k_cells / max(k_launch, 1))
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "k_dense_fwd2 / k_dense_bwd2 (two DP rows per launch)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": "profiles/r1_dense_pair_dram.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per cell x cells per launch)", "peak_source": "MEASURED_PEAKS.json (burst copy)" if peaks else "fallback B200_PROFILING.md",
                "avg_launch_ms": k_ms / max(k_launch, 1), "launches": k_launch, "cells_per_launch": k_cells / max(k_launch, 1),
                "algorithmic_bytes_per_cell": ALGO_BYTES_PER_CELL, "kernel_share_of_step": k_ms / max(tot_ms, 1e-9)}
        out = {"metric": "PHMM forward-backward GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "reads_per_s": world * R / (ms_step * 1e-3),
               "config": {"workload": f"{workload_tag(args)}: {args.genome_len} bp diploid 1% het, {args.read_len} bp HiFi reads p=0.001, k={args.k}, run_sparse + node freqs",
                          "n_nodes": int(N), "reads_per_gpu_per_step": R, "n_active_nodes": 40, "n_warmup": args.k,
                          "l2": "DP rows of one step exceed L2 (each dense row is 28 B x N, hundreds of rows in flight)"},
               "roofline": roof,
               "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": int(launches), "clocks": clocks}
        if not args.no_cpu_baseline:
            n_sample = args.cpu_sample_reads or auto_cpu_sample(N, args.k)
            cb = cpu_reference(args, g, li, lt, reads, min(n_sample, len(reads)))
            cb.pop("logp_fwd", None); cb.pop("seconds", None)
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    try:
        rc = main()
    except Exception:
        # single-process runs get one more attempt after a transient failure (e.g. the device still being released by the
        # process that ran before this one); a multi-rank job cannot re-enter its rendezvous and fails as it is
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            raise
        import traceback
        traceback.print_exc()
        time.sleep(5.0)
        rc = main()
    sys.exit(rc)
