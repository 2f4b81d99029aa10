#!/usr/bin/env python
"""BASELINE.json configs[3] (C4): batched P(R|X) over candidate copy-number assignments on a KIR-like tandem-repeat region.

200 kbp region = 20 kbp flank + 16 x 10 kbp repeat units (0.5 % diverged copies) + 20 kbp flank, two haplotypes (0.2 % het), k = 40,
10 kbp HiFi reads (p = 0.001); mappings are generated once on the current X (generate_mappings), then B candidate X (current
copy numbers +-1 on a few repeat nodes, as neighbors.rs:239-270 produces them) are scored by to_full_prob_reads restricted to the
mappings (freq.rs:175-192 under posterior.rs:504-515).  Prints one JSON line: reads x candidates / s and GCUPS of that call.

usage: python tools/bench_c4.py [--candidates 64] [--reads 200] [--steps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from dbgphmm_b200 import graphs, synth  # noqa: E402
from dbgphmm_b200 import hmmv2 as H  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=64)
    ap.add_argument("--reads", type=int, default=200)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--read-len", type=int, default=10_000)
    a = ap.parse_args()
    t0 = time.time()
    hap = synth.tandem_repeat_genome(10_000, 16, 20_000, seed=3, divergence=0.005)
    hap2 = synth.mutate_substitutions(hap, 0.002, 77)
    sg, _ = graphs.build_dbg([hap.tobytes(), hap2.tobytes()], 40, seed=9)
    reads = synth.sample_reads([hap, hap2], 20, a.read_len, 0.001, 13)[:a.reads]
    par = H.params_uniform(0.001)
    par.n_warmup = 40
    li, lt = sg.to_probs("non_zero")
    m = H.PHMMModel(sg.src, sg.dst, sg.base, li, lt, par)
    rd = H.Reads(reads)
    print(f"graph N={sg.n_nodes} E={sg.n_edges} reads={len(reads)} built in {time.time() - t0:.1f}s", file=sys.stderr)
    t1 = time.time()
    maps = m.generate_mappings(rd, None, False)   # top-n active sets (score-ratio sets overflow the 400-entry SparseVec inside a 16-copy repeat, as in the reference)
    t_map = time.time() - t1
    rng = np.random.default_rng(1)
    B = a.candidates
    cn = sg.node_copy_num
    X = np.repeat(cn[None, :], B, 0).astype(np.uint32)
    rep = np.where(cn >= 2)[0]
    for b in range(1, B):
        idx = rng.choice(rep, size=min(40, len(rep)), replace=False)
        X[b, idx] = np.maximum(1, X[b, idx].astype(np.int64) + rng.choice([-1, 1], size=len(idx))).astype(np.uint32)
    m.set_copy_nums_batch(X, "non_zero")
    tot, per = m.to_full_prob_reads(rd, maps)   # warm-up
    ts = []
    for _ in range(a.steps):
        t2 = time.time()
        tot, per = m.to_full_prob_reads(rd, maps)
        ts.append(time.time() - t2)
    n_rows = sum(len(r) for r in reads)
    cells = float(len(maps.nodes)) * B   # forward only: |mapping.nodes(i)| per base per candidate
    dt = min(ts)
    print(json.dumps({"workload": "C4: 200 kbp tandem-repeat region, batched P(R|X) with mappings", "n_nodes": int(sg.n_nodes), "reads": len(reads),
                      "candidates": B, "rows_per_candidate": n_rows, "generate_mappings_s": round(t_map, 3), "to_full_prob_reads_ms": round(dt * 1e3, 2),
                      "reads_x_candidates_per_s": round(len(reads) * B / dt, 1), "gcups": round(cells / dt / 1e9, 3),
                      "best_candidate": int(np.argmax(tot)), "logp_current_x": float(tot[0])}))


if __name__ == "__main__":
    main()
