"""Generates tests/golden/oracle_c2_small.json: outputs of the CPU oracle (oracle/dbgphmm_oracle.cpp) on a seeded, scaled-down C2
case (BASELINE.json configs[1]: haploid genome, HiFi-like reads, sparse mode with n_active_nodes = 40, n_warmup = k).

    python tests/golden/make_oracle_fixture.py

The reference itself cannot run here (Rust nightly + crates.io), so this fixture pins the ORACLE, which is in turn
pinned on the reference's own known answers (reference_kat.json).  It gives the GPU parity tests a committed target that does not
depend on the oracle being rebuilt on the GPU box, and the CPU suite a regression check of the oracle itself.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dbgphmm_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASE = dict(genome_len=3000, k=24, coverage=4, read_len=400, p_err=0.002, ploidy=1, het=0.0, seed=21, n_reads=5)
ROWS = lambda n, k: sorted({k, k + 1, k + 7, n // 2, n - 2, n - 1})  # noqa: E731  forward rows whose active sets are stored


def build():
    w = synth.make_workload("c2_small", CASE["genome_len"], CASE["k"], CASE["coverage"], CASE["read_len"], CASE["p_err"],
                            ploidy=CASE["ploidy"], het=CASE["het"], seed=CASE["seed"], n_reads=CASE["n_reads"])
    par = O.params_uniform(0.001)
    par.n_warmup = w.k
    li, lt = w.graph.to_probs("normal")
    return w, par, O.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, li, lt, par)


def main():
    w, par, o = build()
    N = w.graph.n_nodes
    fr, lf, lb = o.run_node_freqs(O.Reads(w.reads), "sparse")
    out = {"_comment": "oracle outputs; regenerate with tests/golden/make_oracle_fixture.py", "case": CASE, "n_nodes": int(N),
           "n_edges": int(w.graph.n_edges), "read_lens": [int(len(r)) for r in w.reads],
           "logp_forward": [float(x) for x in lf], "logp_backward": [float(x) for x in lb],
           "node_freq_sum": float(fr.sum()), "node_freq_top": [[int(i), float(fr[i])] for i in np.argsort(-fr, kind="stable")[:40]],
           "rows": []}
    for ri, read in enumerate(w.reads[:2]):
        f, b = o.forward_sparse(read, False), o.backward_sparse(read)
        n = len(read)
        for r in ROWS(n, w.k):
            fr_, br_ = f.row(r), b.row(n - 1 - r)
            out["rows"].append({"read": ri, "row": int(r), "fwd_is_dense": bool(fr_.is_dense), "fwd_e": float(fr_.e),
                                "fwd_ids": [] if fr_.is_dense else sorted(int(x) for x in fr_.ids),
                                "fwd_ids_d": [] if fr_.is_dense else sorted(int(x) for x in fr_.ids_d),
                                "bwd_row": int(n - 1 - r), "bwd_is_dense": bool(br_.is_dense), "bwd_mb": float(br_.mb),
                                "bwd_ids": [] if br_.is_dense else sorted(int(x) for x in br_.ids),
                                "bwd_ids_d": [] if br_.is_dense else sorted(int(x) for x in br_.ids_d)})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_c2_small.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(path, "N =", N, "reads", out["read_lens"])


if __name__ == "__main__":
    main()
