"""Generates tests/golden/oracle_fullsize_<case>.npz: outputs of the CPU oracle (oracle/dbgphmm_oracle.cpp) for FULL-LENGTH reads on the
FULL-SIZE graphs the benchmark is quoted on (BASELINE.json configs[2] "C3": 1 Mbp diploid, N = 1,332,435; a configs[4] "C5" case:
5 Mbp diploid, N = 6.66 M, 20 kbp read).  One read costs minutes of CPU time and GBs of memory, which is why it is a committed fixture
and not a live oracle call inside the GPU tests.

    python tests/golden/make_fullsize_fixture.py c3 [n_reads=2]
    python tests/golden/make_fullsize_fixture.py c5 [n_reads=1]

Per read: ln P forward / backward; per DP row and direction: density, number of m/i entries and of d entries, a 64-bit hash of the
SORTED node ids of both lists (the active sets -- bit-exact quantities), the row scalar (forward e / backward mb); the non-zero node
frequencies of the read; and how many sparse rows had a tie (within 1e-9 relative) across the top-n boundary of the selection that
produced them (SURVEY.md 8c asks for that count: the reference's tie order is unpinned).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dbgphmm_b200 import graphs, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # the bench.py graph (same seeds) and reads drawn like bench.py's rank 0
    "c3": dict(genome_len=1_000_000, read_len=10_000, k=40, read_seed=1000, n_reads=2),
    "c5": dict(genome_len=5_000_000, read_len=20_000, k=40, read_seed=1000, n_reads=1),
}


def ids_hash(ids):
    """FNV-1a (64 bit) over the sorted ids as little-endian u32 -- trivially restated in the test."""
    h = np.uint64(0xcbf29ce484222325)
    p = np.uint64(0x100000001b3)
    with np.errstate(over="ignore"):
        for b in np.sort(np.asarray(ids, np.uint32)).astype("<u4").tobytes():
            h = (h ^ np.uint64(b)) * p
    return h


def build_case(name):
    c = CASES[name]
    h0 = synth.random_genome(c["genome_len"], 0)
    h1 = synth.mutate_substitutions(h0, 0.01, 1)
    g, _ = graphs.build_dbg([h0.tobytes(), h1.tobytes()], c["k"], seed=100)
    reads = synth.sample_reads([h0, h1], 1.0, c["read_len"], 0.001, c["read_seed"])
    return c, g, reads


def boundary_tie(prev_row, n_nodes, k, rel=1e-9):
    """Bit 0: the k-th and (k+1)-th largest merged values of a sparse row tie within rel on the ln scale (the criterion of
    tests/common.py::same_up_to_ties); bit 1: they are exactly equal.  The top-k set is then not determined by the values."""
    with np.errstate(divide="ignore", invalid="ignore"):
        idx = np.union1d(prev_row.ids, prev_row.ids_d)
        v = np.full(len(idx), -np.inf)
        pos = np.searchsorted(idx, prev_row.ids)
        v[pos] = np.logaddexp(prev_row.m, prev_row.i)
        posd = np.searchsorted(idx, prev_row.ids_d)
        v[posd] = np.logaddexp(v[posd], prev_row.d)
    if len(v) <= k:
        return 0
    s = np.sort(v)[::-1]
    a, b = s[k - 1], s[k]
    if a == b:
        return 3
    return 1 if abs(a - b) <= rel * max(1.0, abs(a)) else 0


def one_read(o, read, n_nodes, n_active):
    t0 = time.time()
    f = o.forward_sparse(read, False)
    b = o.backward_sparse(read)
    n = len(read)
    out = {"logp_fwd": f.full_prob(), "logp_bwd": b.full_prob()}
    fr = O.PHMMOutput(f, b).to_node_freqs()
    nz = np.nonzero(fr > 0)[0]
    out["freq_idx"] = nz.astype(np.uint32); out["freq_val"] = fr[nz]
    ties = [0, 0]
    for d, t in enumerate((f, b)):
        dense = np.zeros(n, np.uint8); n_mi = np.zeros(n, np.uint32); n_d = np.zeros(n, np.uint32)
        h_mi = np.zeros(n, np.uint64); h_d = np.zeros(n, np.uint64); sc = np.zeros(n); tie = np.zeros(n, np.uint8)
        prev = None
        order = range(n) if d == 0 else range(n - 1, -1, -1)   # the order the rows were computed in
        for r in order:
            row = t.row(r)
            dense[r] = row.is_dense
            sc[r] = row.e if d == 0 else row.mb
            if not row.is_dense:
                n_mi[r] = len(row.ids); n_d[r] = len(row.ids_d)
                h_mi[r] = ids_hash(row.ids); h_d[r] = ids_hash(row.ids_d)
                if prev is not None and not prev.is_dense:
                    tie[r] = boundary_tie(prev, n_nodes, n_active)
                    ties[d] += int(tie[r] != 0)
            prev = row if not row.is_dense else None
        p = "f_" if d == 0 else "b_"
        out.update({p + "dense": dense, p + "n_mi": n_mi, p + "n_d": n_d, p + "h_mi": h_mi, p + "h_d": h_d, p + "scalar": sc, p + "tie": tie})
    out["ties"] = np.array(ties, np.uint32)
    out["seconds"] = time.time() - t0
    return out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    c, g, reads = build_case(name)
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else c["n_reads"]
    par = O.params_uniform(0.001)
    par.n_warmup = c["k"]
    li, lt = g.to_probs("normal")
    o = O.PHMMModel(g.src, g.dst, g.base, li, lt, par)
    N = g.n_nodes
    print(name, "N =", N, "E =", g.n_edges, flush=True)
    data = {"n_nodes": np.array([N], np.uint64), "n_edges": np.array([g.n_edges], np.uint64), "n_reads": np.array([n_reads], np.uint32),
            "genome_len": np.array([c["genome_len"]]), "read_len": np.array([c["read_len"]]), "k": np.array([c["k"]]), "read_seed": np.array([c["read_seed"]])}
    for i in range(n_reads):
        r = one_read(o, reads[i], N, int(par.n_active_nodes))
        print(f" read {i}: len {len(reads[i])} ln P fwd {r['logp_fwd']:.9f} bwd {r['logp_bwd']:.9f} ties at the top-n boundary (fwd, bwd) {r['ties'].tolist()} "
              f"freq entries {len(r['freq_idx'])} in {r['seconds']:.0f} s", flush=True)
        data[f"r{i}_read"] = np.asarray(reads[i], np.uint8)
        for kx, v in r.items():
            data[f"r{i}_{kx}"] = np.asarray(v)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"oracle_fullsize_{name}.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
