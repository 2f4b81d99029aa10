"""Shared helpers of the parity tests: build the same model in the oracle and on the GPU, compare rows."""
import json
import os

import numpy as np

from dbgphmm_b200 import graphs
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
REL_TOL = 1e-9  # north_star: log P(R|X) and node frequencies within 1e-9 relative in f64


def kat():
    with open(os.path.join(HERE, "golden", "reference_kat.json")) as f:
        return json.load(f)


def oracle_params(p, n_warmup=None, n_active=None, warmup_threshold=None):
    q = O.params_uniform(p)
    if warmup_threshold is not None:
        q.warmup_threshold = warmup_threshold
    if n_warmup is not None:
        q.n_warmup = n_warmup
    if n_active is not None:
        q.n_active_nodes = n_active
    return q


def to_gpu_params(q):
    from dbgphmm_b200 import hmmv2 as H
    g = H.Params()
    for name, _ in H.Params._fields_:
        setattr(g, name, getattr(q, name))
    return g


def oracle_model(sg, param, mode="normal", probs=None):
    li, lt = probs if probs is not None else sg.to_probs(mode)
    return O.PHMMModel(sg.src, sg.dst, sg.base, li, lt, param)


def gpu_model(sg, param, mode="normal", probs=None, **kw):
    from dbgphmm_b200 import hmmv2 as H
    li, lt = probs if probs is not None else sg.to_probs(mode)
    return H.PHMMModel(sg.src, sg.dst, sg.base, li, lt, to_gpu_params(param), **kw)


def close_log(a, b, rel=REL_TOL, abs_floor=1e-9):
    """|a-b| <= rel * max(1, |b|) elementwise on natural logs, -inf == -inf."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    both_inf = np.isneginf(a) & np.isneginf(b)
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= np.maximum(abs_floor, rel * np.abs(b))
    return ok | both_inf


def assert_rows_match(g, o, n_nodes, what=""):
    """g: GPU row, o: oracle row (both natural logs).  Index sets and order must be identical; values within REL_TOL.

    The GPU stores one exponent per cell, so a state more than 2^-1022 below the largest state of the SAME node in the
    same row is flushed to zero; such cells are accepted when the oracle value is that small."""
    assert g.is_dense == o.is_dense, f"{what}: density differs"
    for name in ("mb", "ib", "e"):
        assert close_log(getattr(g, name), getattr(o, name)).all(), f"{what}: {name} {getattr(g, name)} vs {getattr(o, name)}"
    if o.is_dense:
        cellmax = np.maximum(np.maximum(o.m, o.i), o.d)
        for name in ("m", "i", "d"):
            a, b = getattr(g, name), getattr(o, name)
            ok = close_log(a, b) | (np.isneginf(a) & (b < cellmax - 700.0))
            assert ok.all(), f"{what}: dense {name} mismatch at {np.nonzero(~ok)[0][:5]}: {a[~ok][:5]} vs {b[~ok][:5]}"
    else:
        # Index SETS must be identical.  The order inside a row follows the order of the selected top nodes, which
        # is only reproducible up to values that tie within rounding (log-space f64 on the CPU vs exponent-extended
        # linear f64 on the GPU round differently) — see same_up_to_ties for the check on the selections themselves.
        assert sorted(g.ids) == sorted(o.ids), f"{what}: m/i node set differs\n gpu {sorted(map(int, g.ids))}\n ref {sorted(map(int, o.ids))}"
        assert sorted(g.ids_d) == sorted(o.ids_d), f"{what}: d node set differs\n gpu {sorted(map(int, g.ids_d))}\n ref {sorted(map(int, o.ids_d))}"
        g = _reorder_like(g, o)
        # largest state of each node in this row (the GPU cell exponent follows it)
        cmax = {}
        for k, a_, b_ in zip(o.ids, o.m, o.i):
            cmax[int(k)] = max(a_, b_)
        for k, d_ in zip(o.ids_d, o.d):
            cmax[int(k)] = max(cmax.get(int(k), -np.inf), d_)
        for name, ids in (("m", o.ids), ("i", o.ids), ("d", o.ids_d)):
            a_, b_ = getattr(g, name), getattr(o, name)
            ok = close_log(a_, b_)
            for j in np.nonzero(~ok)[0]:
                assert np.isneginf(a_[j]) and b_[j] < cmax[int(ids[j])] - 700.0, \
                    f"{what}: sparse {name}[{j}] node {ids[j]}: {a_[j]} vs {b_[j]}"


class _R:
    pass


def _reorder_like(g, o):
    """view of sparse row g with its entries permuted into o's order"""
    r = _R()
    pos = {int(k): j for j, k in enumerate(g.ids)}
    idx = [pos[int(k)] for k in o.ids]
    r.ids = o.ids; r.m = g.m[idx]; r.i = g.i[idx]
    posd = {int(k): j for j, k in enumerate(g.ids_d)}
    idxd = [posd[int(k)] for k in o.ids_d]
    r.ids_d = o.ids_d; r.d = g.d[idxd]
    r.is_dense = False; r.mb, r.ib, r.e = g.mb, g.ib, g.e
    return r


def same_up_to_ties(a_ids, b_ids, b_vals, rel=1e-9):
    """Two descending selections are equal up to permutations inside groups of values that tie within `rel`."""
    a_ids = [int(x) for x in a_ids]; b_ids = [int(x) for x in b_ids]
    if len(a_ids) != len(b_ids):
        return False
    n = len(b_ids)
    start = 0
    for i in range(1, n + 1):
        tie = False
        if i < n:
            x, y = b_vals[i - 1], b_vals[i]
            tie = (np.isneginf(x) and np.isneginf(y)) or abs(x - y) <= rel * max(1.0, abs(x))
        if not tie:
            if sorted(a_ids[start:i]) != sorted(b_ids[start:i]):
                return False
            start = i
    return True


def row_order_exact(g, o):
    return o.is_dense or (list(g.ids) == list(o.ids) and list(g.ids_d) == list(o.ids_d))


def assert_tables_match(gt, ot, n_nodes, what=""):
    assert len(gt) == len(ot), f"{what}: number of rows"
    for r in range(len(ot)):
        assert_rows_match(gt.row(r), ot.row(r), n_nodes, f"{what} row {r}")


def random_linear_graph(n, seed):
    rng = np.random.default_rng(seed)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].tobytes()
    return graphs.genome_graph_to_seq_graph([(seq, 1)]), seq
