"""Value-level parity at the sizes the headline is quoted on: full-length reads on the full-size C3 graph (N = 1,332,435) against the
committed oracle fixture tests/golden/oracle_fullsize_c3.npz (written by tests/golden/make_fullsize_fixture.py: one read costs ~4 CPU
minutes in the oracle).  Active sets of every DP row through the single-read tables (bit-exact quantities, compared by hash), and
ln P / node frequencies through the bulk call in the STREAM strategy (the path bench.py times)."""
import os

import numpy as np
import pytest

from dbgphmm_b200 import graphs, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def ids_hash(ids):
    """64-bit FNV-1a over the sorted ids as little-endian u32 (the fixture's h_mi / h_d); Python ints: ~4x faster than numpy scalars"""
    h = 0xcbf29ce484222325
    for b in np.sort(np.asarray(ids, np.uint32)).astype("<u4").tobytes():
        h = ((h ^ b) * 0x100000001b3) & 0xffffffffffffffff
    return np.uint64(h)


def load(name):
    path = os.path.join(HERE, "golden", f"oracle_fullsize_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path)


def build(fx):
    h0 = synth.random_genome(int(fx["genome_len"][0]), 0)
    h1 = synth.mutate_substitutions(h0, 0.01, 1)
    g, _ = graphs.build_dbg([h0.tobytes(), h1.tobytes()], int(fx["k"][0]), seed=100)
    assert g.n_nodes == int(fx["n_nodes"][0]) and g.n_edges == int(fx["n_edges"][0])
    return g


def model_of(H, g, k):
    li, lt = g.to_probs("normal")
    par = H.params_uniform(0.001); par.n_warmup = k
    return H.PHMMModel(g.src, g.dst, g.base, li, lt, par)


def close(a, b, rel=1e-9):
    return abs(a - b) <= rel * max(1.0, abs(b))


def compare_rows(tables, fx, pre, n, what):
    """-> (rows whose active sets differ, of which explained by a tie of the selection that produced them or of an earlier row)"""
    diff, diff_tied, scalar_bad = [], 0, []
    order = range(n) if pre.endswith("f_") else range(n - 1, -1, -1)
    tie_seen = False
    for r in order:
        row = tables.row(r)
        assert row.is_dense == bool(fx[pre + "dense"][r]), f"{what} row {r}: density"
        sc = row.e if pre.endswith("f_") else row.mb
        if not close(sc, float(fx[pre + "scalar"][r])):
            scalar_bad.append((r, sc, float(fx[pre + "scalar"][r])))
        if row.is_dense:
            continue
        tie_seen = tie_seen or fx[pre + "tie"][r] != 0
        same = (len(row.ids) == int(fx[pre + "n_mi"][r]) and len(row.ids_d) == int(fx[pre + "n_d"][r])
                and ids_hash(row.ids) == fx[pre + "h_mi"][r] and ids_hash(row.ids_d) == fx[pre + "h_d"][r])
        if not same:
            diff.append(r)
            diff_tied += int(tie_seen)
    return diff, diff_tied, scalar_bad


def test_c3_fullsize_rows_and_bulk_against_the_oracle_fixture(monkeypatch, capsys):
    from dbgphmm_b200 import hmmv2 as H
    fx = load("c3")
    g = build(fx)
    m = model_of(H, g, int(fx["k"][0]))
    N = g.n_nodes
    reads = [np.asarray(fx[f"r{i}_read"], np.uint8) for i in range(int(fx["n_reads"][0]))]
    report = []
    for i, read in enumerate(reads):
        n = len(read)
        f = m.forward_sparse(read, False)
        assert close(f.full_prob(), float(fx[f"r{i}_logp_fwd"]))
        df, df_t, sf = compare_rows(f, fx, f"r{i}_f_", n, f"read {i} fwd")
        del f
        b = m.backward_sparse(read)
        assert close(b.full_prob(), float(fx[f"r{i}_logp_bwd"]))
        db, db_t, sb = compare_rows(b, fx, f"r{i}_b_", n, f"read {i} bwd")
        del b
        ties = fx[f"r{i}_ties"].tolist()
        report.append(f"read {i} ({n} rows): active sets differ in {len(df)} forward / {len(db)} backward rows; the oracle saw a tie across the top-n boundary "
                      f"(1e-9 on the ln scale) in {ties[0]} / {ties[1]} rows")
        assert not sf and not sb, (sf[:3], sb[:3])
        # a set may differ only where the selection was not determined by the values (a tie at the top-n boundary, there or upstream)
        assert len(df) == df_t and len(db) == db_t, (df[:5], db[:5])
        assert len(df) <= ties[0] and len(db) <= ties[1]
    with capsys.disabled():
        print("\n" + "\n".join(report))
    # ---- the bulk call in the stream strategy (what bench.py runs), all fixture reads at once
    monkeypatch.setenv("DBGPHMM_STRATEGY", "stream")
    fr, lf, lb, cells = m.run_node_freqs(H.Reads(reads), "sparse")
    want = np.zeros(N)
    for i in range(len(reads)):
        assert close(lf[i], float(fx[f"r{i}_logp_fwd"])) and close(lb[i], float(fx[f"r{i}_logp_bwd"]))
        np.add.at(want, fx[f"r{i}_freq_idx"], fx[f"r{i}_freq_val"])
    assert np.allclose(fr, want, rtol=1e-9, atol=1e-12), np.abs(fr - want).max()
    top = np.argsort(-want, kind="stable")[:100]
    assert np.allclose(fr[top], want[top], rtol=1e-9, atol=0)
    # the same reads in the store strategy give the same answer (strategy independence at full size)
    monkeypatch.setenv("DBGPHMM_STRATEGY", "store")
    fr2, lf2, lb2, cells2 = m.run_node_freqs(H.Reads(reads), "sparse")
    assert cells2 == cells and np.array_equal(lf2, lf) and np.array_equal(lb2, lb)
    assert np.allclose(fr2, fr, rtol=1e-11, atol=1e-14)


def test_c3_fullsize_two_batches_and_arena_repeat(monkeypatch):
    """The memory-pressure paths at full N: a budget that holds two reads' rows forces a second batch; an arena sized at a
    fraction of the estimate is outgrown, the phase is repeated with the worst-case bound; dense groups of one read.  All of them must give
    the fixture's answers."""
    from dbgphmm_b200 import hmmv2 as H
    fx = load("c3")
    g = build(fx)
    N = g.n_nodes
    r0, r1 = (np.asarray(fx[f"r{i}_read"], np.uint8) for i in range(2))
    reads = [r0, r1, r0]
    want_lf = [float(fx["r0_logp_fwd"]), float(fx["r1_logp_fwd"]), float(fx["r0_logp_fwd"])]
    want = np.zeros(N)
    for i in (0, 1, 0):
        np.add.at(want, fx[f"r{i}_freq_idx"], fx[f"r{i}_freq_val"])
    li, lt = g.to_probs("normal")
    par = H.params_uniform(0.001); par.n_warmup = int(fx["k"][0])
    slab = ((N + 1) // 2 * 2 * 28 + 255) // 256 * 256
    per_read = 2 * slab + 2 * 10_000 * (40 * 48 + 256 + 3 * 64) + (1 << 20)
    monkeypatch.setenv("DBGPHMM_STRATEGY", "stream")
    monkeypatch.setenv("DBGPHMM_VERIFY", "1")
    for env, budget in (({}, int(2.4 * per_read)), ({"DBGPHMM_ARENA_EST_PCT": "30"}, 0), ({"DBGPHMM_DENSE_GROUP": "1"}, 0)):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        m = H.PHMMModel(g.src, g.dst, g.base, li, lt, par, mem_budget_bytes=budget)
        fr, lf, lb, cells = m.run_node_freqs(H.Reads(reads), "sparse")
        assert all(close(a, b) for a, b in zip(lf, want_lf)), (env, lf, want_lf)
        assert np.allclose(fr, want, rtol=1e-9, atol=1e-12), (env, np.abs(fr - want).max())
        m.close()
        for k in env:
            monkeypatch.delenv(k)


def test_c5_fullsize_read_against_the_oracle_fixture(monkeypatch, capsys):
    """BASELINE configs[4] graph (5 Mbp diploid, N = 6.66 M) with one full 20 kbp read (31 CPU-minutes in the oracle): active sets of every
    row through the single-read tables, then the bulk call in the stream strategy with the dense warm-up in GROUPS sharing one pool of
    slabs (the mode the library chooses by itself on this graph when a batch holds more reads than two slabs each leave room for)."""
    from dbgphmm_b200 import hmmv2 as H
    fx = load("c5")
    g = build(fx)
    m = model_of(H, g, int(fx["k"][0]))
    N = g.n_nodes
    read = np.asarray(fx["r0_read"], np.uint8)
    n = len(read)
    f = m.forward_sparse(read, False)
    assert close(f.full_prob(), float(fx["r0_logp_fwd"]))
    df, df_t, sf = compare_rows(f, fx, "r0_f_", n, "C5 fwd")
    del f
    b = m.backward_sparse(read)
    assert close(b.full_prob(), float(fx["r0_logp_bwd"]))
    db, db_t, sb = compare_rows(b, fx, "r0_b_", n, "C5 bwd")
    del b
    ties = fx["r0_ties"].tolist()
    with capsys.disabled():
        print(f"\nC5 read ({n} rows): active sets differ in {len(df)} forward / {len(db)} backward rows; ties across the top-n boundary in {ties[0]} / {ties[1]} rows")
    assert not sf and not sb, (sf[:3], sb[:3])
    assert len(df) == df_t and len(db) == db_t and len(df) <= ties[0] and len(db) <= ties[1], (df[:5], db[:5])
    want = np.zeros(N)
    np.add.at(want, fx["r0_freq_idx"], fx["r0_freq_val"])
    monkeypatch.setenv("DBGPHMM_STRATEGY", "stream")
    monkeypatch.setenv("DBGPHMM_VERIFY", "1")
    for group in ("2", "0"):     # three copies of the read in groups of two (one full group, one partial) ; then ungrouped
        monkeypatch.setenv("DBGPHMM_DENSE_GROUP", group)
        fr, lf, lb, cells = m.run_node_freqs(H.Reads([read, read, read]), "sparse")
        assert all(close(x, float(fx["r0_logp_fwd"])) for x in lf) and all(close(x, float(fx["r0_logp_bwd"])) for x in lb), group
        assert np.allclose(fr, 3.0 * want, rtol=1e-9, atol=1e-12), (group, np.abs(fr - 3.0 * want).max())
