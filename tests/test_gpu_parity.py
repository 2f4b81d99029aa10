"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): active-node sets and selected indices identical; ln P(R|X) and node frequencies within
1e-9 relative (f64).  The reference's own known-answer values (tests/golden/reference_kat.json) are also asserted
directly on the GPU results."""
import numpy as np
import pytest

from dbgphmm_b200 import graphs, synth
from oracle import oracle as O
from tests.common import (REL_TOL, assert_rows_match, assert_tables_match, close_log, gpu_model, kat, oracle_model,
                          oracle_params, random_linear_graph, row_order_exact, same_up_to_ties, to_gpu_params)

pytestmark = pytest.mark.gpu
K = kat()


@pytest.fixture(scope="module")
def H():
    from dbgphmm_b200 import hmmv2
    assert hmmv2.device_count() > 0, "no B200 visible"
    return hmmv2


def both(sg, param, mode="normal"):
    return gpu_model(sg, param, mode), oracle_model(sg, param, mode)


# ---------------------------------------------------------------- reference KATs straight on the GPU
@pytest.mark.parametrize("case", ["forward_zero_error", "forward_high_error"])
def test_kat_forward(H, case):
    c = K[case]
    g, _ = both(graphs.mock_linear(), oracle_params(c["p"]))
    f = g.forward(c["read"].encode())
    assert len(f) == 5
    for row, node, val in c.get("m", []):
        assert abs(f.row(row).m[node] - val) < c["eps"]
    for row, val in c["e"]:
        assert abs(f.row(row).e - val) < c["eps"]
    if c.get("all_i_d_zero"):
        for r in range(5):
            assert np.isneginf(f.row(r).i).all() and np.isneginf(f.row(r).d).all()
        assert np.isneginf(g.forward(c["impossible_read"].encode()).row(4).e)
    if "read2" in c:
        f2 = g.forward(c["read2"].encode())
        for row, val in c["e2"]:
            assert abs(f2.row(row).e - val) < c["eps"]


@pytest.mark.parametrize("case", ["backward_zero_error", "backward_high_error"])
def test_kat_backward(H, case):
    c = K[case]
    g, _ = both(graphs.mock_linear(), oracle_params(c["p"]))
    b = g.backward(c["read"].encode())
    for row, node, val in c.get("m", []):
        assert abs(b.row(row).m[node] - val) < c["eps"]
    for row, val in c["mb"]:
        assert abs(b.row(row).mb - val) < c["eps"]
    if "read2" in c:
        b2 = g.backward(c["read2"].encode())
        for row, val in c["mb2"]:
            assert abs(b2.row(row).mb - val) < c["eps"]


def test_kat_mapping_node_lists(H):
    c = K["hint_mock_linear_high_error"]
    g, _ = both(graphs.mock_linear(), oracle_params(c["p"]))
    o = g.run(c["read"].encode())
    hint = o.to_mapping(c["n_active"])
    assert [list(map(int, x)) for x in hint.nodes] == c["nodes"]
    maps = H.Mappings.from_list([hint])
    p1 = g.forward(c["read"].encode()).full_prob()
    p2 = g.forward_with_mapping(c["read"].encode(), maps, 0).full_prob()
    assert abs(p1 - p2) < c["max_log_diff_dense_vs_hint"]


def test_kat_hint_for_toy(H):
    c = K["hint_for_toy"]
    sg, k = graphs.toy_repeat()
    par = oracle_params(c["p"], n_warmup=k)
    g = gpu_model(sg, par, "non_zero")
    for case in c["cases"]:
        mp = g.generate_mappings(H.Reads([case["read"].encode()]), None, True)[0]
        assert [int(x[0]) for x in mp.nodes] == case["top1"]


# ---------------------------------------------------------------- row-by-row parity with the oracle
def _dbg_case(seed, glen=600, k=12, ploidy=2, het=0.02, p_err=0.01, read_len=150, n_reads=4):
    w = synth.make_workload("t", glen, k, 4, read_len, p_err, ploidy=ploidy, het=het, seed=seed, n_reads=n_reads)
    return w


@pytest.mark.parametrize("p", [0.0, 0.001, 0.1])
def test_dense_rows_linear(H, p):
    sg, seq = random_linear_graph(300, 1)
    g, o = both(sg, oracle_params(p))
    read = seq[100:160]
    assert_tables_match(g.forward(read), o.forward(read), sg.n_nodes, f"fwd p={p}")
    assert_tables_match(g.backward(read), o.backward(read), sg.n_nodes, f"bwd p={p}")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_dense_rows_dbg(H, seed):
    w = _dbg_case(seed)
    par = oracle_params(0.01, n_warmup=w.k)
    g, o = both(w.graph, par)
    for read in w.reads[:2]:
        assert_tables_match(g.forward(read), o.forward(read), w.graph.n_nodes, "fwd")
        assert_tables_match(g.backward(read), o.backward(read), w.graph.n_nodes, "bwd")


def test_dense_multi_chunk_graph(H):
    # more nodes than one tile (DENSE_CORE = 1024) so halos between chunks are exercised
    w = _dbg_case(7, glen=3000, k=16, read_len=60, n_reads=2)
    par = oracle_params(0.01, n_warmup=w.k)
    g, o = both(w.graph, par)
    assert w.graph.n_nodes > 2048
    for read in w.reads:
        assert_tables_match(g.forward(read), o.forward(read), w.graph.n_nodes, "fwd")
        assert_tables_match(g.backward(read), o.backward(read), w.graph.n_nodes, "bwd")


def test_dense_rows_on_a_chain_with_cross_edges(H):
    """Forks and merges inside the halos, short cycles: the tile planner used to lose a halo node whose first upstream neighbour lies
    on the other branch of a fork (host-side out-of-bounds write, garbage positions on the device).  Found by the sanitizer run of
    the host logic (tests/test_host_logic_asan.py); here the rows of such a graph against the oracle."""
    rng = np.random.default_rng(5)
    n = 1500
    src, dst = list(range(n - 1)), list(range(1, n))
    for _ in range(60):
        v = int(rng.integers(0, n)); w = int(np.clip(v + rng.integers(-30, 30), 0, n - 1))
        if (v, w) not in set(zip(src, dst)):
            src.append(v); dst.append(w)
    base = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
    sg = graphs.SeqGraph(src, dst, base, np.ones(n, np.int64))
    g, o = both(sg, oracle_params(0.01, n_warmup=20))
    read = bytes(base[400:440])
    assert_tables_match(g.forward(read), o.forward(read), n, "fwd")
    assert_tables_match(g.backward(read), o.backward(read), n, "bwd")
    gs, os_ = g.run(read), o.run(read)
    assert close_log(gs.to_full_prob_forward(), os_.to_full_prob_forward()).all()
    assert np.allclose(gs.to_node_freqs(), os_.to_node_freqs(), rtol=REL_TOL, atol=1e-12)


def test_dense_rows_on_a_small_k_graph_with_cycles(H):
    """k = 8: a branchy de Bruijn graph with cycles of halo nodes (the planner's climb to the top of a chain used to spin on them)."""
    w = _dbg_case(5, glen=3000, k=8, ploidy=1, het=0.0, read_len=40, n_reads=2)
    par = oracle_params(0.01, n_warmup=w.k)
    g, o = both(w.graph, par)
    for read in w.reads:
        assert_tables_match(g.forward(read), o.forward(read), w.graph.n_nodes, "fwd")
        assert_tables_match(g.backward(read), o.backward(read), w.graph.n_nodes, "bwd")


@pytest.mark.parametrize("seed,n_active", [(0, 40), (1, 10), (3, 80)])
def test_sparse_topn_rows_bit_exact_sets(H, seed, n_active):
    w = _dbg_case(seed)
    par = oracle_params(0.01, n_warmup=w.k, n_active=n_active)
    g, o = both(w.graph, par)
    for read in w.reads[:3]:
        assert_tables_match(g.forward_sparse(read, False), o.forward_sparse(read, False), w.graph.n_nodes, "fwd_sparse")
        assert_tables_match(g.backward_sparse(read), o.backward_sparse(read), w.graph.n_nodes, "bwd_sparse")


@pytest.mark.parametrize("seed", [0, 4])
def test_sparse_ratio_and_by_forward(H, seed):
    w = _dbg_case(seed, k=16, p_err=0.003)
    par = oracle_params(0.001, n_warmup=w.k, warmup_threshold=30)
    g, o = both(w.graph, par, "non_zero")
    for read in w.reads[:3]:
        gf, of = g.forward_sparse(read, True), o.forward_sparse(read, True)
        assert_tables_match(gf, of, w.graph.n_nodes, "fwd_ratio")
        assert_tables_match(g.backward_by_forward(read, gf), o.backward_by_forward(read, of), w.graph.n_nodes, "bwd_by_fwd")


@pytest.mark.parametrize("n_active", [10, 40])
def test_dense_selection_through_the_tile_prefilter(H, monkeypatch, n_active):
    """Top-n of a dense row via tile maxima (dense.cu: k_select_tilemax + the pre-filter of k_dense_select, the default on rows of
    >= 64 K cells) must pick exactly what the full radix sweep and the oracle pick.  ~90 tiles here: n_active = 40 leaves few spare
    tiles; the ratio selection (k = 400 > number of tiles) cannot use the pre-filter and must fall through to the full sweep."""
    w = _dbg_case(11, glen=30000, k=16, het=0.01, p_err=0.003, read_len=120, n_reads=2)
    assert w.graph.n_nodes > 40 * 512
    par = oracle_params(0.003, n_warmup=w.k, n_active=n_active, warmup_threshold=30)
    g, o = both(w.graph, par)
    read = w.reads[0]
    of, ob = o.forward_sparse(read, False), o.backward_sparse(read)
    before = H.launch_count()
    for flag in ("1", "0"):
        monkeypatch.setenv("DBGPHMM_SELECT_TILES", flag)
        gf = g.forward_sparse(read, False)
        assert_tables_match(gf, of, w.graph.n_nodes, f"fwd_sparse tiles={flag}")
        assert_tables_match(g.backward_sparse(read), ob, w.graph.n_nodes, f"bwd_sparse tiles={flag}")
        for r in (0, w.k // 2, w.k - 1):
            merged = of.row(r).merged(w.graph.n_nodes)
            for k in (1, n_active):
                a, b = gf.top_nodes(r, k), of.top_nodes(r, k)
                assert same_up_to_ties(a, b, merged[b]), (flag, r, k, list(a), list(b))
            a, b = gf.top_nodes_by_score_ratio(r, 30.0), of.top_nodes_by_score_ratio(r, 30.0)
            assert same_up_to_ties(a, b, merged[b]), (flag, r, list(a), list(b))
    assert H.launch_count() > before


def test_two_handles_driven_from_two_host_threads(H):
    """The reference's model is Sync and shared by rayon threads; here a handle belongs to one host thread at a time, but different
    handles may run concurrently (include/dbgphmm_b200.h): the device-memory cache they share must not hand a block of one
    thread's in-flight work to the other."""
    import threading
    w = _dbg_case(5, glen=2000, k=16, read_len=200, n_reads=6)
    par = oracle_params(0.01, n_warmup=w.k)
    models = [gpu_model(w.graph, par) for _ in range(2)]
    reads = [H.Reads(w.reads) for _ in range(2)]
    ref = models[0].run_node_freqs(reads[0], "sparse")
    out, err = [None, None], []

    def work(i):
        try:
            for mode in ("sparse", "dense", "sparse"):
                r = models[i].run_node_freqs(reads[i], mode)
                if mode == "sparse":
                    out[i] = r
        except Exception as e:  # noqa: BLE001
            err.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not err, err
    for o in out:
        assert np.allclose(o[0], ref[0], rtol=1e-12, atol=1e-15)
        assert np.array_equal(o[1], ref[1]) and np.array_equal(o[2], ref[2])


def _table_diff(a, b, n_nodes):
    """PHMMTable::diff (table.rs:183-191): sum of |pa - pb| over m, i, d and the three scalars, in linear space."""
    def dense(r):
        if r.is_dense:
            return np.exp(r.m), np.exp(r.i), np.exp(r.d)
        m = np.zeros(n_nodes); i = np.zeros(n_nodes); d = np.zeros(n_nodes)
        m[r.ids] = np.exp(r.m); i[r.ids] = np.exp(r.i); d[r.ids_d] = np.exp(r.d)
        return m, i, d
    x, y = dense(a), dense(b)
    t = sum(np.abs(u - v).sum() for u, v in zip(x, y))
    return t + sum(abs(np.exp(u) - np.exp(v)) for u, v in ((a.mb, b.mb), (a.ib, b.ib), (a.e, b.e)))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_dense_and_sparse_tables_sweep_of_the_reference(H, seed):
    """tests/hmm.rs:141-214 (hmmv2_forward_dense_and_sparse / hmmv2_backward_dense_and_sparse, #[ignore]d there for their run time):
    1000-node random linear PHMM, 100-base reads, (n_warmup, n_active_nodes) in [(10,10), (20,10), (40,10), (40,40), (40,80)]:
    the sparse tables equal the dense ones exactly inside the warm-up and differ by < 1e-6 at the far end."""
    sg, seq = random_linear_graph(1000, 2)
    rng = np.random.default_rng(seed)
    start = int(rng.integers(0, 900))
    read = bytearray(seq[start:start + 100])
    for pos in rng.integers(0, 100, 2):   # a couple of substitutions, like a sampled read at p = 0.01
        read[pos] = b"ACGT"[(b"ACGT".index(read[pos]) + 1) % 4]
    read = bytes(read)
    g = gpu_model(sg, oracle_params(0.01))
    fd, bd = g.forward(read), g.backward(read)
    n = len(read)
    for n_warmup, n_active in [(10, 10), (20, 10), (40, 10), (40, 40), (40, 80)]:
        g.set_params(to_gpu_params(oracle_params(0.01, n_warmup=n_warmup, n_active=n_active)))
        fs, bs = g.forward_sparse(read, False), g.backward_sparse(read)
        for i in range(n):
            if i < n_warmup:
                assert fs.row(i).is_dense and _table_diff(fd.row(i), fs.row(i), sg.n_nodes) == 0.0, (n_warmup, n_active, i)
            if n - i < n_warmup:
                assert bs.row(i).is_dense and _table_diff(bd.row(i), bs.row(i), sg.n_nodes) == 0.0, (n_warmup, n_active, i)
        assert _table_diff(fd.row(n - 1), fs.row(n - 1), sg.n_nodes) < 1e-6, (n_warmup, n_active)
        assert _table_diff(bd.row(0), bs.row(0), sg.n_nodes) < 1e-6, (n_warmup, n_active)


def test_cuda_matches_the_committed_oracle_fixture(H):
    """tests/golden/oracle_c2_small.json (written by tests/golden/make_oracle_fixture.py from the CPU oracle): ln P(R|X) and node
    frequencies within 1e-9 relative (north_star), active-node sets of the stored rows identical."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_oracle_fixture", os.path.join(here, "golden", "make_oracle_fixture.py"))
    fx = importlib.util.module_from_spec(spec); spec.loader.exec_module(fx)
    with open(os.path.join(here, "golden", "oracle_c2_small.json")) as fh:
        gold = json.load(fh)
    w, par, _ = fx.build()
    assert w.graph.n_nodes == gold["n_nodes"] and [len(r) for r in w.reads] == gold["read_lens"]
    g = gpu_model(w.graph, par)
    fr, lf, lb, cells = g.run_node_freqs(H.Reads(w.reads), "sparse")
    assert np.allclose(lf, gold["logp_forward"], rtol=REL_TOL, atol=0) and np.allclose(lb, gold["logp_backward"], rtol=REL_TOL, atol=0)
    assert abs(fr.sum() - gold["node_freq_sum"]) <= REL_TOL * gold["node_freq_sum"]
    for i, v in gold["node_freq_top"]:
        assert abs(fr[i] - v) <= REL_TOL * max(1.0, abs(v)), (i, fr[i], v)
    for ri in (0, 1):
        f, b = g.forward_sparse(w.reads[ri], False), g.backward_sparse(w.reads[ri])
        for x in (x for x in gold["rows"] if x["read"] == ri):
            fr_, br_ = f.row(x["row"]), b.row(x["bwd_row"])
            assert fr_.is_dense == x["fwd_is_dense"] and br_.is_dense == x["bwd_is_dense"]
            assert sorted(int(v) for v in fr_.ids) == x["fwd_ids"] and sorted(int(v) for v in fr_.ids_d) == x["fwd_ids_d"], (ri, x["row"])
            assert sorted(int(v) for v in br_.ids) == x["bwd_ids"] and sorted(int(v) for v in br_.ids_d) == x["bwd_ids_d"], (ri, x["bwd_row"])
            assert abs(fr_.e - x["fwd_e"]) <= REL_TOL * abs(x["fwd_e"]) and abs(br_.mb - x["bwd_mb"]) <= REL_TOL * abs(x["bwd_mb"])


def test_gathered_first_sparse_row_inputs_match_the_dense_slab(H, monkeypatch):
    """DBGPHMM_GATHER=1 (engine.cu: gather_prev0): the first sparse row of a top-n job reads the last dense row through a per-job
    list of gathered cells instead of the slab, which is released before the sparse phase.  Same active sets, hence the very same
    numbers, as the default path (and as the oracle, through the other tests of the stream strategy)."""
    w = _dbg_case(9, glen=4000, k=16, het=0.02, read_len=300, n_reads=6)
    par = oracle_params(0.01, n_warmup=w.k)
    g = gpu_model(w.graph, par)
    reads = H.Reads(w.reads)
    monkeypatch.setenv("DBGPHMM_STRATEGY", "stream")
    a = g.run_node_freqs(reads, "sparse")
    monkeypatch.setenv("DBGPHMM_GATHER", "1")
    b = g.run_node_freqs(reads, "sparse")
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert np.allclose(a[0], b[0], rtol=1e-12, atol=1e-15)
    of, olf, olb = oracle_model(w.graph, par).run_node_freqs(O.Reads(w.reads), "sparse")
    assert np.allclose(b[1], olf, rtol=REL_TOL, atol=0) and np.allclose(b[2], olb, rtol=REL_TOL, atol=0)
    assert np.allclose(b[0], of, rtol=REL_TOL, atol=1e-12)


@pytest.mark.parametrize("group", [1, 2, 4])
def test_dense_warmup_in_groups_sharing_one_pool_of_slabs(H, monkeypatch, group):
    """DBGPHMM_DENSE_GROUP=G (stream strategy): the dense warm-up rows, the top-n selection, the gather of the first sparse row's inputs
    and both recompute passes run over groups of G reads that reuse one pool of 2 G slabs.  Nothing about the arithmetic changes, so
    ln P is identical to the ungrouped run, node frequencies agree to rounding of the atomics, and both match the oracle."""
    w = _dbg_case(13, glen=4000, k=16, het=0.02, read_len=300, n_reads=7)
    par = oracle_params(0.01, n_warmup=w.k)
    g = gpu_model(w.graph, par)
    reads = H.Reads(w.reads)
    monkeypatch.setenv("DBGPHMM_STRATEGY", "stream")
    a = g.run_node_freqs(reads, "sparse")
    monkeypatch.setenv("DBGPHMM_DENSE_GROUP", str(group))
    b = g.run_node_freqs(reads, "sparse")
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert np.allclose(a[0], b[0], rtol=1e-12, atol=1e-15)
    of, olf, olb = oracle_model(w.graph, par).run_node_freqs(O.Reads(w.reads), "sparse")
    assert np.allclose(b[1], olf, rtol=REL_TOL, atol=0) and np.allclose(b[2], olb, rtol=REL_TOL, atol=0)
    assert np.allclose(b[0], of, rtol=REL_TOL, atol=1e-12)


def test_top_nodes_of_rows(H):
    w = _dbg_case(2)
    par = oracle_params(0.01, n_warmup=w.k)
    g, o = both(w.graph, par)
    read = w.reads[0]
    gf, of = g.forward_sparse(read, False), o.forward_sparse(read, False)
    for r in list(range(0, w.k + 3)) + [len(read) // 2, len(read) - 1]:
        merged = of.row(r).merged(w.graph.n_nodes)
        for k in (1, 10, 40):
            a, b = gf.top_nodes(r, k), of.top_nodes(r, k)
            assert same_up_to_ties(a, b, merged[b]), (r, k, list(a), list(b))
        a, b = gf.top_nodes_by_score_ratio(r, 30.0), of.top_nodes_by_score_ratio(r, 30.0)
        assert same_up_to_ties(a, b, merged[b]), (r, list(a), list(b))


def test_sparse_row_order_is_exact_when_nothing_ties(H):
    # on a unique-sequence linear genome with a clean read nothing ties: entry ORDER must then be identical too
    sg, seq = random_linear_graph(400, 9)
    par = oracle_params(0.01, n_warmup=20)
    g, o = both(sg, par)
    read = seq[150:300]
    for gt, ot in ((g.forward_sparse(read, False), o.forward_sparse(read, False)), (g.backward_sparse(read), o.backward_sparse(read))):
        exact = sum(row_order_exact(gt.row(r), ot.row(r)) for r in range(len(ot)))
        assert exact >= 0.9 * len(ot), exact


def _assert_mappings_equal(gm, om):
    """same candidate nodes per base (order up to ties within rounding) and the same ln probabilities per node"""
    for row in range(len(om.row_off) - 1):
        a, b = int(om.row_off[row]), int(om.row_off[row + 1])
        assert same_up_to_ties(gm.nodes[a:b], om.nodes[a:b], om.probs[a:b]), (row, gm.nodes[a:b], om.nodes[a:b])
        gp = dict(zip(gm.nodes[a:b].tolist(), gm.probs[a:b].tolist()))
        assert close_log([gp[int(k)] for k in om.nodes[a:b]], om.probs[a:b]).all(), row


def test_mapping_tables_and_generate_mappings(H):
    w = _dbg_case(5, k=16, p_err=0.003)
    par = oracle_params(0.001, n_warmup=w.k, warmup_threshold=30)
    g, o = both(w.graph, par, "non_zero")
    reads = w.reads[:3]
    gm = g.generate_mappings(H.Reads(reads), None, True)
    om = o.generate_mappings(O.Reads(reads), None, True)
    assert np.array_equal(gm.read_off, om.read_off)
    assert np.array_equal(gm.row_off, om.row_off), "per-base candidate counts differ"
    _assert_mappings_equal(gm, om)
    # forward / backward restricted to the mapping (forward.rs:51, backward.rs:59)
    gom = H.Mappings(om.read_off, om.row_off, om.nodes, om.probs)  # identical hint for both sides
    for r, read in enumerate(reads):
        gt, ot = g.forward_with_mapping(read, gom, r), o.forward_with_mapping(read, om[r])
        assert_tables_match(gt, ot, w.graph.n_nodes, "fwd_map")
        assert all(row_order_exact(gt.row(i), ot.row(i)) for i in range(len(ot))), "mapping rows keep the hint's node order"
        assert_tables_match(g.backward_with_mapping(read, gom, r), o.backward_with_mapping(read, om[r]), w.graph.n_nodes, "bwd_map")
    # Mappings::to_node_freqs
    assert np.allclose(gm.to_node_freqs(w.graph.n_nodes), om.to_node_freqs(w.graph.n_nodes), rtol=1e-9, atol=1e-12)
    # top-n mapping mode (use_max_ratio = false)
    gm2 = g.generate_mappings(H.Reads(reads), None, False)
    om2 = o.generate_mappings(O.Reads(reads), None, False)
    assert np.array_equal(gm2.row_off, om2.row_off)
    _assert_mappings_equal(gm2, om2)


@pytest.mark.parametrize("mode", ["dense", "sparse", "sparse_adaptive", "with_mapping"])
def test_run_node_freqs_and_logp(H, mode):
    w = _dbg_case(6, n_reads=6, k=16, p_err=0.003)
    par = oracle_params(0.001, n_warmup=w.k, warmup_threshold=30)
    g, o = both(w.graph, par, "non_zero")
    reads = w.reads
    gmaps = omaps = None
    if mode == "with_mapping":
        omaps = o.generate_mappings(O.Reads(reads), None, True)
        gmaps = H.Mappings(omaps.read_off, omaps.row_off, omaps.nodes, omaps.probs)
    gf, glf, glb, cells = g.run_node_freqs(H.Reads(reads), mode, True, gmaps)
    of, olf, olb = o.run_node_freqs(O.Reads(reads), mode, True, omaps)
    assert close_log(glf, olf).all(), (glf, olf)
    assert close_log(glb, olb).all(), (glb, olb)
    assert np.allclose(gf, of, rtol=REL_TOL, atol=1e-12), np.abs(gf - of).max()
    if mode != "with_mapping":
        ref_cells = [sum(o.count_cells(r, mode, True, d) for r in reads) for d in (1, 2)]
        assert list(cells) == ref_cells


def test_stream_strategy_matches_store_and_oracle(H, monkeypatch):
    """run_sparse with dense rows in ping-pong slabs + on-the-fly products (the strategy used when N is large)."""
    w = _dbg_case(9, n_reads=6)
    par = oracle_params(0.01, n_warmup=w.k)
    g, o = both(w.graph, par)
    reads = w.reads
    of, olf, olb = o.run_node_freqs(O.Reads(reads), "sparse", True, None)
    res = {}
    for strat in ("store", "stream"):
        monkeypatch.setenv("DBGPHMM_STRATEGY", strat)
        res[strat] = g.run_node_freqs(H.Reads(reads), "sparse", True, None)
        gf, glf, glb, cells = res[strat]
        assert close_log(glf, olf).all() and close_log(glb, olb).all()
        assert np.allclose(gf, of, rtol=REL_TOL, atol=1e-12), (strat, np.abs(gf - of).max())
    assert res["store"][3] == res["stream"][3]
    assert np.allclose(res["store"][0], res["stream"][0], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("cap", ["32", "48", "80"])
def test_sparse_jobs_that_outgrow_their_tables_are_carried_on_by_the_rescue_launch(H, monkeypatch, cap):
    """A small entry capacity makes most sparse jobs hand their previous row over to the rescue launch, many of them in the middle of
    a read; results, active sets and cell counts must not depend on where (or whether) that happens.  (32 entries = a 64-cell hash
    table: rows of this graph offer more distinct candidates than that -- the bounded probing must report the row as too large.)"""
    w = _dbg_case(13, glen=900, n_reads=6, read_len=300)
    par = oracle_params(0.01, n_warmup=w.k)
    g, o = both(w.graph, par)
    of, olf, olb = o.run_node_freqs(O.Reads(w.reads), "sparse", True, None)
    res = {}
    for tag, env in (("default", {}), ("small", {"DBGPHMM_SPARSE_CAP": cap}), ("rerun", {"DBGPHMM_SPARSE_CAP": cap, "DBGPHMM_SPARSE_RESCUE": "0"})):
        for k in ("DBGPHMM_SPARSE_CAP", "DBGPHMM_SPARSE_RESCUE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for strat in ("store", "stream"):
            monkeypatch.setenv("DBGPHMM_STRATEGY", strat)
            gf, glf, glb, cells = g.run_node_freqs(H.Reads(w.reads), "sparse", True, None)
            assert close_log(glf, olf).all() and close_log(glb, olb).all(), (tag, strat)
            assert np.allclose(gf, of, rtol=REL_TOL, atol=1e-12), (tag, strat)
            res[tag, strat] = (glf, glb, cells)
        t, u = g.forward_sparse(w.reads[0], False), o.forward_sparse(w.reads[0], False)
        tb, ub = g.backward_sparse(w.reads[0]), o.backward_sparse(w.reads[0])
        for a, b in ((t, u), (tb, ub)):
            for r in range(len(b)):
                ra, rb = a.row(r), b.row(r)
                assert ra.is_dense == rb.is_dense
                if not rb.is_dense:
                    assert list(ra.ids) == list(rb.ids), (tag, r)
    for key, v in res.items():
        assert np.array_equal(v[0], res["default", "store"][0]) and np.array_equal(v[1], res["default", "store"][1]) and v[2] == res["default", "store"][2], key


@pytest.mark.parametrize("strat", ["store", "stream"])
def test_batch_that_runs_out_of_device_memory_is_split(H, monkeypatch, strat):
    """The sparse-row arenas are sized for typical rows; a batch that outgrows what it can get (forced here by capping the arena)
    is split in two and tried again, with the same results as the unsplit batch."""
    w = _dbg_case(17, glen=900, n_reads=7, read_len=300)
    par = oracle_params(0.01, n_warmup=w.k)
    g = gpu_model(w.graph, par)
    monkeypatch.setenv("DBGPHMM_STRATEGY", strat)
    ref = g.run_node_freqs(H.Reads(w.reads), "sparse", True, None)
    monkeypatch.setenv("DBGPHMM_ARENA_MAX_BYTES", str(3 * 256 * 1024))   # three 256 KB pages: at most three jobs per direction
    got = g.run_node_freqs(H.Reads(w.reads), "sparse", True, None)
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2]) and got[3] == ref[3]
    assert np.allclose(got[0], ref[0], rtol=1e-12, atol=1e-15)
    monkeypatch.setenv("DBGPHMM_ARENA_MAX_BYTES", "4096")                 # not even one read fits: the error surfaces
    with pytest.raises(H.DbgphmmError) as ei:
        g.run_node_freqs(H.Reads(w.reads), "sparse", True, None)
    assert ei.value.status == H.ERR_OOM


def test_two_rows_per_launch_forward_kernel_and_its_fallback(H, monkeypatch):
    """Stream strategy: the forward warm-up runs two rows per launch (k_dense_fwd2) and must reproduce the single-row steps bit for
    bit; a two-row frame that is too narrow (forced here) makes the phase fall back to single-row steps with the same result."""
    w = _dbg_case(11, n_reads=5)
    par = oracle_params(0.003, n_warmup=w.k)
    g, o = both(w.graph, par)
    of, olf, olb = o.run_node_freqs(O.Reads(w.reads), "sparse", True, None)
    monkeypatch.setenv("DBGPHMM_STRATEGY", "stream")
    res = {}
    for tag, span2 in (("pair", None), ("fallback", "1")):
        if span2 is None:
            monkeypatch.delenv("DBGPHMM_DENSE_SPAN2", raising=False)
        else:
            monkeypatch.setenv("DBGPHMM_DENSE_SPAN2", span2)
        res[tag] = g.run_node_freqs(H.Reads(w.reads), "sparse", True, None)
        gf, glf, glb, cells = res[tag]
        assert close_log(glf, olf).all() and close_log(glb, olb).all()
        assert np.allclose(gf, of, rtol=REL_TOL, atol=1e-12)
    assert np.array_equal(res["pair"][1], res["fallback"][1]) and res["pair"][3] == res["fallback"][3]
    monkeypatch.setenv("DBGPHMM_STRATEGY", "store")
    monkeypatch.delenv("DBGPHMM_DENSE_SPAN2", raising=False)
    st = g.run_node_freqs(H.Reads(w.reads), "sparse", True, None)
    assert np.array_equal(res["pair"][1], st[1])


def test_full_prob_reads_batched_over_candidates(H):
    """to_full_prob_reads for a batch of candidate copy-number vectors X (posterior.rs:504-515)."""
    w = _dbg_case(8, n_reads=5, k=16, p_err=0.003)
    sg = w.graph
    par = oracle_params(0.001, n_warmup=w.k, warmup_threshold=30)
    g = gpu_model(sg, par, "non_zero")
    o = oracle_model(sg, par, "non_zero")
    reads = w.reads
    omaps = o.generate_mappings(O.Reads(reads), None, True)
    gmaps = H.Mappings(omaps.read_off, omaps.row_off, omaps.nodes, omaps.probs)
    rng = np.random.default_rng(0)
    # candidates: the true copy numbers, +1 on a few nodes, and +-1 (some nodes drop to 0 copies)
    X = np.stack([sg.node_copy_num] + [sg.node_copy_num + (rng.random(sg.n_nodes) < 0.03) for _ in range(2)]
                 + [np.maximum(0, sg.node_copy_num + rng.integers(-1, 2, sg.n_nodes)) for _ in range(3)])
    g.set_copy_nums_batch(X, "normal")
    tot, per = g.to_full_prob_reads(H.Reads(reads), gmaps)
    g.set_copy_nums_batch(X[:3], "normal")
    tot2, per2 = g.to_full_prob_reads(H.Reads(reads), None, True)
    g.set_copy_nums_batch(X, "normal")
    for x in range(len(X)):
        li, lt = sg.to_probs("normal", X[x])
        gl, gt = g.get_probs(x)
        assert close_log(gl, li, rel=1e-14).all() and close_log(gt, lt, rel=1e-14).all()
        o.set_probs(li, lt)
        s, p = o.to_full_prob_reads(O.Reads(reads), omaps)
        assert close_log(per[x], p).all(), (x, per[x], p)
        assert close_log(tot[x], s).all()
        if x < 3:
            s2, p2 = o.to_full_prob_reads(O.Reads(reads), None, True)
            assert close_log(per2[x], p2).all(), (x, per2[x], p2)


def test_mapping_rows_beyond_the_sparsevec_capacity(H):
    """Mapping rows: up to 400 distinct nodes (the capacity of the reference's SparseVec, table.rs:22) are scored -- through the wide
    shape of k_mapx for several candidates, through k_sparse for one -- and agree with the oracle; 401 nodes are an error where the
    reference panics (round 1 silently truncated such a row)."""
    w = _dbg_case(13, glen=900, n_reads=2, read_len=300)      # (a k-mer graph: candidates change init / trans through node copy numbers alone)
    sg = w.graph
    assert sg.n_nodes > 500
    par = oracle_params(0.01, n_warmup=w.k)
    g, o = both(sg, par)
    read = w.reads[0][:12]
    base = o.generate_mappings(O.Reads([read]), None, False)
    rng = np.random.default_rng(2)

    def maps(width):
        nodes, row_off = [], [0]
        for i in range(len(read)):
            have = [int(v) for v in base.nodes[int(base.row_off[i]):int(base.row_off[i + 1])]]
            row = have + [int(v) for v in rng.permutation(sg.n_nodes) if int(v) not in have][:width - len(have)]
            nodes += row; row_off.append(len(nodes))
        om = O.Mappings(np.array([0, len(read)], np.uint64), np.array(row_off, np.uint64), np.array(nodes, np.uint32), np.zeros(len(nodes)))
        return om, H.Mappings(om.read_off, om.row_off, om.nodes, om.probs)

    om, gm = maps(400)
    X = np.stack([sg.node_copy_num, sg.node_copy_num + 1, sg.node_copy_num + (rng.random(sg.n_nodes) < 0.3)])
    g.set_copy_nums_batch(X, "normal")
    tot, per = g.to_full_prob_reads(H.Reads([read]), gm)
    for x in range(len(X)):
        li, lt = sg.to_probs("normal", X[x])
        o.set_probs(li, lt)
        s, p = o.to_full_prob_reads(O.Reads([read]), om)
        assert close_log(per[x], p).all(), (x, per[x], p)
    g.set_copy_nums_batch(X[:1], "normal")
    tot1, per1 = g.to_full_prob_reads(H.Reads([read]), gm)
    assert per1[0, 0] == per[0, 0]
    _, gm401 = maps(401)
    for cand in (X, X[:1]):
        g.set_copy_nums_batch(cand, "normal")
        with pytest.raises(H.DbgphmmError) as ei:
            g.to_full_prob_reads(H.Reads([read]), gm401)
        assert ei.value.status == H.ERR_CAPACITY


def test_capacity_overflow_is_an_error_like_the_reference_panic(H):
    # ratio mode leaving warm-up with > 200 candidates overflows the 400-entry SparseVec (params.rs:37-38)
    w = _dbg_case(0)
    par = oracle_params(0.01, n_warmup=w.k)
    g, o = both(w.graph, par, "non_zero")
    with pytest.raises(RuntimeError):
        o.forward_sparse(w.reads[0], True)
    with pytest.raises(H.DbgphmmError) as ei:
        g.forward_sparse(w.reads[0], True)
    assert ei.value.status == H.ERR_CAPACITY


def test_zero_probability_read_is_an_error(H):
    g, _ = both(graphs.mock_linear(), oracle_params(0.0))
    with pytest.raises(H.DbgphmmError) as ei:
        g.run(b"CGATT").to_node_freqs()  # the reference divides by P = 0 and panics on NaN ordering
    assert ei.value.status == H.ERR_ZERO_PROB


def test_read_shorter_than_warmup_and_single_base(H):
    sg, seq = random_linear_graph(120, 3)
    par = oracle_params(0.01, n_warmup=40)
    g, o = both(sg, par)
    for read in (seq[5:6], seq[10:25]):
        assert_tables_match(g.forward_sparse(read, False), o.forward_sparse(read, False), sg.n_nodes, "fwd")
        assert_tables_match(g.backward_sparse(read), o.backward_sparse(read), sg.n_nodes, "bwd")
        gf, of = g.forward_sparse(read, True), o.forward_sparse(read, True)
        assert_tables_match(gf, of, sg.n_nodes, "fwd ratio")
        assert_tables_match(g.backward_by_forward(read, gf), o.backward_by_forward(read, of), sg.n_nodes, "bwd by fwd")


# ---------------------------------------------------------------- edge / init frequencies (freq.rs:276-389)
def test_edge_freqs_reference_kat_on_mock_crossing(H):
    """graph/seq_graph.rs:440-504 asserted on the GPU results, plus the oracle on every edge."""
    ra, rb = b"ATTAGGAGCA", b"ATTAGGAGCAGCTGATAGGG"
    for flag in (False, True):
        sg = graphs.mock_crossing(flag)
        g, o = both(sg, oracle_params(0.01))
        for r in (ra, rb):
            gef, gnf = g.run(r).to_edge_and_init_freqs()
            oef, onf = o.run(r).to_edge_and_init_freqs(o, r)
            assert np.allclose(gef, oef, rtol=REL_TOL, atol=1e-14), np.abs(gef - oef).max()
            assert np.allclose(gnf, onf, rtol=REL_TOL, atol=1e-14)
        ef = g.run(rb).to_edge_freqs()
        if not flag:
            assert ef[36] < 0.0001 and ef[37] > 0.9 and ef[38] < 0.0001 and ef[39] < 0.0001
        else:
            assert ef[37] == 0.0 and ef[38] == 0.0


@pytest.mark.parametrize("mode", ["dense", "sparse", "sparse_adaptive", "with_mapping"])
def test_edge_and_init_freqs_all_row_kinds(H, mode):
    """dense x dense, dense x sparse, sparse x sparse and b_init pairs of rows, against the oracle."""
    w = _dbg_case(11, n_reads=3, k=12, p_err=0.004)
    par = oracle_params(0.002, n_warmup=8, warmup_threshold=30)
    g, o = both(w.graph, par, "non_zero")
    for read in w.reads[:3]:
        read = read[:60]
        if mode == "dense":
            go, oo = g.run(read), o.run(read)
        elif mode == "sparse":
            go, oo = g.run_sparse(read), o.run_sparse(read)
        elif mode == "sparse_adaptive":
            go, oo = g.run_sparse_adaptive(read, False), o.run_sparse_adaptive(read, False)
        else:
            om = o.generate_mappings(O.Reads([read]), None, False)
            gm = H.Mappings(om.read_off, om.row_off, om.nodes, om.probs)
            go, oo = g.run_with_mapping(read, gm, 0), o.run_with_mapping(read, om[0])
        gef, gnf = go.to_edge_and_init_freqs()
        oef, onf = oo.to_edge_and_init_freqs(o, read)
        assert np.allclose(gef, oef, rtol=REL_TOL, atol=1e-13), (mode, np.abs(gef - oef).max())
        assert np.allclose(gnf, onf, rtol=REL_TOL, atol=1e-13), (mode, np.abs(gnf - onf).max())
        gq, oq = g.q_score_exact(gef, gnf), o.q_score_exact(oef, onf)   # q.rs:66-96
        assert np.allclose(gq, oq, rtol=1e-9, atol=1e-12), (gq, oq)


def test_batched_candidates_wide_and_duplicated_mapping_rows(H, monkeypatch):
    """Mapping rows the 8-candidates-per-CTA kernel does not take (more than 64 nodes, a duplicated node) fall back to the
    general sparse kernel: both routes must agree with the oracle and with each other."""
    w = _dbg_case(13, n_reads=3, k=14, p_err=0.003)
    sg = w.graph
    par = oracle_params(0.001, n_warmup=w.k, warmup_threshold=30)
    g, o = both(sg, par, "non_zero")
    reads = [r[:80] for r in w.reads]
    base = o.generate_mappings(O.Reads(reads), None, False)
    rng = np.random.default_rng(4)
    read_off, row_off, nodes = [0], [0], []
    for r in range(len(reads)):
        for i in range(len(reads[r])):
            a, b = int(base.row_off[int(base.read_off[r]) + i]), int(base.row_off[int(base.read_off[r]) + i + 1])
            row = list(base.nodes[a:b])
            if r == 0 and i % 7 == 3:      # wide rows: 70 - 90 distinct nodes
                extra = [int(v) for v in rng.permutation(sg.n_nodes)[:90] if int(v) not in row]
                row = row + extra[:rng.integers(70, 90) - len(row)]
            if r == 1 and i % 5 == 2:      # a duplicated node (first occurrence wins, active_nodes.rs:15-56 style)
                row = row + [row[0]]
            nodes += row; row_off.append(len(nodes))
        read_off.append(len(row_off) - 1)
    probs = np.zeros(len(nodes))
    om = O.Mappings(np.array(read_off, np.uint64), np.array(row_off, np.uint64), np.array(nodes, np.uint32), probs)
    gm = H.Mappings(om.read_off, om.row_off, om.nodes, om.probs)
    X = np.stack([sg.node_copy_num, sg.node_copy_num + (rng.random(sg.n_nodes) < 0.05), sg.node_copy_num + 1])
    g.set_copy_nums_batch(X, "normal")
    tot, per = g.to_full_prob_reads(H.Reads(reads), gm)
    monkeypatch.setenv("DBGPHMM_NO_MAPX", "1")
    tot1, per1 = g.to_full_prob_reads(H.Reads(reads), gm)
    assert np.allclose(per, per1, rtol=1e-12, atol=0)
    for x in range(len(X)):
        li, lt = sg.to_probs("normal", X[x])
        o.set_probs(li, lt)
        s, p = o.to_full_prob_reads(O.Reads(reads), om)
        assert close_log(per[x], p).all(), (x, per[x], p)


def test_reference_surface_compositions(H):
    """The rest of the reference's method surface, which the Python mirror composes on the host from the same C-ABI
    calls: PHMMOutput::{to_emit_probs, to_state_probs} (table.rs:500-505, freq.rs:226-239) over exported device rows, and
    PHMMModel::{to_node_freqs, to_full_prob(_parallel), to_full_prob_sparse(_backward), forward_*_score_only} (freq.rs:87-165,
    forward.rs:79-89,158-206) as one bulk call each.  Against the oracle doing it the reference's way, read by read
    (tests/test_surface_host.py runs the same compositions over the oracle on the CPU)."""
    w = _dbg_case(8, n_reads=5, k=16, p_err=0.003)
    par = oracle_params(0.001, n_warmup=w.k, warmup_threshold=30)
    g, o = both(w.graph, par, "non_zero")
    seqs = w.reads
    N = w.graph.n_nodes
    for go, oo in ((g.run(seqs[0]), o.run(seqs[0])), (g.run_sparse(seqs[0]), o.run_sparse(seqs[0]))):
        sp = go.to_state_probs()
        freqs = np.exp(sp.merged(N))
        assert np.allclose(freqs, go.to_node_freqs(), rtol=1e-9, atol=1e-12)      # the device's own product kernels
        assert np.allclose(freqs, oo.to_node_freqs(), rtol=1e-9, atol=1e-12)      # the oracle
        n = go.n_emissions()
        t = go.to_emit_probs(n // 2)
        f, b, p = oo.forward.row(n // 2 - 1), oo.backward.row(n // 2), oo.forward.full_prob()
        want = H._dense_states(f, N)[0] + H._dense_states(b, N)[0] - p
        big = want > -600.0     # (below that a state may have been flushed beside a much larger state of the same cell, tests/common.py)
        assert close_log(t.m[big], want[big]).all() and (t.m[~big] < -590.0).all()
        assert sum(1 for _ in go.iter_emit_probs()) == n + 1
    want = sum(o.run(x).to_node_freqs() for x in seqs)
    assert np.allclose(g.to_node_freqs(seqs), want, rtol=1e-9, atol=1e-12)
    want = sum(o.forward(x).full_prob() for x in seqs)
    assert close_log(g.to_full_prob(seqs), want).all() and close_log(g.to_full_prob_parallel(H.Reads(seqs)), want).all()
    for ratio in (False, True):
        want = sum(o.forward_sparse(x, ratio).full_prob() for x in seqs)
        assert close_log(g.to_full_prob_sparse(seqs, ratio), want).all()
        assert close_log(g.forward_sparse_score_only(seqs[1], ratio), o.forward_sparse(seqs[1], ratio).full_prob()).all()
    want = sum(o.backward_sparse(x).full_prob() for x in seqs)
    assert close_log(g.to_full_prob_sparse_backward(seqs), want).all()
    om = o.generate_mappings(O.Reads(seqs), None, True)
    hm = H.Mappings(om.read_off, om.row_off, om.nodes, om.probs)
    for r in (0, 3):
        want = o.forward_with_mapping(seqs[r], om[r]).full_prob()
        assert close_log(g.forward_with_mapping_score_only(seqs[r], hm, r), want).all()
