"""GPU tests on the BASELINE.json configurations (or scaled versions that the CPU oracle finishes in seconds) plus
size-independent properties at larger N."""
import numpy as np
import pytest

from dbgphmm_b200 import graphs, synth
from oracle import oracle as O
from tests.common import REL_TOL, assert_tables_match, close_log, gpu_model, oracle_model, oracle_params, same_up_to_ties

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    from dbgphmm_b200 import hmmv2
    assert hmmv2.device_count() > 0
    return hmmv2


@pytest.fixture(scope="module")
def c1():
    """configs[0] / configs[1]: 10 kbp haploid genome, 1 kbp HiFi reads (p = 0.001), k = 40 draft DBG."""
    return synth.make_workload("C1", 10_000, 40, 20, 1_000, 0.001, ploidy=1, seed=0, n_reads=6)


def test_c1_dense_forward_backward_on_the_real_graph(H, c1):
    par = oracle_params(0.001, n_warmup=c1.k)
    g, o = gpu_model(c1.graph, par), oracle_model(c1.graph, par)
    assert c1.graph.n_nodes >= 10_000
    read = c1.reads[0][:120]   # dense rows over all N nodes; the oracle needs ~1 s per 100 rows here
    assert_tables_match(g.forward(read), o.forward(read), c1.graph.n_nodes, "C1 fwd")
    assert_tables_match(g.backward(read), o.backward(read), c1.graph.n_nodes, "C1 bwd")
    gf, glf, glb, _ = g.run_node_freqs(H.Reads([read]), "dense")
    of, olf, olb = o.run_node_freqs(O.Reads([read]), "dense")
    assert close_log(glf, olf).all() and close_log(glb, olb).all()
    assert np.allclose(gf, of, rtol=REL_TOL, atol=1e-12)


def test_c2_sparse_full_reads_active_sets_and_freqs(H, c1):
    """configs[1]: sparse mode, n_active_nodes = 40, n_warmup = k = 40, full 1 kbp reads."""
    par = oracle_params(0.001, n_warmup=c1.k)
    g, o = gpu_model(c1.graph, par), oracle_model(c1.graph, par)
    reads = c1.reads[:3]
    for read in reads[:2]:
        assert_tables_match(g.forward_sparse(read, False), o.forward_sparse(read, False), c1.graph.n_nodes, "C2 fwd")
        assert_tables_match(g.backward_sparse(read), o.backward_sparse(read), c1.graph.n_nodes, "C2 bwd")
    gf, glf, glb, cells = g.run_node_freqs(H.Reads(reads), "sparse")
    of, olf, olb = o.run_node_freqs(O.Reads(reads), "sparse")
    assert close_log(glf, olf).all() and close_log(glb, olb).all()
    assert np.allclose(gf, of, rtol=REL_TOL, atol=1e-12)
    assert list(cells) == [sum(o.count_cells(r, "sparse", True, d) for r in reads) for d in (1, 2)]


def test_c3_scaled_diploid_stream_strategy_properties(H, monkeypatch):
    """configs[2] scaled to 50 kbp: diploid 1 % het, k = 40, stream strategy; oracle on one read, properties on all."""
    w = synth.make_workload("C3s", 50_000, 40, 1, 2_000, 0.001, ploidy=2, het=0.01, seed=5, n_reads=12)
    par = oracle_params(0.001, n_warmup=w.k)
    g, o = gpu_model(w.graph, par), oracle_model(w.graph, par)
    res = {}
    for strat in ("store", "stream"):
        monkeypatch.setenv("DBGPHMM_STRATEGY", strat)
        res[strat] = g.run_node_freqs(H.Reads(w.reads), "sparse")
    fs, lfs, lbs, cs = res["stream"]
    ft, lft, lbt, ct = res["store"]
    assert cs == ct and np.array_equal(lfs, lft) and np.array_equal(lbs, lbt)
    assert np.allclose(fs, ft, rtol=1e-12, atol=1e-15)
    n_bases = sum(len(r) for r in w.reads)
    # every base is emitted by exactly one Match/Ins state; silent Del mass adds a little, the active-set truncation removes a little
    assert abs(fs.sum() / n_bases - 1.0) < 1e-3
    assert np.allclose(lfs, lbs, rtol=1e-3)              # forward and backward totals differ only by the bounded Del chain (SURVEY §8a gotcha 10)
    of, olf, olb = o.run_node_freqs(O.Reads(w.reads[:1]), "sparse")
    assert close_log(lfs[:1], olf).all() and close_log(lbs[:1], olb).all()
    f2 = g.run_node_freqs(H.Reads(w.reads[:1]), "sparse")[0]
    assert np.allclose(f2, of, rtol=REL_TOL, atol=1e-12)


def test_c4_batched_candidates_on_a_tandem_repeat(H):
    """configs[3] scaled: tandem-repeat region, mapping-restricted P(R|X) for a batch of candidate copy numbers."""
    hap = synth.tandem_repeat_genome(400, 6, 600, seed=3, divergence=0.01)
    hap2 = synth.mutate_substitutions(hap, 0.003, 77)
    sg, _ = graphs.build_dbg([hap.tobytes(), hap2.tobytes()], 24, seed=9)
    reads = synth.sample_reads([hap, hap2], 3, 800, 0.002, 13)[:8]
    par = oracle_params(0.001, n_warmup=24, warmup_threshold=40)
    g = gpu_model(sg, par, "non_zero")
    o = oracle_model(sg, par, "non_zero")
    omaps = o.generate_mappings(O.Reads(reads), None, True)
    gmaps = H.Mappings(omaps.read_off, omaps.row_off, omaps.nodes, omaps.probs)
    rng = np.random.default_rng(1)
    B = 24
    X = np.stack([sg.node_copy_num + (rng.random(sg.n_nodes) < 0.02) * rng.integers(1, 3, sg.n_nodes) for _ in range(B)])
    X[0] = sg.node_copy_num
    g.set_copy_nums_batch(X, "normal")
    tot, per = g.to_full_prob_reads(H.Reads(reads), gmaps)
    assert tot.shape == (B,) and per.shape == (B, len(reads))
    for x in (0, 5, B - 1):
        li, lt = sg.to_probs("normal", X[x])
        o.set_probs(li, lt)
        s, p = o.to_full_prob_reads(O.Reads(reads), omaps)
        assert close_log(per[x], p).all(), (x, per[x], p)
        assert close_log(tot[x], s).all()


def test_c4_full_size_region_64_candidates(H):
    """configs[3] at its size: 200 kbp KIR-like region (16 x 10 kbp units, two haplotypes, N ~ 94 k), 10 kbp reads, B = 64 candidate copy-number
    vectors scored with mappings in one batched call (posterior.rs:504-515) -- the workload of bench.py's C4 line.  The mappings
    come from the GPU (generate_mappings), are checked against the oracle's on the first reads, and every candidate's per-read
    ln P(R|X) is compared with the oracle given the same mappings."""
    import os
    hap = synth.tandem_repeat_genome(10_000, 16, 20_000, seed=3, divergence=0.005)
    hap2 = synth.mutate_substitutions(hap, 0.002, 77)
    sg, _ = graphs.build_dbg([hap.tobytes(), hap2.tobytes()], 40, seed=9)
    assert sg.n_nodes > 80_000
    reads = synth.sample_reads([hap, hap2], 20, 10_000, 0.001, 13)[:6]
    par = oracle_params(0.001, n_warmup=40)
    g = gpu_model(sg, par, "non_zero")
    o = oracle_model(sg, par, "non_zero")
    gmaps = g.generate_mappings(H.Reads(reads), None, False)
    omaps2 = o.generate_mappings(O.Reads(reads[:2]), None, False, n_threads=2)
    n2 = int(omaps2.read_off[2])
    assert np.array_equal(gmaps.row_off[:n2 + 1], omaps2.row_off)
    # A row lists the 40 most probable nodes of a base; inside a 16-copy repeat its tail holds nodes at e^-50 of the row's mass whose
    # presence (and, through the active sets upstream, whose value) depends on how (near-)ties at the top-n boundary of the forward / backward active sets were broken (log-space f64 in the
    # oracle, exponent-extended linear f64 here: SURVEY 8c, unpinned in the reference).  Everything within e^-30 of the row's best node
    # -- the part `to_mapping_by_score_ratio(30)` would keep -- must agree exactly as a set and in value; the count of rows that are
    # identical lists is reported and must be the overwhelming majority.
    n_same = 0
    for r in range(n2):
        a, b = int(omaps2.row_off[r]), int(omaps2.row_off[r + 1])
        gi, gp, oi, op = gmaps.nodes[a:b], gmaps.probs[a:b], omaps2.nodes[a:b], omaps2.probs[a:b]
        n_same += int(np.array_equal(gi, oi))
        floor = op.max() - 30.0
        gh = {int(i): p for i, p in zip(gi, gp) if p >= floor}
        oh = {int(i): p for i, p in zip(oi, op) if p >= floor}
        assert set(gh) == set(oh), (r, sorted(set(gh) ^ set(oh)))
        assert all(abs(gh[i] - oh[i]) <= 1e-7 * max(1.0, abs(oh[i])) for i in oh), r
    print(f"C4 mappings: {n_same} of {n2} rows are identical lists; the others differ below e^-30 of the row's best node")
    assert n_same >= 0.98 * n2
    omaps = O.Mappings(gmaps.read_off, gmaps.row_off, gmaps.nodes, gmaps.probs)
    rng = np.random.default_rng(1)
    B = 64
    cn = sg.node_copy_num
    X = np.repeat(cn[None, :], B, 0).astype(np.uint32)
    rep = np.where(cn >= 2)[0]
    for b in range(1, B):
        idx = rng.choice(rep, size=40, replace=False)
        X[b, idx] = np.maximum(1, X[b, idx].astype(np.int64) + rng.choice([-1, 1], size=len(idx))).astype(np.uint32)
    g.set_copy_nums_batch(X, "non_zero")
    tot, per = g.to_full_prob_reads(H.Reads(reads), gmaps)
    assert tot.shape == (B,) and per.shape == (B, len(reads))
    threads = max(1, min(8, len(os.sched_getaffinity(0))))
    for x in range(B):
        li, lt = sg.to_probs("non_zero", X[x])
        o.set_probs(li, lt)
        s, p = o.to_full_prob_reads(O.Reads(reads), omaps, n_threads=threads)
        assert close_log(per[x], p).all(), (x, per[x], p)
        assert close_log(tot[x], s).all()
    assert len(set(np.round(tot, 6))) > B // 2          # the candidates are really different models


def test_fast_and_exact_dense_kernels_agree_on_a_large_graph(H, monkeypatch):
    """The common-frame register-stencil kernel against the per-value-exponent kernel (which the small cases pin to the oracle)
    on a graph large enough to contain every tile shape: merge nodes, irregular chains, padded layouts, halo-only tiles."""
    w = synth.make_workload("C3m", 200_000, 40, 1, 1_500, 0.001, ploidy=2, het=0.01, seed=5, n_reads=6)
    par = oracle_params(0.001, n_warmup=w.k)
    g = gpu_model(w.graph, par)
    read = w.reads[0][:8]
    N = w.graph.n_nodes
    rows, runs = {}, {}
    for exact in ("1", "0"):
        monkeypatch.setenv("DBGPHMM_FORCE_EXACT", exact)
        tf, tb = g.forward(read), g.backward(read)
        rows[exact] = [tf.row(i).merged(N) for i in range(len(read))] + [tb.row(i).merged(N) for i in range(len(read))]
        runs[exact] = g.run_node_freqs(H.Reads(w.reads), "sparse")
    for a, b in zip(rows["1"], rows["0"]):
        assert not np.isnan(b).any()
        fin = np.isfinite(a)
        assert np.array_equal(fin, np.isfinite(b))
        assert np.allclose(a[fin], b[fin], rtol=REL_TOL, atol=0)
    fa, lfa, lba, ca = runs["1"]
    fb, lfb, lbb, cb = runs["0"]
    assert ca == cb and close_log(lfa, lfb).all() and close_log(lba, lbb).all()
    assert np.allclose(fa, fb, rtol=REL_TOL, atol=1e-12)
