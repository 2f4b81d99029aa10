"""GPU tests on the BASELINE.json configurations (or scaled versions that the CPU oracle finishes in seconds) plus
size-independent properties at larger N."""
import numpy as np
import pytest

from dbgphmm_b200 import graphs, synth
from oracle import oracle as O
from tests.common import REL_TOL, assert_tables_match, close_log, gpu_model, oracle_model, oracle_params

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
