"""Pins the CPU oracle against the reference's own known-answer tests (tests/golden/reference_kat.json).

These are the values the reference asserts in its CI-run unit tests; an oracle that reproduces them (and the exact
node lists) follows the same recurrence, fold order and selection order as src/hmmv2."""
import math

import numpy as np
import pytest

from dbgphmm_b200 import graphs
from oracle import oracle as O
from tests.common import kat, oracle_model, oracle_params

K = kat()


def linear_model(p):
    sg = graphs.mock_linear()
    return oracle_model(sg, oracle_params(p))


def test_prob_add_identities():
    # prob.rs:181-197 special cases
    ninf = -math.inf
    assert O.padd(ninf, 0.0) == 0.0 and O.padd(0.0, ninf) == 0.0
    assert O.padd(ninf, ninf) == ninf
    assert O.padd(math.log(0.3), math.log(0.3)) == math.log(0.3) + math.log(2.0)
    assert abs(O.padd(math.log(0.3), math.log(0.2)) - math.log(0.5)) < 1e-15


def test_params_uniform_matches_reference_formula():
    q = O.params_uniform(0.01)  # params.rs:73-124
    assert q.n_active_nodes == 40 and q.n_warmup == 50 and q.n_max_gaps == 4 and q.warmup_threshold == 200
    assert q.active_node_max_ratio == 30.0
    assert abs(math.exp(q.p_MM) - (1 - 0.02 - 1e-5)) < 1e-15
    assert abs(math.exp(q.p_DM) - (1 - 0.02 - 1e-5)) < 1e-15
    assert q.p_random == math.log(0.25) and abs(math.exp(q.p_match) - 0.99) < 1e-15


@pytest.mark.parametrize("case", ["forward_zero_error", "forward_high_error"])
def test_forward_kat(case):
    c = K[case]
    m = linear_model(c["p"])
    f = m.forward(c["read"].encode())
    assert len(f) == 5
    for row, node, val in c.get("m", []):
        assert abs(f.row(row).m[node] - val) < c["eps"]
    for row, val in c["e"]:
        assert abs(f.row(row).e - val) < c["eps"]
    if c.get("all_i_d_zero"):
        for r in range(5):
            assert np.isneginf(f.row(r).i).all() and np.isneginf(f.row(r).d).all()
        assert np.isneginf(m.forward(c["impossible_read"].encode()).row(4).e)
    if "read2" in c:
        f2 = m.forward(c["read2"].encode())
        for row, val in c["e2"]:
            assert abs(f2.row(row).e - val) < c["eps"]
        assert abs(f2.row(3).e - f.row(3).e) < c["eps"]


@pytest.mark.parametrize("case", ["backward_zero_error", "backward_high_error"])
def test_backward_kat(case):
    c = K[case]
    m = linear_model(c["p"])
    b = m.backward(c["read"].encode())
    for row, node, val in c.get("m", []):
        assert abs(b.row(row).m[node] - val) < c["eps"]
    for row, val in c["mb"]:
        assert abs(b.row(row).mb - val) < c["eps"]
    if "read2" in c:
        b2 = m.backward(c["read2"].encode())
        for row, val in c["mb2"]:
            assert abs(b2.row(row).mb - val) < c["eps"]


def test_mapping_node_lists_exact():
    c = K["hint_mock_linear_high_error"]
    m = linear_model(c["p"])
    o = m.run(c["read"].encode())
    hint = o.to_mapping(c["n_active"])
    assert [list(map(int, x)) for x in hint.nodes] == c["nodes"]
    p1 = m.forward(c["read"].encode()).full_prob()
    p2 = m.forward_with_mapping(c["read"].encode(), hint).full_prob()
    assert abs(p1 - p2) < c["max_log_diff_dense_vs_hint"]
    cb = K["backward_with_hint"]
    hint5 = o.to_mapping(cb["n_active"])
    b1 = m.backward(cb["read"].encode()).full_prob()
    b2 = m.backward_with_mapping(cb["read"].encode(), hint5).full_prob()
    assert abs(b1 - b2) < cb["max_log_diff"]


def test_hint_for_toy_repeat():
    c = K["hint_for_toy"]
    sg, k = graphs.toy_repeat()
    par = oracle_params(c["p"], n_warmup=k)  # MultiDbg::generate_mappings sets n_warmup = k (posterior.rs:615)
    m = oracle_model(sg, par, "non_zero")    # to_non_zero_phmm (posterior.rs:617)
    for case in c["cases"]:
        mp = m.generate_mappings(O.Reads([case["read"].encode()]), None, True)[0]
        assert [int(x[0]) for x in mp.nodes] == case["top1"]


def test_forward_sparse_equals_dense_inside_warmup():
    # hmm_forward_mock_sparse (forward.rs:621-638): 100-node linear, 32-base read, all rows inside warm-up
    rng = np.random.default_rng(0)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 100)].tobytes()
    sg = graphs.genome_graph_to_seq_graph([(seq, 1)])
    m = oracle_model(sg, oracle_params(0.01))
    read = seq[30:62]
    r1, r2 = m.forward(read), m.forward_sparse(read, False)
    for i in range(len(r1)):
        a, b = r1.row(i), r2.row(i)
        assert a.is_dense and b.is_dense
        assert np.array_equal(a.m, b.m) and np.array_equal(a.i, b.i) and np.array_equal(a.d, b.d)


def test_freq_invariants():
    # freq.rs:434-610: zero-error forward total == backward total; node freqs ~1 on the true path
    sg = graphs.mock_linear()
    m0 = oracle_model(sg, oracle_params(0.0))
    o = m0.run(b"CGATC")
    assert abs(o.to_full_prob_forward() - o.to_full_prob_backward()) < 1e-7
    fr = o.to_node_freqs()
    assert (fr[3:8] > 0.98).all() and (np.delete(fr, range(3, 8)) < 0.01).all()
    m1 = oracle_model(sg, oracle_params(0.01))
    fr1 = m1.run(b"CGATC").to_node_freqs()
    assert (fr1[3:8] > 0.98).all()


def test_seqgraph_copy_numbers_to_probs_match_oracle():
    # graphs.copy_nums_to_probs (product-side numpy) vs the oracle's restatement of seq_graph.rs:160-273
    rng = np.random.default_rng(5)
    for trial in range(5):
        hap = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 300)].tobytes()
        g, _ = graphs.build_dbg([hap, hap[:150] + b"A" + hap[151:]], 8, seed=trial)
        cn = rng.integers(0, 4, g.n_nodes)
        for mi, mode in enumerate(("normal", "non_zero", "uniform")):
            a = graphs.copy_nums_to_probs(g.src, g.dst, g.base, cn, None, mode)
            b = O.seqgraph_to_phmm(g.src, g.dst, g.base, cn, None, mi)
            for x, y in zip(a, b):
                assert np.array_equal(np.isneginf(x), np.isneginf(y))
                fin = ~np.isneginf(y)
                assert np.allclose(x[fin], y[fin], rtol=0, atol=1e-15)


def test_seq_graph_edge_copy_num_kat_edge_freqs():
    """graph/seq_graph.rs:440-504 (seq_graph_edge_copy_num): full probabilities on mock_crossing with and without edge copy
    numbers and PHMMOutput::to_edge_freqs (freq.rs:276-315) on the junction edges 36..39."""
    from dbgphmm_b200 import graphs
    ra, rb = b"ATTAGGAGCA", b"ATTAGGAGCAGCTGATAGGG"
    outs = {}
    for flag in (False, True):
        sg = graphs.mock_crossing(flag)
        li, lt = sg.to_probs("normal")
        m = O.PHMMModel(sg.src, sg.dst, sg.base, li, lt, O.params_uniform(0.01))   # PHMMParams::default() (params.rs:126-128)
        outs[flag] = (m, m.run(ra), m.run(rb))
    (g1, o1a, o1b), (g2, o2a, o2b) = outs[False], outs[True]
    for o in (o1a, o1b, o2a, o2b):
        assert abs(o.to_full_prob_forward() - o.to_full_prob_backward()) < 0.1
    assert abs(o1a.to_full_prob_forward() - o2a.to_full_prob_forward()) < 0.1
    assert o1b.to_full_prob_forward() > -17.0
    assert o2b.to_full_prob_forward() < -39.0
    ef1 = o1b.to_edge_freqs(g1, rb)
    assert len(ef1) == 40
    assert ef1[36] < 0.0001 and ef1[37] > 0.9 and ef1[38] < 0.0001 and ef1[39] < 0.0001
    ef2 = o2b.to_edge_freqs(g2, rb)
    assert ef2[37] == 0.0 and ef2[38] == 0.0
    # every path leaves Begin exactly once (up to the Begin -> Ins mass)
    _, nf = o1b.to_edge_and_init_freqs(g1, rb)
    assert abs(nf.sum() - 1.0) < 1e-2


@pytest.mark.parametrize("flag", [True, False])
def test_hmm_crossing_edge_on_off(flag):
    """hmmv2/common.rs:381-417 (hmm_crossing_edge_on / _off): mock_crossing -> to_seq_graph -> to_phmm has 40 nodes and 40 edges;
    with edge copy numbers the transitions 9 -> 20 and 19 -> 30 are 1 and the crossing ones are zero, without them all four are
    1/2 (seq_graph.rs:180-211).  Checked on the product-side builder and on the oracle's restatement."""
    from dbgphmm_b200 import graphs
    sg = graphs.mock_crossing(flag)
    assert sg.n_nodes == 40 and sg.n_edges == 40
    for which, (li, lt) in (("graphs", sg.to_probs("normal")), ("oracle", O.seqgraph_to_phmm(sg.src, sg.dst, sg.base, sg.node_copy_num, sg.edge_copy_num, 0))):
        t = {(int(s), int(d)): float(p) for s, d, p in zip(sg.src, sg.dst, lt)}
        assert sorted(d for (s, d) in t if s == 9) == [20, 30] and sorted(s for (s, d) in t if d == 20) == [9, 19], which
        if flag:
            assert abs(t[(9, 20)]) < 1e-12 and abs(t[(19, 30)]) < 1e-12, which
            assert t[(9, 30)] == -np.inf and t[(19, 20)] == -np.inf, which
        else:
            for e in ((9, 20), (9, 30), (19, 20), (19, 30)):
                assert abs(t[e] - np.log(0.5)) < 1e-12, (which, e)


def _load_fixture_module():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_oracle_fixture.py")
    spec = importlib.util.spec_from_file_location("make_oracle_fixture", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_oracle_reproduces_its_committed_fixture():
    """tests/golden/oracle_c2_small.json (scaled-down C2: sparse mode, n_active_nodes = 40, n_warmup = k) is what the GPU parity
    test compares against as well; here the oracle must still produce it (index sets exactly, values to 1e-12)."""
    import json
    import os
    fx = _load_fixture_module()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_c2_small.json")) as fh:
        gold = json.load(fh)
    w, par, o = fx.build()
    assert w.graph.n_nodes == gold["n_nodes"] and w.graph.n_edges == gold["n_edges"] and [len(r) for r in w.reads] == gold["read_lens"]
    fr, lf, lb = o.run_node_freqs(O.Reads(w.reads), "sparse")
    assert np.allclose(lf, gold["logp_forward"], rtol=1e-12, atol=0) and np.allclose(lb, gold["logp_backward"], rtol=1e-12, atol=0)
    assert abs(fr.sum() - gold["node_freq_sum"]) < 1e-9 * gold["node_freq_sum"]
    for i, v in gold["node_freq_top"]:
        assert abs(fr[i] - v) <= 1e-12 * max(1.0, abs(v))
    for ri in (0, 1):
        f, b = o.forward_sparse(w.reads[ri], False), o.backward_sparse(w.reads[ri])
        for g in (g for g in gold["rows"] if g["read"] == ri):
            fr_, br_ = f.row(g["row"]), b.row(g["bwd_row"])
            assert fr_.is_dense == g["fwd_is_dense"] and br_.is_dense == g["bwd_is_dense"]
            assert sorted(int(x) for x in fr_.ids) == g["fwd_ids"] and sorted(int(x) for x in fr_.ids_d) == g["fwd_ids_d"]
            assert sorted(int(x) for x in br_.ids) == g["bwd_ids"] and sorted(int(x) for x in br_.ids_d) == g["bwd_ids_d"]
            assert abs(fr_.e - g["fwd_e"]) <= 1e-12 * abs(g["fwd_e"]) and abs(br_.mb - g["bwd_mb"]) <= 1e-12 * abs(g["bwd_mb"])
