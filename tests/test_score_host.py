"""Host-side terms of MultiDbg::to_score beside the likelihood (multi_dbg/posterior.rs:164-277): Euler-circuit count, genome size,
prior, Score / Posterior bookkeeping.  Pinned on the reference's own known answers: graph/euler.rs:142-238 (`n_euler`),
multi_dbg.rs:2320-2351 (`n_euler_circuits_test_toy`); the log-determinant is cross-checked against numpy on random multigraphs."""
import math

import numpy as np
import pytest

from dbgphmm_b200 import hmmv2 as H
from tests.test_formats import README_DBG

# toy::one_in_n_repeat (multi_dbg/toy.rs:450-493) as a DBG text: compact nodes nn, gg, ga, gt; the copy numbers of its full edges
ONE_IN_N_REPEAT = ("K\t3\nN\t0\tnn\nN\t1\tGG\nN\t2\tGA\nN\t3\tGT\n"
                   "E\t0\t0\t1\tnnGG\t1\t0,1\n"
                   "E\t1\t1\t2\tGGA\t5\t2\n"
                   "E\t2\t2\t3\tGACGT\t4\t3,4,5\n"
                   "E\t3\t3\t1\tGTGG\t4\t6,7\n"
                   "E\t4\t2\t3\tGAAGT\t1\t8,9,10\n"
                   "E\t5\t3\t0\tGTnn\t1\t11,12\n")


@pytest.mark.parametrize("edges,n,n0,n1", [
    ([(0, 0, 1)], 1, 1, 1),                                                            # self loop
    ([(0, 1, 2)], 2, 0, 0),                                                            # single edge has no euler circuit
    ([(0, 1, 1), (1, 0, 1)], 2, 1, 1),                                                 # loop
    ([(0, 0, 1), (0, 0, 1), (0, 0, 2)], 1, 3, 3),                                      # three euler circuits: AXXB AXBX ABXX
    ([(0, 1, 1), (0, 1, 1), (1, 2, 2), (2, 3, 1), (2, 3, 1), (3, 0, 2)], 4, 2, 2),     # two bubbles
    ([(0, 0, 1), (1, 1, 1)], 2, 0, 1),                                                 # separate components
])
def test_euler_circuit_count_reference_known_answers(edges, n, n0, n1):
    # graph/euler.rs:142-238: assert_euler_approx_eq(graph, n0 (single component required), n1 (components multiplied))
    assert abs(math.exp(H.euler_circuit_count(n, edges, False)) - n0) < 1e-3
    assert abs(math.exp(H.euler_circuit_count(n, edges, True)) - n1) < 1e-3


def test_n_euler_circuits_toy():
    # multi_dbg.rs:2320-2342
    d = H.MultiDbg.from_dbg_str(README_DBG)               # toy::repeat
    assert abs(math.exp(d.n_euler_circuits()) - 1.0) < 1e-4
    d = H.MultiDbg.from_dbg_str(ONE_IN_N_REPEAT)
    assert abs(math.exp(d.n_euler_circuits()) - 5.0) < 1e-4
    zeros = np.zeros(d.n_edges_compact, np.uint32)
    cands = np.stack([d.get_copy_nums(), zeros, [2, 6, 4, 4, 2, 2]])
    n = d.n_euler_circuits(cands)
    assert abs(math.exp(n[0]) - 5.0) < 1e-4 and math.exp(n[1]) == 0.0 and np.isfinite(n[2])
    d.set_copy_nums(zeros)
    assert math.exp(d.n_euler_circuits()) == 0.0
    with pytest.raises(H.DbgphmmError):                   # copies do not balance: set_copy_nums would assert (multi_dbg.rs:1051)
        d.n_euler_circuits(np.array([[1, 5, 4, 4, 1, 2]], np.uint32))


def test_genome_size_and_prior():
    d = H.MultiDbg.from_dbg_str(ONE_IN_N_REPEAT)
    assert d.genome_size() == 30                          # 1 + 1 + 5 + 5 x 4 + 3 x 1 ; the two n-edges do not count (multi_dbg.rs:1018-1028)
    cands = np.stack([d.get_copy_nums(), [2, 6, 4, 4, 2, 2]])
    assert list(d.genome_size(cands)) == [30, 2 * 2 + 6 + 12 + 8 + 6]
    # distribution.rs:22-25
    want = -0.5 * math.log(2 * math.pi * 25.0) - (30 - 28) ** 2 / (2 * 25.0)
    assert abs(d.to_prior(28, 5) - want) < 1e-15
    assert abs(d.to_prior(28, 5, cands)[1] - (-0.5 * math.log(2 * math.pi * 25.0) - (36 - 28) ** 2 / 50.0)) < 1e-15


def _reference_formula(n, edges):
    """euler_circuit_count_in_connected (graph/euler.rs:22-84) with numpy's slogdet, for a strongly connected balanced multigraph"""
    L = np.zeros((n, n)); out = np.zeros(n, np.int64)
    for s, t, w in edges:
        L[s, s] += w; L[s, t] -= w; out[s] += w
    L[0, 0] += 1.0
    sign, ln = np.linalg.slogdet(L)
    c = sign * ln
    for v in range(n):
        c += math.lgamma(out[v])           # ln (out - 1)!
    for _, _, w in edges:
        c -= math.lgamma(w + 1)
    return c


def test_log_determinant_against_numpy_on_random_balanced_multigraphs():
    rng = np.random.default_rng(7)
    for trial in range(20):
        n = int(rng.integers(2, 40))
        # a sum of random closed walks is balanced and (through the Hamiltonian cycle added first) strongly connected
        mult = {}
        perm = rng.permutation(n)
        walks = [list(perm) + [perm[0]]] + [list(rng.integers(0, n, int(rng.integers(2, 8)))) for _ in range(int(rng.integers(0, 30)))]
        for wk in walks:
            if wk[0] != wk[-1]:
                wk = wk + [wk[0]]
            for a, b in zip(wk[:-1], wk[1:]):
                mult[(int(a), int(b))] = mult.get((int(a), int(b)), 0) + 1
        edges = [(a, b, w) for (a, b), w in mult.items()]
        if trial % 3 == 0:                  # parallel edges given separately add up like one edge with the summed multiplicity in L ...
            a, b, w = edges[0]
            edges[0] = (a, b, w + 1); edges.append((a, b, 1)); edges.append((b, a, 2)) if a != b else None
            edges = [e for e in edges if e is not None]
            # ... re-balance: the extra a -> b copies return over b -> a
        got = H.euler_circuit_count(n, edges, False)
        want = _reference_formula(n, edges)
        bal = np.zeros(n, np.int64)
        for s, t, w in edges:
            bal[s] -= w; bal[t] += w
        if (bal != 0).any():
            assert np.isneginf(got)
        else:
            assert abs(got - want) < 1e-9 * max(1.0, abs(want)), (trial, got, want)


def test_score_and_posterior_bookkeeping():
    s1 = H.Score(-100.0, -3.0, 30, math.log(5.0)); s2 = H.Score(-101.0, -2.5, 31, 0.0); s3 = H.Score(-100.0, -3.0, 30, math.log(5.0))
    assert s1.p() == -103.0 + math.log(5.0)
    post = H.Posterior()
    assert np.isneginf(post.p())
    post.add([1, 5, 4], s1); post.add([1, 4, 4], s2); post.add([1, 5, 4], s3)        # the third is a duplicate: ignored (posterior.rs:93-98)
    assert len(post.samples) == 2 and post.contains([1, 4, 4]) and not post.contains([0, 0, 0]) and post.find([1, 5, 4]) is s1
    assert abs(post.p() - np.logaddexp(s1.p(), s2.p())) < 1e-12
    assert list(post.max_copy_nums()) == [1, 5, 4]
    # P(X[1] = 5 | R) = p(s1) / (p(s1) + p(s2)) ; edge 0 has copy number 1 in every sample
    assert abs(post.p_edge_x(1, 5) - (s1.p() - post.p())) < 1e-12 and abs(post.p_edge_x(0, 1)) < 1e-12 and np.isneginf(post.p_edge_x(2, 7))


def test_to_scores_batches_the_candidates_through_one_model_call():
    """MultiDbg::to_score for a batch (posterior.rs:259-277,504-515): the composition calls expand -> set_copy_nums_batch ->
    to_full_prob_reads once each; checked here with a stand-in model (the two device calls are covered by the GPU parity tests)."""
    d = H.MultiDbg.from_dbg_str(ONE_IN_N_REPEAT)
    calls = []

    class Model:
        def set_copy_nums_batch(self, full, mode):
            calls.append(("set", full.copy(), mode))

        def to_full_prob_reads(self, reads, mappings, use_max_ratio):
            calls.append(("score", reads, mappings, use_max_ratio))
            return np.array([-10.0, -20.0]), None

    cands = np.stack([d.get_copy_nums(), [2, 6, 4, 4, 2, 2]])
    scores = d.to_scores(Model(), "reads", "maps", cands, 28, 5)
    assert [c[0] for c in calls] == ["set", "score"] and calls[1][1:] == ("reads", "maps", True) and calls[0][2] == "normal"
    full = calls[0][1]
    assert full.shape == (2, d.n_edges_full) and list(full[0][:3]) == [1, 1, 5] and list(full[1][8:11]) == [2, 2, 2]
    assert [s.likelihood for s in scores] == [-10.0, -20.0] and [s.genome_size for s in scores] == [30, 36]
    assert abs(math.exp(scores[0].n_euler_circuits) - 5.0) < 1e-4
    assert abs(scores[0].p() - (-10.0 + d.to_prior(28, 5) + math.log(5.0))) < 1e-9


def test_sample_posterior_greedy_search_over_the_oracle():
    """MultiDbg::sample_posterior (posterior.rs:314-420) with a caller-supplied neighbour generator, end to end on toy::repeat with
    real likelihoods: the model calls are answered by the oracle (the checker standing in for the device, as in
    tests/test_surface_host.py).  Reads come from the 3-unit genome of the toy; the search starts from 1 unit and walks +-1 steps."""
    from dbgphmm_b200 import graphs
    from oracle import oracle as O
    d = H.MultiDbg.from_dbg_str(README_DBG)
    sg, k = graphs.toy_repeat()
    op = O.params_uniform(0.01); op.n_warmup = k
    li, lt = sg.to_probs("normal")
    o = O.PHMMModel(sg.src, sg.dst, sg.base, li, lt, op)
    genome = b"TCCCAGCAGCAGCAGGAA"
    reads = O.Reads([genome] * 4 + [genome[2:14], genome[5:]])
    evaluated = []

    class OracleModel:
        def set_copy_nums_batch(self, full, mode):
            self.full, self.mode = full, mode

        def to_full_prob_reads(self, rd, mappings, use_max_ratio):
            out = []
            for x in self.full:
                o.set_probs(*sg.to_probs(self.mode, x))
                out.append(o.to_full_prob_reads(rd, mappings, use_max_ratio)[0])
                evaluated.append(tuple(int(v) for v in x[6:9]))
            return np.array(out), None

    def neighbors(dbg):
        c = dbg.get_copy_nums()
        up = c.copy(); up[1] += 1
        out = [up]
        if c[1] > 0:
            dn = c.copy(); dn[1] -= 1
            out.append(dn)
        return [out]

    d.set_copy_nums(np.array([1, 1, 1], np.uint32))
    post = d.sample_posterior(OracleModel(), reads, None, 18, 3, neighbors, max_iter=10)
    assert list(d.get_copy_nums()) == [1, 1, 1]                        # self untouched
    seen = [s[0] for s in post.samples]
    assert seen[0] == (1, 1, 1) and len(set(seen)) == len(seen)          # every vector scored once (posterior.rs:93-98, 507-511)
    assert len(evaluated) == len(seen)
    best = post.max_sample()
    assert best[0] == (1, 3, 1), (best, seen)                            # the truth: three units
    # the walk went 1 -> 2 -> 3, looked at 4 and 2 from there and stopped
    assert set(seen) == {(1, 1, 1), (1, 2, 1), (1, 0, 1), (1, 3, 1), (1, 4, 1)}
    for key, sc in post.samples:
        o.set_probs(*sg.to_probs("normal", d.expand_copy_nums(np.array([key], np.uint32))[0]))
        want = o.to_full_prob_reads(reads, None, True)[0]
        assert abs(sc.likelihood - want) < 1e-12 and sc.genome_size == 9 + 3 * key[1]
    assert abs(post.p() - np.logaddexp.reduce([sc.p() for _, sc in post.samples])) < 1e-9
    assert post.p_edge_x(1, 3) > post.p_edge_x(1, 2) > post.p_edge_x(1, 0)


def _dbg_text_of_genome(genome, k):
    """DBG text (multi_dbg/output.rs:155-199) of one linear genome with n-padded ends: (k-1)-mer nodes, k-mer edges with their
    multiplicity as copy number, compact edges = maximal simple paths between nodes that branch (or the all-n terminal)."""
    s = "n" * (k - 1) + genome + "n" * (k - 1)
    kmers = {}
    for j in range(len(s) - k + 1):
        kmers[s[j:j + k]] = kmers.get(s[j:j + k], 0) + 1
    out_e, in_e = {}, {}
    for km in kmers:
        out_e.setdefault(km[:-1], []).append(km); in_e.setdefault(km[1:], []).append(km)
    nodes = sorted(set(out_e) | set(in_e))
    is_compact = lambda v: v == "n" * (k - 1) or len(out_e.get(v, [])) != 1 or len(in_e.get(v, [])) != 1
    cnodes = [v for v in nodes if is_compact(v)]
    cid = {v: i for i, v in enumerate(cnodes)}
    full_id = {km: i for i, km in enumerate(sorted(kmers))}
    lines = [f"K\t{k}"] + [f"N\t{i}\t{v}" for i, v in enumerate(cnodes)]
    eid = 0
    for v in cnodes:
        for km in sorted(out_e.get(v, [])):
            path, seq, w = [km], km, km[1:]
            while not is_compact(w):
                km2 = out_e[w][0]
                path.append(km2); seq += km2[-1]; w = km2[1:]
            assert len({kmers[x] for x in path}) == 1
            lines.append(f"E\t{eid}\t{cid[v]}\t{cid[w]}\t{seq}\t{kmers[km]}\t" + ",".join(str(full_id[x]) for x in path))
            eid += 1
    return "\n".join(lines) + "\n"


def test_multi_move_accepts_independent_improvements_at_once():
    """posterior.rs:533-588 (the mode of rescue-only rounds): two tandem repeats, both under-counted; the +1 moves on the two repeat
    edges both improve the score and touch different edges, so they are accepted together and the combined vector is scored in the
    same round; a third neighbour that shares an edge with an accepted one is left out.  Likelihoods from the oracle."""
    from dbgphmm_b200 import graphs
    from oracle import oracle as O
    k = 4
    genome = "TCC" + "CAG" * 3 + "GAATACT" + "TGA" * 3 + "CCGT"
    d = H.MultiDbg.from_dbg_str(_dbg_text_of_genome(genome, k))
    src, dst, em, cn, ce = d.phmm_graph()
    truth = d.get_copy_nums()
    loops = [e for e in range(d.n_edges_compact) if truth[e] == 2]        # the two self-loop edges CAG -> CAG and TGA -> TGA
    assert len(loops) == 2 and d.genome_size() == len(genome)
    op = O.params_uniform(0.01); op.n_warmup = k
    li, lt = graphs.copy_nums_to_probs(src, dst, em, cn, None, "normal")
    o = O.PHMMModel(src, dst, em, li, lt, op)
    reads = O.Reads([genome.encode()] * 5)

    class OracleModel:
        def set_copy_nums_batch(self, full, mode):
            self.full, self.mode = full, mode

        def to_full_prob_reads(self, rd, mappings, use_max_ratio):
            out = []
            for x in self.full:
                o.set_probs(*graphs.copy_nums_to_probs(src, dst, em, x, None, self.mode))
                out.append(o.to_full_prob_reads(rd, mappings, use_max_ratio)[0])
            return np.array(out), None

    start = truth.copy(); start[loops] = 1
    d.set_copy_nums(start)
    up = lambda *es: np.array([start[e] + (1 if e in es else 0) for e in range(len(start))], np.uint32)
    both_plus_a_again = up(loops[0]); both_plus_a_again[loops[0]] += 1          # +2 on loop A: shares the edge with the +1 move
    neighbors = [up(loops[0]), up(loops[1]), both_plus_a_again]
    post = H.Posterior()
    post.add(start, d.to_scores(OracleModel(), reads, None, [start], len(genome), 3)[0])
    best = d.sample_posterior_once(OracleModel(), reads, None, neighbors, post, len(genome), 3, multi_move=True)
    seen = [s[0] for s in post.samples]
    assert tuple(truth) in seen and len(seen) == 5                             # start, 3 neighbours, the combined move
    assert best is not None and best[0] == tuple(int(v) for v in truth)        # both repeats fixed in one round
    # without multi-move the same round only gets as far as the better single move
    post1 = H.Posterior()
    post1.add(start, d.to_scores(OracleModel(), reads, None, [start], len(genome), 3)[0])
    best1 = d.sample_posterior_once(OracleModel(), reads, None, neighbors, post1, len(genome), 3)
    assert tuple(truth) not in [s[0] for s in post1.samples] and best1[0] in (tuple(up(loops[0])), tuple(up(loops[1])))
    with pytest.raises(H.DbgphmmError):                                        # "current copy number was not sampled"
        d.sample_posterior_once(OracleModel(), reads, None, neighbors, H.Posterior(), len(genome), 3, multi_move=True)
