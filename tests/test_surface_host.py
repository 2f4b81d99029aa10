"""CPU checks of the host-side compositions of the Python mirror (dbgphmm_b200/hmmv2.py) that complete the reference's
method surface: PHMMOutput::{to_emit_probs, iter_emit_probs, to_state_probs} and PHMMModel::{to_node_freqs,
to_full_prob, to_full_prob_parallel, to_full_prob_sparse, to_full_prob_sparse_backward, forward_*_score_only}.

The compositions only call methods that exist on the GPU objects AND on the oracle's objects (row / len / full_prob; the two
bulk calls), so here they are run over the oracle -- the checker standing in for the device -- and compared with what the
oracle computes the way the reference does: per read, table by table (freq.rs:87-165,226-255, table.rs:414-434,500-505).
The same compositions over the CUDA path are compared with the oracle in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from dbgphmm_b200 import graphs, synth
from dbgphmm_b200 import hmmv2 as H
from oracle import oracle as O
from tests.common import oracle_model, oracle_params


class OracleBacked(H.PHMMModel):
    """hmmv2.PHMMModel whose two bulk C calls are answered by the oracle: what the compositions are built on."""

    def __init__(self, o):            # no device handle
        self._o, self._h, self.n_nodes = o, None, o.n_nodes

    def __del__(self):
        pass

    @staticmethod
    def _oreads(reads):
        return O.Reads([reads[r] for r in range(len(reads))])

    @staticmethod
    def _omaps(mappings):
        return None if mappings is None else O.Mappings(mappings.read_off, mappings.row_off, mappings.nodes, mappings.probs)

    def run_node_freqs(self, reads, mode, use_max_ratio=True, mappings=None, want_freqs=True):
        fr, lf, lb = self._o.run_node_freqs(self._oreads(reads), mode, use_max_ratio, self._omaps(mappings), want_freqs=want_freqs)
        return fr, lf, lb, (0, 0)

    def to_full_prob_reads(self, reads, mappings=None, use_max_ratio=True):
        s, per = self._o.to_full_prob_reads(self._oreads(reads), self._omaps(mappings), use_max_ratio)
        return np.array([s]), per[None, :]


@pytest.fixture(scope="module")
def case():
    w = synth.make_workload("surface", 600, 16, 4, 150, 0.003, ploidy=2, het=0.02, seed=8, n_reads=5)
    par = oracle_params(0.001, n_warmup=w.k, warmup_threshold=30)
    o = oracle_model(w.graph, par, "non_zero")
    return w, o


def test_emit_and_state_probs_reproduce_the_node_frequencies(case):
    w, o = case
    x = w.reads[0]
    for out in (o.run(x), o.run_sparse(x), o.run_sparse_adaptive(x, True)):
        sp = H.state_probs(out.forward, out.backward)
        with np.errstate(over="ignore"):
            freqs = np.exp(sp.merged(o.n_nodes))
        ref = out.to_node_freqs()            # the oracle's own to_node_freqs (freq.rs:245-255)
        assert np.allclose(freqs, ref, rtol=1e-9, atol=1e-13), np.abs(freqs - ref).max()
    # dense run: every base is emitted by exactly one Match or Ins state -> the posteriors of the emitting states add up to 1
    # up to the difference between the forward and the backward total (bounded Del chains, SURVEY §8a gotcha 10)
    out = o.run(x)
    n = len(out.forward)
    for i in (1, n // 2, n):
        t = H.emit_probs(out.forward, out.backward, i)
        tot = np.logaddexp(np.logaddexp.reduce(t.m), np.logaddexp.reduce(t.i))
        assert abs(tot) < 1e-2, (i, tot)
    # merged index 0 / n use the init tables (table.rs:414-434)
    t0 = H.emit_probs(out.forward, out.backward, 0)
    assert np.isneginf(t0.m).all() and abs(t0.mb - (out.backward.row(0).mb - out.forward.full_prob())) < 1e-12
    tn = H.emit_probs(out.forward, out.backward, n)
    assert np.allclose(tn.m, out.forward.row(n - 1).m + out.backward.row(-1).m - out.forward.full_prob(), rtol=0, atol=1e-12, equal_nan=True)


def test_table_merged_indexing():
    o = oracle_model(graphs.mock_linear(), oracle_params(0.1))
    f, b = o.forward(b"CGATC"), o.backward(b"CGATC")
    assert H.table_merged(f, True, 0).mb == f.row(-1).mb == 0.0
    assert H.table_merged(f, True, 5).e == f.row(4).e == f.full_prob()
    assert H.table_merged(b, False, 0).mb == b.row(0).mb == b.full_prob()
    assert np.array_equal(H.table_merged(b, False, 5).m, b.row(-1).m)
    assert np.array_equal(H.table_merged(b, False, 2).m, b.row(2).m)


def test_read_set_methods_follow_the_reference_definitions(case):
    w, o = case
    g = OracleBacked(o)
    seqs = w.reads
    # freq.rs:87-102
    want = sum(o.run(x).to_node_freqs() for x in seqs)
    assert np.allclose(g.to_node_freqs(seqs), want, rtol=1e-12, atol=1e-15)
    # freq.rs:105-135
    want = sum(o.forward(x).full_prob() for x in seqs)
    assert abs(g.to_full_prob(seqs) - want) < 1e-9 and abs(g.to_full_prob_parallel(H.Reads(seqs)) - want) < 1e-9
    # freq.rs:138-150
    for ratio in (False, True):
        want = sum(o.forward_sparse(x, ratio).full_prob() for x in seqs)
        assert abs(g.to_full_prob_sparse(seqs, ratio) - want) < 1e-9
        assert abs(g.forward_sparse_score_only(seqs[1], ratio) - o.forward_sparse(seqs[1], ratio).full_prob()) < 1e-10
    # freq.rs:153-164
    want = sum(o.backward_sparse(x).full_prob() for x in seqs)
    assert abs(g.to_full_prob_sparse_backward(seqs) - want) < 1e-9
    # forward.rs:79-89
    om = o.generate_mappings(O.Reads(seqs), None, True)
    hm = H.Mappings(om.read_off, om.row_off, om.nodes, om.probs)
    for r in (0, 3):
        want = o.forward_with_mapping(seqs[r], om[r]).full_prob()
        assert abs(g.forward_with_mapping_score_only(seqs[r], hm, r) - want) < 1e-10


def test_emit_probs_of_an_impossible_read_fail_like_the_reference():
    o = oracle_model(graphs.mock_linear(), oracle_params(0.0))
    f, b = o.forward(b"CGATT"), o.backward(b"CGATT")     # P = 0 with zero error rates (forward.rs:592-596)
    assert np.isneginf(f.full_prob())
    with pytest.raises(H.DbgphmmError):
        H.emit_probs(f, b, 1)


def test_table_diff_and_log_diff_like_the_reference_tests(case):
    """PHMMTable::diff / log_diff (table.rs:174-195) as the reference's own tests use them: dense vs sparse rows are identical inside
    the warm-up (`diff == 0.0`, tests/hmm.rs:155-173) and close after it (forward.rs:621-638 asserts < 1e-9 on its linear mock)."""
    w, o = case
    x = w.reads[2]
    N = o.n_nodes
    dense, sparse = o.forward(x), o.forward_sparse(x, False)
    for r in range(len(dense)):
        a, b = dense.row(r), sparse.row(r)
        d = H.table_diff(a, b, N)
        if b.is_dense:
            assert d == 0.0 and H.table_log_diff(a, b, N) == 0.0
        else:
            assert d < 1e-6 and np.isinf(H.table_log_diff(a, b, N))      # absent sparse entries are zeros: ln-diff is infinite
    assert not sparse.row(len(dense) - 1).is_dense
    # hand-made rows: |p_a - p_b| summed over states and the three scalars
    a, b = H.Row(), H.Row()
    a.is_dense = b.is_dense = False
    a.ids = np.array([1, 3], np.uint32); a.m = np.log([0.5, 0.25]); a.i = np.log([0.125, 1e-300]); a.ids_d = np.array([3], np.uint32); a.d = np.log([0.0625])
    b.ids = np.array([3], np.uint32); b.m = np.log([0.25]); b.i = np.array([-np.inf]); b.ids_d = np.zeros(0, np.uint32); b.d = np.zeros(0)
    a.mb, a.ib, a.e = np.log(0.5), -np.inf, np.log(0.1); b.mb, b.ib, b.e = np.log(0.25), -np.inf, np.log(0.1)
    assert abs(a.diff(b, 5) - (0.5 + 0.125 + 1e-300 + 0.0625 + 0.25)) < 1e-15
    assert np.isinf(a.log_diff(b, 5)) and a.log_diff(a, 5) == 0.0 and a.n_active_nodes() == 2
    q = H.QScore(-1.0, -2.0, 0.0).sub(H.QScore(-0.5, -1.0, 0.0))
    assert (q.init, q.trans, q.total()) == (-0.5, -1.0, -1.5)
