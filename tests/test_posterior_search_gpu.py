"""GPU tests written after the round's GPU budget was spent: they only compose calls the rest of the suite
covers on the device, and the same compositions pass on the CPU with the oracle answering the model calls
(tests/test_score_host.py).  Kept in a file that sorts after the other GPU tests so that `pytest -x` reaches them last.

  * BASELINE configs[2] at its full size (N = 1.33 M nodes): size-independent properties of run_sparse + node frequencies;
  * MultiDbg::to_score for a batch of candidates (posterior.rs:259-277) through the Python layer, against the oracle;
  * the greedy search of MultiDbg::sample_posterior (posterior.rs:314-420) through the C++ host layer on toy::repeat."""
import numpy as np
import pytest

from dbgphmm_b200 import graphs, hmmv2 as H
from tests.test_cpp_header import RC_NO_DEVICE, _build_and_run
from tests.test_formats import README_DBG

pytestmark = pytest.mark.gpu

CPP = r'''#include "dbgphmm_b200.hpp"
#include <cstdio>
#define CHECK(cond, code) do { if (!(cond)) { std::fprintf(stderr, "check failed (%d): %s\n", code, #cond); return code; } } while (0)
int main() {
    if (dbgphmm_device_count() == 0) return ''' + str(RC_NO_DEVICE) + r''';
    // MultiDbg::to_score / sample_posterior over toy::repeat (README.md:174-191) with +-1 steps on the repeat edge as the neighbours:
    // reads from the 3-unit genome, start at 1 unit -> the search ends on 3 units (tests/test_score_host.py does the same over the oracle)
    {
        auto toy = dbgphmm::MultiDbg::from_dbg_str("K\t4\nN\t0\tnnn\nN\t1\tCAG\nE\t0\t1\t0\tCAGGAAnnn\t1\t9,10,11,12,13,14\n"
                                                  "E\t1\t1\t1\tCAGCAG\t3\t6,7,8\nE\t2\t0\t1\tnnnTCCCAG\t1\t0,1,2,3,4,5\n");
        auto tp = toy->to_phmm(dbgphmm::uniform(0.01));
        const std::string genome = "TCCCAGCAGCAGCAGGAA";
        dbgphmm::Reads rs(std::vector<std::string>{genome, genome, genome, genome, genome.substr(2, 12), genome.substr(5)});
        toy->set_copy_nums({1, 1, 1});
        auto nb = [](const dbgphmm::MultiDbg& g) {
            std::vector<uint32_t> c = g.get_copy_nums(), up = c, dn = c;
            up[1]++;
            std::vector<std::vector<uint32_t>> set{up};
            if (c[1] > 0) { dn[1]--; set.push_back(dn); }
            return std::vector<std::vector<std::vector<uint32_t>>>{set};
        };
        dbgphmm::Posterior found = toy->sample_posterior(*tp, rs, nullptr, 18, 3, nb, 10);
        CHECK(toy->get_copy_nums() == std::vector<uint32_t>({1, 1, 1}), 48);
        CHECK(found.max_copy_nums() == std::vector<uint32_t>({1, 3, 1}) && found.samples().size() == 5, 49);
        CHECK(found.p_edge_x(1, 3) > found.p_edge_x(1, 2), 50);
    }
    std::puts("posterior search ok");
    return 0;
}
'''


def test_to_scores_of_a_candidate_batch_against_the_oracle():
    from oracle import oracle as O
    d = H.MultiDbg.from_dbg_str(README_DBG)
    sg, k = graphs.toy_repeat()
    g = d.to_phmm(H.params_uniform(0.01))
    op = O.params_uniform(0.01); op.n_warmup = k
    o = O.PHMMModel(sg.src, sg.dst, sg.base, *sg.to_probs("normal"), op)
    reads = [b"TCCCAGCAGCAGCAGGAA", b"CCAGCAGG"]
    X = np.array([[1, 3, 1], [1, 2, 1], [1, 6, 1]], np.uint32)
    scores = d.to_scores(g, H.Reads(reads), None, X, 18, 3)
    for b in range(3):
        o.set_probs(*sg.to_probs("normal", d.expand_copy_nums(X[b])[0]))
        s, _ = o.to_full_prob_reads(O.Reads(reads), None, True)         # to_likelihood uses use_max_ratio = true (posterior.rs:247-255)
        assert abs(scores[b].likelihood - s) <= 1e-9 * abs(s)
        assert scores[b].genome_size == [18, 15, 27][b] and abs(scores[b].n_euler_circuits) < 1e-12
        assert abs(scores[b].p() - (s - 0.5 * np.log(2 * np.pi * 9.0) - ([18, 15, 27][b] - 18) ** 2 / 18.0)) <= 1e-9 * abs(s)


def test_cpp_sample_posterior_walks_to_the_true_copy_number(tmp_path):
    rc, out = _build_and_run(tmp_path, CPP)
    assert rc == 0 and "posterior search ok" in out, (rc, out)


@pytest.fixture(scope="module")
def c3_full():
    """BASELINE.json configs[2] at its FULL size (1 Mbp diploid, 1 % het, k = 40: N = 1.33 M nodes, 10 kbp reads; the graph bench.py
    measures) and a handful of reads."""
    from dbgphmm_b200 import synth
    h0 = synth.random_genome(1_000_000, 0)
    h1 = synth.mutate_substitutions(h0, 0.01, 1)
    sg, _ = graphs.build_dbg([h0.tobytes(), h1.tobytes()], 40, seed=100)
    reads = synth.sample_reads([h0, h1], 0.03, 10_000, 0.001, 1000)[:6]
    assert len(reads) == 6 and sg.n_nodes > 1_300_000
    par = H.params_uniform(0.001); par.n_warmup = 40
    li, lt = sg.to_probs("normal")
    return sg, reads, H.PHMMModel(sg.src, sg.dst, sg.base, li, lt, par)


def test_c3_full_size_properties(c3_full, monkeypatch):
    """Size-independent properties instead of the oracle (one full read costs ~100 CPU-seconds there), on the path bench.py times
    (stream strategy).  The same bounds hold against the oracle at 50 kbp in
    tests/test_gpu_configs.py::test_c3_scaled_diploid_stream_strategy_properties."""
    sg, reads, g = c3_full
    monkeypatch.setenv("DBGPHMM_STRATEGY", "stream")
    fs, lfs, lbs, cs = g.run_node_freqs(H.Reads(reads), "sparse")
    # cells: 2 x 40 dense warm-up rows of N nodes per read dominate
    assert cs[0] > 6 * 40 * sg.n_nodes and cs[1] > 6 * 40 * sg.n_nodes
    n_bases = sum(len(r) for r in reads)
    assert np.isfinite(lfs).all() and (lfs < 0).all() and (lfs > -0.1 * 10_000).all()       # HiFi reads of the genome itself: likely
    assert abs(fs.sum() / n_bases - 1.0) < 5e-3              # every base is emitted by exactly one Match / Ins state (+ ~0.1 % silent Del mass)
    assert np.allclose(lfs, lbs, rtol=1e-3)                  # forward and backward totals differ only by the bounded Del chain
    # reads are independent: a batch is the sum of its parts, whatever the batching
    fa, lfa, _, _ = g.run_node_freqs(H.Reads(reads[:2]), "sparse")
    fb, lfb, _, _ = g.run_node_freqs(H.Reads(reads[2:]), "sparse")
    assert np.allclose(fa + fb, fs, rtol=1e-9, atol=1e-12) and np.allclose(np.concatenate([lfa, lfb]), lfs, rtol=1e-12, atol=0)
    # the expected frequencies follow the reads: the nodes a read passes through carry it
    assert (fs > 0.5).sum() > 0.9 * n_bases / 2 and fs.max() < len(reads) + 0.5


def test_c3_full_size_strategy_independence(c3_full, monkeypatch):
    """Which rows are kept (store: every dense row of the batch, single-row kernels; stream: two ping-pong slabs per read, two rows
    per launch, products by cone recompute) does not change the result -- at full size."""
    sg, reads, g = c3_full
    res = {}
    for strat in ("stream", "store"):
        monkeypatch.setenv("DBGPHMM_STRATEGY", strat)
        res[strat] = g.run_node_freqs(H.Reads(reads[:3]), "sparse")
    fs, lfs, lbs, cs = res["stream"]
    ft, lft, lbt, ct = res["store"]
    assert cs == ct and np.allclose(lfs, lft, rtol=1e-12, atol=0) and np.allclose(lbs, lbt, rtol=1e-12, atol=0)
    assert np.allclose(fs, ft, rtol=1e-9, atol=1e-12)
