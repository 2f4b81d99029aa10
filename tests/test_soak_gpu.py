"""Soak of the bulk path on the GPU: many passes of run_sparse + node frequencies in the stream strategy over a reduced bench shard,
two model handles alternating, every stored row checked structurally (DBGPHMM_VERIFY=1), every pass bit-identical to the first.

Round 1 shipped a race in the sparse-row store (a warp could write its entries at the next row's offset) that only showed up as an
intermittent CUDA error 700 after tens of steps; this test is the regression guard: it fails within a few passes on that code."""
import numpy as np
import pytest

from dbgphmm_b200 import synth

pytestmark = pytest.mark.gpu


def _workload(seed, n_reads):
    w = synth.make_workload("soak", 60_000, 40, 4, 1_500, 0.001, ploidy=2, het=0.01, seed=seed)
    reads = list(w.reads)
    while len(reads) < n_reads:
        reads += synth.sample_reads(w.haplotypes, 4, 1_500, 0.001, 900 + len(reads))
    return w, reads[:n_reads]


@pytest.mark.parametrize("cap", [None, "48"])
def test_soak_stream_strategy_two_handles(monkeypatch, cap):
    from dbgphmm_b200 import hmmv2 as H
    monkeypatch.setenv("DBGPHMM_STRATEGY", "stream")
    monkeypatch.setenv("DBGPHMM_VERIFY", "1")
    if cap:
        monkeypatch.setenv("DBGPHMM_SPARSE_CAP", cap)   # small primary tables: many jobs are handed to the rescue launch
    models, sets, refs = [], [], []
    for seed in (7, 8):
        w, reads = _workload(seed, 300)
        li, lt = w.graph.to_probs("normal")
        par = H.params_uniform(0.001); par.n_warmup = w.k
        models.append(H.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, li, lt, par))
        sets.append(H.Reads(reads))
        refs.append(None)
    n_pass = 60 if cap is None else 40
    for step in range(n_pass):
        for h in range(2):
            fr, lf, lb, cells = models[h].run_node_freqs(sets[h], "sparse")
            if refs[h] is None:
                refs[h] = (fr, lf, lb, cells)
                n_bases = sets[h].total_bases()
                assert abs(fr.sum() / n_bases - 1.0) < 1e-3
                continue
            rf, rlf, rlb, rc = refs[h]
            assert cells == rc, f"pass {step} handle {h}: cell count changed"
            assert np.array_equal(lf, rlf) and np.array_equal(lb, rlb), f"pass {step} handle {h}: ln P changed"
            # the frequencies are accumulated with f64 atomics: the order of the additions is not fixed
            assert np.allclose(fr, rf, rtol=1e-11, atol=1e-13), f"pass {step} handle {h}: node frequencies changed by {np.abs(fr - rf).max()}"


def test_two_thread_overlap_equals_one_phase_at_a_time(monkeypatch):
    """Stream strategy: both sparse phases side by side (second host thread, second stream set) against one phase at a time."""
    from dbgphmm_b200 import hmmv2 as H
    monkeypatch.setenv("DBGPHMM_STRATEGY", "stream")
    monkeypatch.setenv("DBGPHMM_VERIFY", "1")
    w, reads = _workload(11, 200)
    li, lt = w.graph.to_probs("normal")
    par = H.params_uniform(0.001); par.n_warmup = w.k
    m = H.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, li, lt, par)
    res = {}
    for mode in ("0", "1", "0", "1"):
        monkeypatch.setenv("DBGPHMM_OVERLAP", mode)
        fr, lf, lb, cells = m.run_node_freqs(H.Reads(reads), "sparse")
        if mode in res:
            rf, rlf, rlb, rc = res[mode]
            assert cells == rc and np.array_equal(lf, rlf) and np.array_equal(lb, rlb)
        res[mode] = (fr, lf, lb, cells)
    assert res["0"][3] == res["1"][3]
    assert np.array_equal(res["0"][1], res["1"][1]) and np.array_equal(res["0"][2], res["1"][2])
    assert np.allclose(res["0"][0], res["1"][0], rtol=1e-11, atol=1e-13)
