"""world_size-2 gloo test of the multi-GPU host logic (read sharding + one all-reduce) on CPU.

The compute callback here is the CPU oracle (allowed in tests); on the GPU box bench.py runs the same plumbing with the
CUDA library and NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dbgphmm_b200 import synth
    from dbgphmm_b200.dist import allreduce_results, shard_reads
    from oracle import oracle as O
    w = synth.make_workload("t", 500, 12, 4, 120, 0.01, ploidy=2, het=0.02, seed=3, n_reads=7)
    par = O.params_uniform(0.01); par.n_warmup = w.k
    li, lt = w.graph.to_probs()
    m = O.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, li, lt, par)
    mine, (lo, hi) = shard_reads(w.reads, rank, world)
    fr, lf, lb = m.run_node_freqs(O.Reads(mine), "sparse")
    f_all, l_all = allreduce_results(fr, lf.sum(), dist)
    np.save(os.path.join(out_dir, f"f{rank}.npy"), f_all); np.save(os.path.join(out_dir, f"l{rank}.npy"), l_all)
    np.save(os.path.join(out_dir, f"b{rank}.npy"), np.array([lo, hi]))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from dbgphmm_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 4000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_allreduce_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from dbgphmm_b200 import synth
    from oracle import oracle as O
    w = synth.make_workload("t", 500, 12, 4, 120, 0.01, ploidy=2, het=0.02, seed=3, n_reads=7)
    par = O.params_uniform(0.01); par.n_warmup = w.k
    li, lt = w.graph.to_probs()
    m = O.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, li, lt, par)
    fr, lf, lb = m.run_node_freqs(O.Reads(w.reads), "sparse")
    f0, f1 = np.load(tmp_path / "f0.npy"), np.load(tmp_path / "f1.npy")
    l0, l1 = np.load(tmp_path / "l0.npy"), np.load(tmp_path / "l1.npy")
    b0, b1 = np.load(tmp_path / "b0.npy"), np.load(tmp_path / "b1.npy")
    assert list(b0) == [0, 4] and list(b1) == [4, 7]
    assert np.array_equal(f0, f1) and np.array_equal(l0, l1)          # every rank ends with the same result
    assert np.allclose(f0, fr, rtol=1e-12, atol=1e-15)
    assert abs(l0[0] - lf.sum()) < 1e-9 * abs(lf.sum())


class _OracleCandidates:
    """stand-in for hmmv2.PHMMModel with several parameter sets: the oracle scores the shard once per candidate"""

    def __init__(self, w, X, par):
        from oracle import oracle as O
        self.O, self.sg, self.X = O, w.graph, X
        li, lt = w.graph.to_probs("normal", X[0])
        self.m = O.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, li, lt, par)

    def n_batch(self):
        return len(self.X)

    def set_copy_nums_batch(self, X, mode):
        assert mode == "normal"
        self.X = np.asarray(X)

    def to_full_prob_reads(self, reads, mappings, use_max_ratio):
        O = self.O
        om = None if mappings is None else O.Mappings(mappings.read_off, mappings.row_off, mappings.nodes, mappings.probs)
        tot = []
        for x in self.X:
            self.m.set_probs(*self.sg.to_probs("normal", x))
            tot.append(self.m.to_full_prob_reads(O.Reads([reads[r] for r in range(len(reads))]), om, use_max_ratio)[0])
        return np.array(tot), None


def _case_c4():
    from dbgphmm_b200 import synth
    from oracle import oracle as O
    w = synth.make_workload("t", 500, 12, 4, 120, 0.01, ploidy=2, het=0.02, seed=3, n_reads=5)
    par = O.params_uniform(0.01); par.n_warmup = w.k
    rng = np.random.default_rng(1)
    X = np.stack([w.graph.node_copy_num, w.graph.node_copy_num + (rng.random(w.graph.n_nodes) < 0.05)])
    return w, par, X


def _worker_c4(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dbgphmm_b200 import hmmv2 as H
    from dbgphmm_b200.dist import full_prob_reads_sharded
    from oracle import oracle as O
    w, par, X = _case_c4()
    model = _OracleCandidates(w, X, par)
    om = model.m.generate_mappings(O.Reads(w.reads), None, False)
    maps = H.Mappings(om.read_off, om.row_off, om.nodes, om.probs)          # (host-only handle: no GPU needed)
    with_maps = full_prob_reads_sharded(model, w.reads, maps, dist)
    # more ranks than reads: the last rank's shard is empty
    few = full_prob_reads_sharded(model, w.reads[:2], None, dist, use_max_ratio=False)
    np.save(os.path.join(out_dir, f"m{rank}.npy"), with_maps); np.save(os.path.join(out_dir, f"e{rank}.npy"), few)
    # the other way round: candidates sharded, every rank scores all reads (4 candidates over 3 ranks: 2 + 1 + 1)
    from dbgphmm_b200.dist import full_prob_candidates_sharded
    X4 = np.stack([X[0], X[1], X[0] + 1, X[1] + 1])
    by_cand = full_prob_candidates_sharded(model, X4, H.Reads(w.reads), maps, dist)
    np.save(os.path.join(out_dir, f"c{rank}.npy"), by_cand)
    # mapping generation: per-read output stays on its rank, only the [N] node frequencies are summed
    from dbgphmm_b200.dist import mappings_to_freqs_sharded
    o2 = O.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, *w.graph.to_probs("uniform"), par)
    f_all, mine, (lo, hi) = mappings_to_freqs_sharded(o2, w.reads, dist, use_max_ratio=False, make_reads=O.Reads)
    assert mine.n_reads() == hi - lo
    np.save(os.path.join(out_dir, f"q{rank}.npy"), f_all)
    dist.destroy_process_group()


def test_three_rank_sharded_full_prob_reads_over_candidates(tmp_path):
    """Batched P(R|X) (BASELINE configs[3]) with the reads sharded over ranks, mappings travelling with their reads, one all-reduce
    of the per-candidate sums; world_size 3 so that 2 reads leave one rank without work."""
    world = 3
    mp.spawn(_worker_c4, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from dbgphmm_b200 import hmmv2 as H
    from oracle import oracle as O
    w, par, X = _case_c4()
    model = _OracleCandidates(w, X, par)
    om = model.m.generate_mappings(O.Reads(w.reads), None, False)
    maps = H.Mappings(om.read_off, om.row_off, om.nodes, om.probs)
    want_m = model.to_full_prob_reads(H.Reads(w.reads), maps, True)[0]
    want_e = model.to_full_prob_reads(H.Reads(w.reads[:2]), None, False)[0]
    for r in range(world):
        assert np.allclose(np.load(tmp_path / f"m{r}.npy"), want_m, rtol=1e-12, atol=0)
        assert np.allclose(np.load(tmp_path / f"e{r}.npy"), want_e, rtol=1e-12, atol=0)
    assert want_m[0] != want_m[1]
    X4 = np.stack([X[0], X[1], X[0] + 1, X[1] + 1])
    model.set_copy_nums_batch(X4, "normal")
    want_c = model.to_full_prob_reads(H.Reads(w.reads), maps, True)[0]
    for r in range(world):
        assert np.allclose(np.load(tmp_path / f"c{r}.npy"), want_c, rtol=1e-12, atol=0)
    assert len(set(np.round(want_c, 6))) == 4
    o2 = O.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, *w.graph.to_probs("uniform"), par)
    want_f = o2.generate_mappings(O.Reads(w.reads), None, False).to_node_freqs(o2.n_nodes)
    for r in range(world):
        assert np.allclose(np.load(tmp_path / f"q{r}.npy"), want_f, rtol=1e-12, atol=1e-15)
    assert abs(want_f.sum() - sum(len(x) for x in w.reads)) < 0.02 * want_f.sum()      # about one node per base (top-n lists drop the tail)


def _worker_inplace(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dbgphmm_b200.dist import allreduce_results, packed_buffer
    # the exchange buffer of bench.py: adjacent views of one buffer -> ONE collective, in place, no copy
    buf, freqs, logp = packed_buffer(5, 1)
    freqs[:] = torch.arange(5, dtype=torch.float64) + rank; logp[0] = -10.0 * (rank + 1)
    calls = []
    real = dist.all_reduce
    dist.all_reduce = lambda t, *a, **k: (calls.append(t.numel()), real(t, *a, **k))[1]
    f, l = allreduce_results(freqs, logp, dist)
    assert f.data_ptr() == freqs.data_ptr() and l.data_ptr() == logp.data_ptr() and calls == [6]
    # separate tensors: both are still reduced in place (a second, tiny collective)
    a = torch.full((4,), float(rank + 1), dtype=torch.float64); b = torch.tensor([float(rank)], dtype=torch.float64)
    calls.clear()
    f2, l2 = allreduce_results(a, b, dist)
    assert f2.data_ptr() == a.data_ptr() and calls == [4, 1]
    dist.all_reduce = real
    np.save(os.path.join(out_dir, f"p{rank}.npy"), np.concatenate([buf.numpy(), a.numpy(), b.numpy()]))
    dist.destroy_process_group()


def test_allreduce_results_reduces_torch_tensors_in_place(tmp_path):
    """the torch path used to all-reduce a concatenated copy and leave its inputs untouched."""
    world = 3
    mp.spawn(_worker_inplace, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want = np.concatenate([3 * np.arange(5) + 3, [-60.0], [6.0] * 4, [3.0]])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"p{r}.npy"), want)
