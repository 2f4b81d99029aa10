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
