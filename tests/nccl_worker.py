"""Worker of tests/test_nccl_gpu.py (one process per GPU under torchrun): reads sharded over the ranks, node frequencies and the summed
ln P(R) combined with dbgphmm_b200.dist.allreduce_results over NCCL, compared on every rank with the unsharded run on its own GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from dbgphmm_b200 import dist as D, hmmv2 as H, synth
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = synth.make_workload("nccl", 40_000, 40, 4, 1_200, 0.001, ploidy=2, het=0.01, seed=3)
    reads = list(w.reads)[:101]                                  # an odd count: the shards differ in size
    li, lt = w.graph.to_probs("normal")
    par = H.params_uniform(0.001); par.n_warmup = w.k
    m = H.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, li, lt, par, device=local)
    N = w.graph.n_nodes
    for strat in ("stream", "store"):
        os.environ["DBGPHMM_STRATEGY"] = strat
        mine, (lo, hi) = D.shard_reads(reads, rank, world)
        buf, freqs, logp = D.packed_buffer(N, 1, "cuda")
        lp = torch.zeros(max(hi - lo, 1), dtype=torch.float64, device="cuda")
        for step in range(3):
            buf.zero_(); torch.cuda.current_stream().synchronize()
            cells = m.run_node_freqs_dev(H.Reads(mine), "sparse", freqs.data_ptr(), logp_fwd_ptr=lp.data_ptr())
            logp[0] = lp[:hi - lo].sum()
            f_all, l_all = D.allreduce_results(freqs, logp, dist)
            assert f_all.data_ptr() == freqs.data_ptr()          # in place
            torch.cuda.synchronize()
            fr, lf, lb, cells1 = m.run_node_freqs(H.Reads(reads), "sparse")   # unsharded, this GPU
            got = freqs.cpu().numpy()
            assert np.allclose(got, fr, rtol=1e-11, atol=1e-13), (strat, step, np.abs(got - fr).max())
            assert abs(float(logp.item()) - float(lf.sum())) <= 1e-9 * abs(float(lf.sum())), (strat, step)
            tot = torch.tensor([float(sum(cells))], dtype=torch.float64, device="cuda"); dist.all_reduce(tot)
            assert int(tot.item()) == sum(cells1), (strat, step)
    # numpy path of the same function (one packed payload, new arrays)
    f_np, l_np = D.allreduce_results(np.full(5, rank + 1.0), np.array([1.0]), dist, device="cuda")
    assert np.allclose(f_np, world * (world + 1) / 2) and np.allclose(l_np, world)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok", flush=True)


if __name__ == "__main__":
    main()
