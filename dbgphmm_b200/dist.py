"""Read sharding over the GPUs of one box: every rank holds the full graph and a contiguous shard of the
reads; per-node expected frequencies and the summed ln P(R|X) are combined with ONE all-reduce (NCCL over NVLink on
GPUs, gloo in the CPU tests).  There is no communication inside the DP."""
import numpy as np


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced split: the first (n_items % world) ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_reads(reads, rank, world):
    lo, hi = shard_bounds(len(reads), rank, world)
    return reads[lo:hi], (lo, hi)


def allreduce_results(node_freqs, logp_sum, dist=None, device=None):
    """Sum node_freqs [N] and the scalar(s) logp_sum over all ranks in one collective.  Inputs may be numpy arrays or torch
    tensors (on `device` for NCCL).  Returns (node_freqs, logp_sum) of the same kind as given."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return node_freqs, logp_sum
    import torch
    is_np = isinstance(node_freqs, np.ndarray)
    f = torch.from_numpy(np.ascontiguousarray(node_freqs)) if is_np else node_freqs
    l = torch.as_tensor(np.atleast_1d(np.asarray(logp_sum, np.float64))) if not torch.is_tensor(logp_sum) else logp_sum
    if device is not None:
        f = f.to(device); l = l.to(device)
    buf = torch.cat([f.reshape(-1).to(torch.float64), l.reshape(-1).to(torch.float64)])  # one payload, one collective
    dist.all_reduce(buf)
    f_out, l_out = buf[:f.numel()], buf[f.numel():]
    if is_np:
        return f_out.cpu().numpy(), l_out.cpu().numpy()
    return f_out, l_out
