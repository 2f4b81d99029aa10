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
    """Sum node_freqs [N] and the scalar(s) logp_sum over all ranks.  numpy inputs: packed into one payload, ONE collective, new arrays
    are returned.  torch tensors: reduced IN PLACE (the inputs hold the sums afterwards and are returned); when both are views of one
    buffer -- `packed_buffer` below -- pass the buffer's two views and a single collective over the whole buffer is issued, otherwise
    the (tiny) logp tensor takes a second collective.  No copy of the [N] vector is made on the torch path."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return node_freqs, logp_sum
    import torch
    if isinstance(node_freqs, np.ndarray) or not torch.is_tensor(node_freqs):
        f = torch.from_numpy(np.ascontiguousarray(node_freqs, np.float64))
        l = torch.as_tensor(np.atleast_1d(np.asarray(logp_sum, np.float64)))
        buf = torch.cat([f.reshape(-1), l.reshape(-1)])      # one payload, one collective
        if device is not None:
            buf = buf.to(device)
        dist.all_reduce(buf)
        buf = buf.cpu()
        return buf[:f.numel()].numpy(), buf[f.numel():].numpy()
    l = logp_sum if torch.is_tensor(logp_sum) else torch.as_tensor(np.atleast_1d(np.asarray(logp_sum, np.float64)), device=node_freqs.device)
    nf, nl = node_freqs.numel(), l.numel()
    same = (node_freqs.is_contiguous() and l.is_contiguous() and node_freqs.dtype == l.dtype and node_freqs.device == l.device
            and node_freqs.untyped_storage().data_ptr() == l.untyped_storage().data_ptr()
            and l.storage_offset() == node_freqs.storage_offset() + nf)
    if same:   # adjacent views of one buffer: reduce the buffer itself
        whole = torch.as_strided(node_freqs, (nf + nl,), (1,), node_freqs.storage_offset())
        dist.all_reduce(whole)
    else:
        dist.all_reduce(node_freqs)
        dist.all_reduce(l)
    return node_freqs, l


def packed_buffer(n_nodes, n_scalars=1, device=None):
    """One f64 buffer [n_nodes + n_scalars] and its two views (node_freqs, scalars): what the ranks exchange, laid out so that
    allreduce_results issues a single in-place collective."""
    import torch
    buf = torch.zeros(n_nodes + n_scalars, dtype=torch.float64, device=device)
    return buf, buf[:n_nodes], buf[n_nodes:]


def full_prob_reads_sharded(model, seqs, mappings, dist=None, use_max_ratio=True, device=None, make_reads=None):
    """ln P(R|X) for every candidate X of `model` (PHMMModel.to_full_prob_reads, freq.rs:175-192) with the reads sharded over the
    ranks: each rank scores its contiguous shard (with the shard's own mappings) against all candidates, then ONE all-reduce of
    the [n_batch] per-candidate sums.  `seqs`: the full list of reads, the same on every rank; returns the [n_batch] totals,
    identical on every rank.  `make_reads`: constructor of the read collection (default hmmv2.Reads)."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    lo, hi = shard_bounds(len(seqs), rank, world)
    if make_reads is None:
        from .hmmv2 import Reads as make_reads
    if hi > lo:
        tot = np.asarray(model.to_full_prob_reads(make_reads(seqs[lo:hi]), mappings.slice(lo, hi) if mappings is not None else None, use_max_ratio)[0], np.float64)
    else:
        tot = np.zeros(model.n_batch())      # more ranks than reads: an empty shard contributes ln 1 to every candidate
    _, out = allreduce_results(np.zeros(0), tot, dist, device)
    return np.asarray(out)


def full_prob_candidates_sharded(model, full_copy_nums, reads, mappings, dist=None, use_max_ratio=True, mode="normal", device=None):
    """The fallback of SURVEY.md 8e for fewer reads than GPUs: shard the CANDIDATES instead.  Every rank derives (init, trans) for
    its contiguous slice of `full_copy_nums` [B][n_nodes] (PHMMModel.set_copy_nums_batch), scores ALL reads against it and writes
    its slice of a zeroed [B] vector; ONE all-reduce (a sum with disjoint supports) assembles ln P(R|X) for every candidate on every
    rank.  `reads` / `mappings`: the whole read set, the same on every rank."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    x = np.ascontiguousarray(np.atleast_2d(full_copy_nums), np.uint32)
    lo, hi = shard_bounds(len(x), rank, world)
    out = np.zeros(len(x))
    if hi > lo:
        model.set_copy_nums_batch(x[lo:hi], mode)
        out[lo:hi] = np.asarray(model.to_full_prob_reads(reads, mappings, use_max_ratio)[0], np.float64)
    _, tot = allreduce_results(np.zeros(0), out, dist, device)
    return np.asarray(tot)


def mappings_to_freqs_sharded(model, seqs, dist=None, use_max_ratio=True, device=None, make_reads=None):
    """MultiDbg::generate_mappings + mappings_to_freqs (posterior.rs:609-630, draft.rs:201-212) with the reads sharded over the ranks:
    every rank maps its own shard (the mappings stay with their reads: no collective), and the per-node frequencies
    `Mappings::to_node_freqs` (hint.rs:161-171) are summed with ONE all-reduce.  Returns (node_freqs [N] identical on every rank,
    this rank's Mappings, (lo, hi) = the reads they belong to)."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    lo, hi = shard_bounds(len(seqs), rank, world)
    if make_reads is None:
        from .hmmv2 import Reads as make_reads
    mine = model.generate_mappings(make_reads(seqs[lo:hi]), None, use_max_ratio) if hi > lo else None
    f = np.asarray(mine.to_node_freqs(model.n_nodes), np.float64) if mine is not None else np.zeros(model.n_nodes)
    f_all, _ = allreduce_results(f, np.zeros(1), dist, device)
    return np.asarray(f_all), mine, (lo, hi)
