"""Deterministic synthetic workloads for the PHMM path (BASELINE.json configs C1..C5).

Own seeded generator (numpy PCG64): the reference's generators use rand 0.8.5 (random_seq.rs:65-93,
e2e.rs:163-240) and are not bit-reproduced — parity is judged on the DP given identical graph + reads.

* genome: i.i.d. uniform ACGT; second haplotype = substitutions at the stated heterozygosity
* reads : forward-strand fragments of fixed length from a uniform start (e2e.rs:174-183), then
          substitution / insertion / deletion each with probability p per position (HiFi: p = 0.001)
* graph : k-mer DBG of the haplotypes, copy numbers = true multiplicities (graphs.build_dbg)
"""
import numpy as np
from . import graphs

ACGT = np.frombuffer(b"ACGT", np.uint8)


def random_genome(length, seed):
    rng = np.random.default_rng(seed)
    return ACGT[rng.integers(0, 4, length)]


def mutate_substitutions(seq, rate, seed):
    """Heterozygous haplotype: each position substituted with probability `rate` by a different base."""
    rng = np.random.default_rng(seed)
    out = seq.copy()
    pos = np.nonzero(rng.random(len(seq)) < rate)[0]
    code = np.searchsorted(ACGT, out[pos])
    out[pos] = ACGT[(code + rng.integers(1, 4, len(pos))) % 4]
    return out


def tandem_repeat_genome(unit_len, n_units, flank_len, seed, divergence=0.0):
    """KIR-like layout (genome.rs:294-341 style): unique flank + unit x n (each copy diverged) + unique flank."""
    unit = random_genome(unit_len, seed)
    parts = [random_genome(flank_len, seed + 1)]
    for u in range(n_units):
        parts.append(mutate_substitutions(unit, divergence, seed + 10 + u) if divergence > 0 else unit)
    parts.append(random_genome(flank_len, seed + 2))
    return np.concatenate(parts)


def sample_reads(haplotypes, coverage, read_len, p_err, seed):
    """Fragment reads: total bases = genome_size x coverage (e2e.rs:177); errors sub/ins/del each w.p. p_err."""
    rng = np.random.default_rng(seed)
    genome_size = sum(len(h) for h in haplotypes)
    n_reads = max(1, int(round(genome_size * coverage / read_len)))
    reads = []
    for _ in range(n_reads):
        h = haplotypes[rng.integers(0, len(haplotypes))]
        L = min(read_len, len(h))
        s = int(rng.integers(0, len(h) - L + 1))
        frag = h[s:s + L]
        if p_err > 0:
            u = rng.random(L)
            sub = u < p_err
            ins = (u >= p_err) & (u < 2 * p_err)
            dele = (u >= 2 * p_err) & (u < 3 * p_err)
            frag = frag.copy()
            if sub.any():
                code = np.searchsorted(ACGT, frag[sub])
                frag[sub] = ACGT[(code + rng.integers(1, 4, int(sub.sum()))) % 4]
            keep = ~dele
            if ins.any():
                rep = np.ones(L, np.int64); rep[ins] = 2
                idx = np.repeat(np.arange(L), rep * keep)
                out = frag[idx]
                # the second copy of an `ins` position becomes a random base
                first = np.concatenate([[True], idx[1:] != idx[:-1]])
                n_new = int((~first).sum())
                out[~first] = ACGT[rng.integers(0, 4, n_new)]
                frag = out
            else:
                frag = frag[keep]
        if len(frag) == 0:
            frag = h[s:s + 1]
        reads.append(np.ascontiguousarray(frag))
    return reads


class Workload:
    def __init__(self, name, k, haplotypes, reads, graph, kmer_ids, p_err):
        self.name, self.k, self.haplotypes, self.reads, self.graph, self.kmer_ids, self.p_err = \
            name, k, haplotypes, reads, graph, kmer_ids, p_err


def make_workload(name, genome_len, k, coverage, read_len, p_err, ploidy=1, het=0.0, seed=0, n_reads=None):
    """C1/C2: (10_000, 40, 20, 1_000, 0.001); C3: (1_000_000, 40, 20, 10_000, 0.001, ploidy 2, het 0.01)."""
    h0 = random_genome(genome_len, seed)
    haps = [h0]
    for j in range(1, ploidy):
        haps.append(mutate_substitutions(h0, het, seed + j))
    g, ids = graphs.build_dbg([h.tobytes() for h in haps], k, seed=seed + 100)
    reads = sample_reads(haps, coverage, read_len, p_err, seed + 1000)
    if n_reads is not None:
        reads = reads[:n_reads]
    return Workload(name, k, haps, reads, g, ids, p_err)
