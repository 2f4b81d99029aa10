"""CPU checks of the synthetic workload generator and the host-side graph builder bench.py and the tests rely on
(dbgphmm_b200/synth.py, graphs.build_dbg): determinism per seed, the de Bruijn structure, and that the copy
numbers turn into proper probability distributions (graph/seq_graph.rs:160-223)."""
import numpy as np

from dbgphmm_b200 import graphs, synth


def test_workloads_are_deterministic_per_seed():
    a = synth.make_workload("a", 3000, 16, 3, 200, 0.003, ploidy=2, het=0.01, seed=7, n_reads=10)
    b = synth.make_workload("b", 3000, 16, 3, 200, 0.003, ploidy=2, het=0.01, seed=7, n_reads=10)
    c = synth.make_workload("c", 3000, 16, 3, 200, 0.003, ploidy=2, het=0.01, seed=8, n_reads=10)
    assert all(np.array_equal(x, y) for x, y in zip(a.reads, b.reads)) and np.array_equal(a.graph.src, b.graph.src)
    assert np.array_equal(a.graph.base, b.graph.base) and np.array_equal(a.graph.node_copy_num, b.graph.node_copy_num)
    assert not np.array_equal(a.haplotypes[0], c.haplotypes[0])
    for r in a.reads:
        assert set(bytes(r)) <= set(b"ACGT") and 150 <= len(r) <= 250


def test_build_dbg_is_the_kmer_graph_of_the_haplotypes():
    k = 12
    h0 = synth.random_genome(800, 3)
    h1 = synth.mutate_substitutions(h0, 0.02, 4)
    g, ids = graphs.build_dbg([h0.tobytes(), h1.tobytes()], k, seed=5)
    pad = lambda h: b"n" * (k - 1) + h.tobytes() + b"n" * (k - 1)
    kmers = {}
    for h in (h0, h1):
        s = pad(h)
        for j in range(len(s) - k + 1):
            kmers[s[j:j + k]] = kmers.get(s[j:j + k], 0) + 1
    assert g.n_nodes == len(kmers)
    # node of every k-mer: emission = its last base, copy number = its multiplicity over the haplotypes
    for hid, h in enumerate((h0, h1)):
        s = pad(h)
        for j in (0, 5, k - 1, 400, len(s) - k):
            v = int(ids[hid][j])
            assert g.base[v] == s[j + k - 1] and g.node_copy_num[v] == kmers[s[j:j + k]]
    # edges = (k-1)-overlaps, except through the all-n terminal; consecutive k-mers of a haplotype are connected
    edges = set(zip(g.src.tolist(), g.dst.tolist()))
    for hid in (0, 1):
        for j in (0, 1, 300, 301, len(ids[hid]) - 2):
            assert (int(ids[hid][j]), int(ids[hid][j + 1])) in edges
    assert (int(ids[0][-1]), int(ids[0][0])) not in edges          # nothing passes through nnn...n
    indeg = np.bincount(g.dst, minlength=g.n_nodes); outdeg = np.bincount(g.src, minlength=g.n_nodes)
    assert indeg.max() <= 5 and outdeg.max() <= 5                   # multi_dbg.rs:82


def test_copy_numbers_become_probability_distributions():
    w = synth.make_workload("p", 2000, 14, 2, 150, 0.003, ploidy=2, het=0.02, seed=2, n_reads=4)
    g = w.graph
    for mode in ("normal", "non_zero", "uniform"):
        li, lt = g.to_probs(mode)
        with np.errstate(divide="ignore"):
            assert abs(np.exp(li).sum() - 1.0) < 1e-12                                    # seq_graph.rs:160-179
            out = np.zeros(g.n_nodes); np.add.at(out, g.src, np.exp(lt))                    # seq_graph.rs:180-211
        has_child = np.bincount(g.src, minlength=g.n_nodes) > 0
        emitting_children = np.zeros(g.n_nodes); np.add.at(emitting_children, g.src, (g.base[g.dst] != ord("n")).astype(float))
        ok = has_child & (emitting_children > 0)
        assert np.allclose(out[ok], 1.0, atol=1e-12), mode
        assert np.isneginf(li[g.base == ord("n")]).all()                                  # null-base nodes never start a read
