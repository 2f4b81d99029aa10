"""Host-side graph builders that feed the PHMM path (pure numpy, no GPU, no oracle).

They restate the reference's model constructors on the caller side of the hot path so that the
same (edge list, emissions, copy numbers) reach both the CUDA library and the test oracle:

* GenomeGraph -> per-base SimpleSeqGraph     graph/genome_graph.rs:252-397
* MultiDbg full graph -> node-centric graph   multi_dbg.rs:1551-1604 (add_terminal = false, :1370-1390)
* copy numbers -> init / trans probabilities  graph/seq_graph.rs:110-273
* k-mer de Bruijn graph of a set of haplotypes (synthetic stand-in for draft.rs:322-380)

Edge order matters: petgraph iterates a node's edges most-recently-added first
(graph/iterators.rs:104-155 over petgraph 0.6 adjacency lists), and the PHMM folds parent sums in
that order, so every builder returns edges in the reference's insertion order.
"""
import numpy as np

NULL_BASE = ord("n")  # common.rs:21


class SeqGraph:
    """Node-centric sequence graph: one base per node (seq_graph.rs SeqNode/SeqEdge)."""

    def __init__(self, src, dst, base, node_copy_num, edge_copy_num=None):
        self.src = np.ascontiguousarray(src, np.uint32)
        self.dst = np.ascontiguousarray(dst, np.uint32)
        self.base = np.ascontiguousarray(base, np.uint8)
        self.node_copy_num = np.ascontiguousarray(node_copy_num, np.int64)
        # -1 == None (seq_graph.rs:198-209 branch)
        self.edge_copy_num = None if edge_copy_num is None else np.ascontiguousarray(edge_copy_num, np.int64)

    @property
    def n_nodes(self):
        return len(self.base)

    @property
    def n_edges(self):
        return len(self.src)

    def to_probs(self, mode="normal", node_copy_num=None):
        return copy_nums_to_probs(self.src, self.dst, self.base,
                                  self.node_copy_num if node_copy_num is None else node_copy_num,
                                  self.edge_copy_num, mode)


def _ln(x):
    with np.errstate(divide="ignore"):
        return np.log(np.asarray(x, np.float64))


def copy_nums_to_probs(src, dst, base, node_copy_num, edge_copy_num=None, mode="normal"):
    """SeqGraph::to_phmm / to_non_zero_phmm / to_uniform_phmm (seq_graph.rs:160-273).

    Returns (log_init[N], log_trans[E]) as natural logs, -inf for zero probability.
    init_v  = ln(c_v) - ln(sum_{emittable u} c_u)                       (:166-171, Prob division)
    trans_e = ln(c_child / sum_{emittable children u of parent} c_u)    (:198-209)  edge copy num None
            = ln(c_e / c_parent)                                       (:185-196)  edge copy num Some
    """
    src = np.asarray(src, np.int64); dst = np.asarray(dst, np.int64)
    base = np.asarray(base, np.uint8)
    N, E = len(base), len(src)
    emit = base != NULL_BASE
    if mode == "uniform":
        n_emit = int(emit.sum())
        log_init = np.where(emit, 0.0 - _ln(float(n_emit)), -np.inf)
        n_child = np.bincount(src[emit[dst]], minlength=N).astype(np.float64)
        log_trans = np.where(emit[dst], 0.0 - _ln(n_child[src]), -np.inf)
        return log_init, log_trans
    minc = {"normal": 0, "non_zero": 1}[mode]
    cn = np.maximum(np.asarray(node_copy_num, np.int64), minc)
    total = int(cn[emit].sum())
    log_init = np.where(emit, _ln(cn.astype(np.float64)) - _ln(float(total)), -np.inf)
    child_total = np.bincount(src[emit[dst]], weights=cn[dst][emit[dst]].astype(np.float64), minlength=N)
    tot = child_total[src]
    with np.errstate(divide="ignore", invalid="ignore"):
        lt_none = np.where(emit[dst] & (tot > 0), _ln(cn[dst].astype(np.float64) / np.where(tot > 0, tot, 1.0)), -np.inf)
    if edge_copy_num is None:
        return log_init, lt_none
    ecn = np.asarray(edge_copy_num, np.int64)
    raw_parent = np.asarray(node_copy_num, np.int64)[src].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        lt_some = np.where(emit[dst] & (ecn > 0), _ln(ecn.astype(np.float64) / np.where(raw_parent > 0, raw_parent, 1.0)), -np.inf)
    return log_init, np.where(ecn >= 0, lt_some, lt_none)


def genome_graph_to_seq_graph(nodes, edges=()):
    """GenomeGraph::to_seq_graph (genome_graph.rs:339-397), forward strand only.

    nodes: [(seq bytes, copy_num)], edges: [(source, target, copy_num or None)].
    Each base becomes a node; intra-sequence edges carry Some(copy_num) (:267-285)."""
    src, dst, base, ncn, ecn = [], [], [], [], []
    head_tail = []
    for seq, cn in nodes:
        first = len(base)
        for b in seq:
            base.append(b); ncn.append(cn)
        for j in range(first + 1, len(base)):
            src.append(j - 1); dst.append(j); ecn.append(cn)
        head_tail.append((first, len(base) - 1))
    for s, t, cn in edges:
        src.append(head_tail[s][1]); dst.append(head_tail[t][0]); ecn.append(-1 if cn is None else cn)
    return SeqGraph(src, dst, base, ncn, ecn)


def mock_linear():
    """graph/mocks.rs:8-12 : 10 bp linear genome ATTCGATCGT, copy number 1."""
    return genome_graph_to_seq_graph([(b"ATTCGATCGT", 1)])


def mock_crossing(has_edge_copy_number):
    """graph/mocks.rs:44-62 : a(x2), b(x2) -> c(x2), d(x2) crossing; sequences as asserted in mocks.rs:78-85."""
    nodes = [(b"TGCTCTGGCG", 2), (b"ATTAGGAGCA", 2), (b"GCTGATAGGG", 2), (b"CGAAGATGAG", 2)]
    if has_edge_copy_number:
        edges = [(0, 2, 2), (1, 2, 0), (0, 3, 0), (1, 3, 2)]
    else:
        edges = [(0, 2, None), (1, 2, None), (0, 3, None), (1, 3, None)]
    return genome_graph_to_seq_graph(nodes, edges)


def multidbg_to_seq_graph(full_is_terminal, full_edges):
    """MultiDbg::to_seq_graph -> to_node_centric_graph(add_terminal=false)  (multi_dbg.rs:1370-1390,1551-1604).

    full_is_terminal[v]: bool per (k-1)-mer node; full_edges: [(src, dst, base, copy_num)] per k-mer, in
    EdgeIndex order.  PHMM node id == full edge id.  PHMM edges: for each non-terminal full node in index
    order, parents x children nested, each adjacency list newest-edge-first (petgraph)."""
    nV = len(full_is_terminal)
    out_e = [[] for _ in range(nV)]; in_e = [[] for _ in range(nV)]
    for e in range(len(full_edges) - 1, -1, -1):
        s, t, _, _ = full_edges[e]
        out_e[s].append(e); in_e[t].append(e)
    src, dst = [], []
    for v in range(nV):
        if full_is_terminal[v]:
            continue
        for e1 in in_e[v]:
            for e2 in out_e[v]:
                src.append(e1); dst.append(e2)
    base = [b for (_, _, b, _) in full_edges]
    cn = [c for (_, _, _, c) in full_edges]
    return SeqGraph(src, dst, base, cn, None)


def toy_repeat():
    """multi_dbg/toy.rs:260-305 : TCCCAGCAGCAGCAGGAA, k=4, repeat CAG x4."""
    names = ["nnn", "nnT", "nTC", "TCC", "CCC", "CCA", "CAG", "AGC", "GCA", "AGG", "GGA", "GAA", "AAn", "Ann"]
    ix = {n: i for i, n in enumerate(names)}
    E = [("nnn", "nnT", "T", 1), ("nnT", "nTC", "C", 1), ("nTC", "TCC", "C", 1), ("TCC", "CCC", "C", 1),
         ("CCC", "CCA", "A", 1), ("CCA", "CAG", "G", 1), ("CAG", "AGC", "C", 3), ("AGC", "GCA", "A", 3),
         ("GCA", "CAG", "G", 3), ("CAG", "AGG", "G", 1), ("AGG", "GGA", "A", 1), ("GGA", "GAA", "A", 1),
         ("GAA", "AAn", "n", 1), ("AAn", "Ann", "n", 1), ("Ann", "nnn", "n", 1)]
    term = [n == "nnn" for n in names]
    return multidbg_to_seq_graph(term, [(ix[a], ix[b], ord(c), k) for a, b, c, k in E]), 4


# --------------------------------------------------------------------------- k-mer DBG of haplotypes
def _unique_rows(a):
    """np.unique(axis=0, return_inverse, return_counts) on a 2-D uint8 array via a void view (memcmp order)."""
    a = np.ascontiguousarray(a)
    v = a.view(np.dtype((np.void, a.shape[1]))).reshape(-1)
    u, inv, cnt = np.unique(v, return_inverse=True, return_counts=True)
    return u.view(np.uint8).reshape(-1, a.shape[1]), inv.reshape(-1), cnt


def build_dbg(haplotypes, k, seed=0, shuffle=True):
    """k-mer de Bruijn graph of linear haplotypes with n-padded ends, as a node-centric SeqGraph.

    Synthetic stand-in for the draft DBG (multi_dbg/draft.rs:322-380: k-mers, terminal `n` padding,
    copy numbers); copy number of a k-mer = its true multiplicity over the haplotypes.  PHMM node =
    k-mer (full-graph edge), emission = last base, PHMM edge = (k-1)-overlap except through the
    all-`n` terminal (k-1)-mer (multi_dbg.rs:1370-1390 passes add_terminal = false).  Node ids are
    shuffled (seeded) because the reference's ids come out of a hash map (hashdbg.rs:289-313) and
    carry no locality.  Returns (SeqGraph, ids) where ids[h][j] is the node of the k-mer ENDING at
    haplotype position j - (k-1) ... i.e. the k-mer starting at padded position j."""
    wins = []
    for h in haplotypes:
        s = np.concatenate([np.full(k - 1, NULL_BASE, np.uint8), np.frombuffer(bytes(h), np.uint8),
                            np.full(k - 1, NULL_BASE, np.uint8)])
        wins.append(np.lib.stride_tricks.sliding_window_view(s, k))
    allw = np.concatenate(wins)
    uniq, inv, counts = _unique_rows(allw)
    N = len(uniq)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(N) if shuffle else np.arange(N)
    node_cn = np.zeros(N, np.int64); node_cn[perm] = counts
    base = np.zeros(N, np.uint8); base[perm] = uniq[:, k - 1]
    # (k-1)-mer nodes of the full graph: prefix / suffix of every k-mer
    ku, kinv, _ = _unique_rows(np.concatenate([uniq[:, :k - 1], uniq[:, 1:]]))
    pre_id = kinv[:N]; suf_id = kinv[N:]
    is_term = (ku == NULL_BASE).all(axis=1)
    nK = len(ku)
    order_in = np.argsort(suf_id, kind="stable"); order_out = np.argsort(pre_id, kind="stable")
    in_cnt = np.bincount(suf_id, minlength=nK); out_cnt = np.bincount(pre_id, minlength=nK)
    in_off = np.concatenate([[0], np.cumsum(in_cnt)]); out_off = np.concatenate([[0], np.cumsum(out_cnt)])
    simple = (in_cnt == 1) & (out_cnt == 1) & ~is_term
    sv = np.nonzero(simple)[0]
    e_src = [order_in[in_off[sv]]]; e_dst = [order_out[out_off[sv]]]; via = [sv]
    for v in np.nonzero(~simple & ~is_term & (in_cnt > 0) & (out_cnt > 0))[0]:
        a = order_in[in_off[v]:in_off[v + 1]]; b = order_out[out_off[v]:out_off[v + 1]]
        e_src.append(np.repeat(a, len(b))); e_dst.append(np.tile(b, len(a))); via.append(np.full(len(a) * len(b), v))
    e_src = np.concatenate(e_src); e_dst = np.concatenate(e_dst); via = np.concatenate(via)
    eo = np.argsort(via, kind="stable")  # edges grouped by the (k-1)-mer they pass through
    g = SeqGraph(perm[e_src[eo]], perm[e_dst[eo]], base, node_cn, None)
    ids = perm[inv].astype(np.uint32)
    out_ids, p = [], 0
    for w in wins:
        out_ids.append(ids[p:p + len(w)]); p += len(w)
    return g, out_ids
