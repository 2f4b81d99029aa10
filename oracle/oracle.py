"""ctypes front-end of the CPU oracle (oracle/dbgphmm_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Never imported by dbgphmm_b200/.

Method names mirror the reference's `impl PHMMModel` surface (hmmv2/forward.rs, backward.rs,
freq.rs, hint.rs) so parity tests read like the reference's own tests.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "dbgphmm_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class Params(C.Structure):
    """hmmv2/params.rs:16-66 ; p_* are natural-log probabilities (Prob.0)."""
    _fields_ = [(n, C.c_double) for n in (
        "p_mismatch", "p_match", "p_random", "p_gap_open", "p_gap_ext", "p_end",
        "p_MM", "p_IM", "p_DM", "p_MI", "p_II", "p_DI", "p_MD", "p_ID", "p_DD")] + [
        ("n_active_nodes", C.c_uint32), ("n_warmup", C.c_uint32),
        ("warmup_threshold", C.c_uint32), ("n_max_gaps", C.c_uint32),
        ("active_node_max_ratio", C.c_double)]

    def copy(self):
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u64, u32, i64, dbl, cint = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_double, C.c_int
        L.orc_last_error.restype = C.c_char_p
        L.orc_params_uniform.argtypes = [dbl, C.POINTER(Params)]
        L.orc_params_new.argtypes = [dbl, dbl, dbl, dbl, u32, u32, C.POINTER(Params)]
        L.orc_padd.argtypes = [dbl, dbl]; L.orc_padd.restype = dbl
        L.orc_seqgraph_to_phmm.argtypes = [u32, u32, vp, vp, vp, vp, vp, cint, vp, vp]
        L.orc_model_create.argtypes = [u32, u32, vp, vp, vp, vp, vp, C.POINTER(Params)]; L.orc_model_create.restype = vp
        L.orc_model_set_probs.argtypes = [vp, vp, vp]
        L.orc_model_set_params.argtypes = [vp, C.POINTER(Params)]
        L.orc_model_destroy.argtypes = [vp]
        L.orc_forward.argtypes = [vp, vp, u64, cint, vp, vp]; L.orc_forward.restype = vp
        L.orc_backward.argtypes = [vp, vp, u64, cint, vp, vp, vp]; L.orc_backward.restype = vp
        L.orc_tables_destroy.argtypes = [vp]
        L.orc_tables_len.argtypes = [vp]; L.orc_tables_len.restype = u64
        L.orc_tables_full_prob.argtypes = [vp]; L.orc_tables_full_prob.restype = dbl
        L.orc_tables_row_info.argtypes = [vp, i64, vp, vp]
        L.orc_tables_row_export.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.orc_tables_row_top_nodes.argtypes = [vp, i64, cint, u64, dbl, vp]; L.orc_tables_row_top_nodes.restype = u64
        L.orc_output_node_freqs.argtypes = [vp, vp, vp]; L.orc_output_node_freqs.restype = cint
        L.orc_output_mapping.argtypes = [vp, vp, cint, u64, dbl, vp, vp, vp]; L.orc_output_mapping.restype = cint
        L.orc_output_edge_init_freqs.argtypes = [vp, vp, vp, vp, u64, vp, vp]; L.orc_output_edge_init_freqs.restype = cint
        L.orc_full_prob_reads.argtypes = [vp, u64, vp, vp, vp, vp, vp, cint, vp, cint]; L.orc_full_prob_reads.restype = dbl
        L.orc_run_node_freqs.argtypes = [vp, u64, vp, vp, cint, cint, vp, vp, vp, vp, vp, vp, cint]; L.orc_run_node_freqs.restype = cint
        L.orc_count_cells.argtypes = [vp, vp, u64, cint, cint, cint]; L.orc_count_cells.restype = u64
        L.orc_generate_mappings.argtypes = [vp, u64, vp, vp, vp, vp, vp, cint, cint]; L.orc_generate_mappings.restype = vp
        L.orc_mappings_destroy.argtypes = [vp]
        L.orc_mappings_n_entries.argtypes = [vp]; L.orc_mappings_n_entries.restype = u64
        L.orc_mappings_n_rows.argtypes = [vp]; L.orc_mappings_n_rows.restype = u64
        L.orc_mappings_export.argtypes = [vp, vp, vp, vp, vp]
        L.orc_max_threads.restype = cint
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _err():
    return lib().orc_last_error().decode()


def params_uniform(p):
    """PHMMParams::uniform (params.rs:116-124)."""
    q = Params()
    lib().orc_params_uniform(float(p), C.byref(q))
    return q


def params_new(p_mismatch, p_gap_open, p_gap_ext, p_end, n_active_nodes, n_warmup):
    q = Params()
    lib().orc_params_new(p_mismatch, p_gap_open, p_gap_ext, p_end, n_active_nodes, n_warmup, C.byref(q))
    return q


def padd(a, b):
    return lib().orc_padd(a, b)


def max_threads():
    return lib().orc_max_threads()


def seqgraph_to_phmm(src, dst, base, node_copy_num, edge_copy_num=None, mode=0):
    """SeqGraph::to_phmm / to_non_zero_phmm / to_uniform_phmm (graph/seq_graph.rs:160-273).

    mode 0 normal, 1 non_zero, 2 uniform.  Returns (log_init[N], log_trans[E])."""
    src = np.ascontiguousarray(src, np.uint32); dst = np.ascontiguousarray(dst, np.uint32)
    base = np.ascontiguousarray(base, np.uint8)
    cn = np.ascontiguousarray(node_copy_num, np.int64)
    ecn = None if edge_copy_num is None else np.ascontiguousarray(edge_copy_num, np.int64)
    li = np.empty(len(base), np.float64); lt = np.empty(len(src), np.float64)
    lib().orc_seqgraph_to_phmm(len(base), len(src), _p(src), _p(dst), _p(base), _p(cn), _p(ecn), mode, _p(li), _p(lt))
    return li, lt


class Row:
    """One PHMMTable (table.rs:42-73) in log space."""
    __slots__ = ("is_dense", "ids", "m", "i", "ids_d", "d", "mb", "ib", "e")

    def merged(self, n_nodes):
        """to_nodevec as a dense log array (absent = -inf)."""
        v = np.full(n_nodes, -np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            if self.is_dense:
                return np.logaddexp(np.logaddexp(self.m, self.i), self.d)
            v[self.ids] = np.logaddexp(self.m, self.i)
            v[self.ids_d] = np.logaddexp(v[self.ids_d], self.d)
        return v


class Tables:
    """PHMMTables (table.rs:365-435)."""

    def __init__(self, handle, n_nodes):
        self._h = handle
        self.n_nodes = n_nodes

    def __del__(self):
        if self._h:
            lib().orc_tables_destroy(self._h)
            self._h = None

    def __len__(self):
        return int(lib().orc_tables_len(self._h))

    def full_prob(self):
        return lib().orc_tables_full_prob(self._h)

    def row(self, i):
        """i = -1 is init_table."""
        info = np.zeros(3, np.uint64); sc = np.zeros(3, np.float64)
        lib().orc_tables_row_info(self._h, i, _p(info), _p(sc))
        r = Row()
        r.is_dense = bool(info[0]); r.mb, r.ib, r.e = sc
        if r.is_dense:
            N = self.n_nodes
            r.ids = r.ids_d = None
            r.m = np.empty(N); r.i = np.empty(N); r.d = np.empty(N)
            lib().orc_tables_row_export(self._h, i, None, _p(r.m), _p(r.i), None, _p(r.d))
        else:
            nm, nd = int(info[1]), int(info[2])
            r.ids = np.empty(nm, np.uint32); r.m = np.empty(nm); r.i = np.empty(nm)
            r.ids_d = np.empty(nd, np.uint32); r.d = np.empty(nd)
            lib().orc_tables_row_export(self._h, i, _p(r.ids), _p(r.m), _p(r.i), _p(r.ids_d), _p(r.d))
        return r

    def top_nodes(self, i, k):
        out = np.empty(400, np.uint32)
        n = lib().orc_tables_row_top_nodes(self._h, i, 0, k, 0.0, _p(out))
        return out[:n].copy()

    def top_nodes_by_score_ratio(self, i, ratio):
        out = np.empty(400, np.uint32)
        n = lib().orc_tables_row_top_nodes(self._h, i, 1, 0, ratio, _p(out))
        return out[:n].copy()


class Mapping:
    """hint.rs:27-30 : per-base candidate nodes with log probs, for ONE read."""

    def __init__(self, nodes, probs):
        self.nodes = nodes  # list of np.uint32 arrays
        self.probs = probs  # list of np.float64 arrays (log)

    def __len__(self):
        return len(self.nodes)

    def csr(self):
        off = np.zeros(len(self.nodes) + 1, np.uint64)
        off[1:] = np.cumsum([len(x) for x in self.nodes])
        nodes = np.concatenate(self.nodes).astype(np.uint32) if len(self.nodes) else np.zeros(0, np.uint32)
        probs = np.concatenate(self.probs).astype(np.float64) if len(self.probs) else np.zeros(0, np.float64)
        return off, np.ascontiguousarray(nodes), np.ascontiguousarray(probs)


class Mappings:
    """hint.rs:150-152 : CSR over reads -> bases -> (node, logp)."""

    def __init__(self, read_off, row_off, nodes, probs):
        self.read_off, self.row_off, self.nodes, self.probs = read_off, row_off, nodes, probs

    @staticmethod
    def from_list(maps):
        read_off = [0]; row_off = [0]; nodes = []; probs = []
        for m in maps:
            for ns, ps in zip(m.nodes, m.probs):
                nodes.append(np.asarray(ns, np.uint32)); probs.append(np.asarray(ps, np.float64))
                row_off.append(row_off[-1] + len(ns))
            read_off.append(read_off[-1] + len(m.nodes))
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs).astype(dt)) if xs else np.zeros(0, dt)
        return Mappings(np.array(read_off, np.uint64), np.array(row_off, np.uint64), cat(nodes, np.uint32), cat(probs, np.float64))

    def n_reads(self):
        return len(self.read_off) - 1

    def __getitem__(self, r):
        a, b = int(self.read_off[r]), int(self.read_off[r + 1])
        ns = [self.nodes[int(self.row_off[i]):int(self.row_off[i + 1])] for i in range(a, b)]
        ps = [self.probs[int(self.row_off[i]):int(self.row_off[i + 1])] for i in range(a, b)]
        return Mapping(ns, ps)

    def to_node_freqs(self, n_nodes):
        """Mappings::to_node_freqs (hint.rs:161-171) == MultiDbg::mappings_to_freqs (draft.rs:201-212)."""
        f = np.zeros(n_nodes)
        np.add.at(f, self.nodes, np.exp(self.probs))
        return f


class Reads:
    def __init__(self, seqs):
        seqs = [np.frombuffer(bytes(s), np.uint8) if not isinstance(s, np.ndarray) else s.astype(np.uint8) for s in seqs]
        self.offsets = np.zeros(len(seqs) + 1, np.uint64)
        self.offsets[1:] = np.cumsum([len(s) for s in seqs])
        self.bases = np.ascontiguousarray(np.concatenate(seqs)) if seqs else np.zeros(0, np.uint8)

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, r):
        return self.bases[int(self.offsets[r]):int(self.offsets[r + 1])]


def _bases(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), np.uint8).copy()
    return np.ascontiguousarray(x, np.uint8)


class PHMMModel:
    """PHMMModel<PNode,PEdge> (hmmv2/common.rs:61-67) in the oracle."""

    def __init__(self, edge_src, edge_dst, emission, log_init, log_trans, param):
        self.src = np.ascontiguousarray(edge_src, np.uint32)
        self.dst = np.ascontiguousarray(edge_dst, np.uint32)
        self.emission = _bases(emission)
        self.log_init = np.ascontiguousarray(log_init, np.float64)
        self.log_trans = np.ascontiguousarray(log_trans, np.float64)
        self.param = param.copy()
        self.n_nodes = len(self.emission)
        self._h = lib().orc_model_create(self.n_nodes, len(self.src), _p(self.src), _p(self.dst), _p(self.emission),
                                         _p(self.log_init), _p(self.log_trans), C.byref(self.param))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_model_destroy(self._h)
            self._h = None

    def set_probs(self, log_init, log_trans):
        self.log_init = np.ascontiguousarray(log_init, np.float64)
        self.log_trans = np.ascontiguousarray(log_trans, np.float64)
        lib().orc_model_set_probs(self._h, _p(self.log_init), _p(self.log_trans))

    def set_params(self, param):
        self.param = param.copy()
        lib().orc_model_set_params(self._h, C.byref(self.param))

    # ---- forward.rs
    def _fwd(self, x, kind, mapping=None):
        x = _bases(x)
        ro = nd = None
        if mapping is not None:
            ro, nd, _ = mapping.csr()
        h = lib().orc_forward(self._h, _p(x), len(x), kind, _p(ro), _p(nd))
        if not h:
            raise RuntimeError(_err())
        return Tables(h, self.n_nodes)

    def forward(self, x):
        return self._fwd(x, 0)

    def forward_sparse(self, x, use_max_ratio):
        return self._fwd(x, 2 if use_max_ratio else 1)

    def forward_with_mapping(self, x, mapping):
        return self._fwd(x, 3, mapping)

    # ---- backward.rs
    def _bwd(self, x, kind, mapping=None, fwd=None):
        x = _bases(x)
        ro = nd = None
        if mapping is not None:
            ro, nd, _ = mapping.csr()
        h = lib().orc_backward(self._h, _p(x), len(x), kind, _p(ro), _p(nd), fwd._h if fwd is not None else None)
        if not h:
            raise RuntimeError(_err())
        return Tables(h, self.n_nodes)

    def backward(self, x):
        return self._bwd(x, 0)

    def backward_sparse(self, x):
        return self._bwd(x, 1)

    def backward_with_mapping(self, x, mapping):
        return self._bwd(x, 2, mapping)

    def backward_by_forward(self, x, fwd):
        return self._bwd(x, 3, fwd=fwd)

    # ---- freq.rs:42-76
    def run(self, x):
        return PHMMOutput(self.forward(x), self.backward(x))

    def run_sparse(self, x):
        return PHMMOutput(self.forward_sparse(x, False), self.backward_sparse(x))

    def run_sparse_adaptive(self, x, use_max_ratio):
        f = self.forward_sparse(x, use_max_ratio)
        return PHMMOutput(f, self.backward_by_forward(x, f))

    def run_with_mapping(self, x, mapping):
        return PHMMOutput(self.forward_with_mapping(x, mapping), self.backward_with_mapping(x, mapping))

    def q_score_exact(self, edge_freqs, init_freqs):
        """q.rs:66-96 -> (init, trans, prior) over emittable nodes (emission != 'n')."""
        emit = self.emission != ord("n")
        if not (np.isfinite(self.log_init[emit]).all()):
            raise RuntimeError("init_prob is not finite (q.rs:79 asserts)")
        keep = emit[self.src] & emit[self.dst]
        if not np.isfinite(self.log_trans[keep]).all():
            raise RuntimeError("trans_prob is not finite (q.rs:88 asserts)")
        return (float((np.asarray(init_freqs)[emit] * self.log_init[emit]).sum()), float((np.asarray(edge_freqs)[keep] * self.log_trans[keep]).sum()), 0.0)

    # ---- bulk (freq.rs:175-192, hint.rs:193-220)
    def to_full_prob_reads(self, reads, mappings=None, use_max_ratio=True, n_threads=1):
        per = np.empty(len(reads))
        mro = mrw = mnd = None
        if mappings is not None:
            mro, mrw, mnd = mappings.read_off, mappings.row_off, mappings.nodes
        s = lib().orc_full_prob_reads(self._h, len(reads), _p(reads.offsets), _p(reads.bases), _p(mro), _p(mrw), _p(mnd),
                                      int(use_max_ratio), _p(per), n_threads)
        return s, per

    def run_node_freqs(self, reads, mode, use_max_ratio=True, mappings=None, n_threads=1, want_freqs=True):
        """mode: 'dense' | 'sparse' | 'sparse_adaptive' | 'with_mapping'. Returns (freqs[N], logp_fwd[R], logp_bwd[R])."""
        mi = {"dense": 0, "sparse": 1, "sparse_adaptive": 2, "with_mapping": 3}[mode]
        fr = np.zeros(self.n_nodes) if want_freqs else None
        lf = np.empty(len(reads)); lb = np.empty(len(reads))
        mro = mrw = mnd = None
        if mappings is not None:
            mro, mrw, mnd = mappings.read_off, mappings.row_off, mappings.nodes
        rc = lib().orc_run_node_freqs(self._h, len(reads), _p(reads.offsets), _p(reads.bases), mi, int(use_max_ratio),
                                      _p(mro), _p(mrw), _p(mnd), _p(fr), _p(lf), _p(lb), n_threads)
        if rc:
            raise RuntimeError(_err())
        return fr, lf, lb

    def count_cells(self, x, mode, use_max_ratio=True, direction=3):
        x = _bases(x)
        mi = {"dense": 0, "sparse": 1, "sparse_adaptive": 2}[mode]
        return int(lib().orc_count_cells(self._h, _p(x), len(x), mi, int(use_max_ratio), direction))

    def generate_mappings(self, reads, mappings=None, use_max_ratio=True, n_threads=1):
        mro = mrw = mnd = None
        if mappings is not None:
            mro, mrw, mnd = mappings.read_off, mappings.row_off, mappings.nodes
        h = lib().orc_generate_mappings(self._h, len(reads), _p(reads.offsets), _p(reads.bases), _p(mro), _p(mrw), _p(mnd),
                                        int(use_max_ratio), n_threads)
        if not h:
            raise RuntimeError(_err())
        try:
            nr, ne = int(lib().orc_mappings_n_rows(h)), int(lib().orc_mappings_n_entries(h))
            ro = np.empty(len(reads) + 1, np.uint64); rw = np.empty(nr + 1, np.uint64)
            nd = np.empty(ne, np.uint32); pr = np.empty(ne, np.float64)
            lib().orc_mappings_export(h, _p(ro), _p(rw), _p(nd), _p(pr))
        finally:
            lib().orc_mappings_destroy(h)
        return Mappings(ro, rw, nd, pr)


class PHMMOutput:
    """table.rs:450-517."""

    def __init__(self, forward, backward):
        assert len(forward) == len(backward)
        self.forward, self.backward = forward, backward

    def to_full_prob_forward(self):
        return self.forward.full_prob()

    def to_full_prob_backward(self):
        return self.backward.full_prob()

    def to_node_freqs(self):
        fr = np.empty(self.forward.n_nodes)
        if lib().orc_output_node_freqs(self.forward._h, self.backward._h, _p(fr)):
            raise RuntimeError(_err())
        return fr

    def to_edge_and_init_freqs(self, phmm, emissions):
        """freq.rs:276-298: (edge_freqs[E], init_freqs[N]) of one read."""
        x = _bases(emissions)
        ef = np.empty(len(phmm.src)); nf = np.empty(phmm.n_nodes)
        if lib().orc_output_edge_init_freqs(phmm._h, self.forward._h, self.backward._h, _p(x), len(x), _p(ef), _p(nf)):
            raise RuntimeError(_err())
        return ef, nf

    def to_edge_freqs(self, phmm, emissions):
        return self.to_edge_and_init_freqs(phmm, emissions)[0]

    def _mapping(self, by_ratio, n_active, ratio):
        n = len(self.forward)
        cnt = np.zeros(n, np.uint64)
        if lib().orc_output_mapping(self.forward._h, self.backward._h, by_ratio, n_active, ratio, _p(cnt), None, None):
            raise RuntimeError(_err())
        tot = int(cnt.sum())
        nd = np.empty(tot, np.uint32); pr = np.empty(tot, np.float64)
        lib().orc_output_mapping(self.forward._h, self.backward._h, by_ratio, n_active, ratio, _p(cnt), _p(nd), _p(pr))
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        return Mapping([nd[off[i]:off[i + 1]] for i in range(n)], [pr[off[i]:off[i + 1]] for i in range(n)])

    def to_mapping(self, n_active_nodes):
        return self._mapping(0, n_active_nodes, 0.0)

    def to_mapping_by_score_ratio(self, max_ratio):
        return self._mapping(1, 0, max_ratio)
