"""Host-side mirror of the reference's `impl PHMMModel` / `impl PHMMOutput` surface (src/hmmv2) over the C ABI.

Same method names, argument meaning and error behaviour as the reference (forward.rs, backward.rs, freq.rs,
hint.rs, table.rs); where the reference panics this raises DbgphmmError.  All compute goes through
lib/libdbgphmm_b200.so (hand-written sm_100a CUDA); there is no CPU path — without the library or without a
GPU every compute call raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdbgphmm_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_ZERO_PROB, ERR_OOM = range(6)
MAX_ACTIVE_NODES = 400  # hmmv2/table.rs:22


class DbgphmmError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"dbgphmm_b200 status {status}: {msg}")
        self.status = status


class Params(C.Structure):
    """PHMMParams (hmmv2/params.rs:16-66); p_* are natural logs.  Layout == struct dbgphmm_params."""
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


# every symbol include/dbgphmm_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "dbgphmm_last_error", "dbgphmm_device_count", "dbgphmm_params_new", "dbgphmm_params_uniform",
    "dbgphmm_model_create", "dbgphmm_model_destroy", "dbgphmm_model_set_params", "dbgphmm_model_set_probs",
    "dbgphmm_model_set_copy_nums_batch", "dbgphmm_model_get_probs", "dbgphmm_model_n_nodes", "dbgphmm_model_n_batch",
    "dbgphmm_reads_create", "dbgphmm_reads_destroy", "dbgphmm_mappings_create", "dbgphmm_mappings_destroy",
    "dbgphmm_mappings_sizes", "dbgphmm_mappings_export", "dbgphmm_mappings_to_node_freqs", "dbgphmm_mappings_map_nodes",
    "dbgphmm_forward", "dbgphmm_backward", "dbgphmm_tables_destroy", "dbgphmm_tables_len", "dbgphmm_tables_full_prob",
    "dbgphmm_tables_row_info", "dbgphmm_tables_row_export", "dbgphmm_tables_row_top_nodes",
    "dbgphmm_output_node_freqs", "dbgphmm_output_edge_and_init_freqs", "dbgphmm_q_score_exact", "dbgphmm_output_mapping", "dbgphmm_to_full_prob_reads", "dbgphmm_run_node_freqs",
    "dbgphmm_run_node_freqs_dev", "dbgphmm_generate_mappings", "dbgphmm_launch_count", "dbgphmm_last_timing",
    "dbgphmm_reads_to_device", "dbgphmm_last_dense_kernel", "dbgphmm_model_wave_reads",
    "dbgphmm_dbg_from_text", "dbgphmm_dbg_from_file", "dbgphmm_dbg_destroy", "dbgphmm_dbg_sizes", "dbgphmm_dbg_phmm_graph",
    "dbgphmm_dbg_get_copy_nums", "dbgphmm_dbg_set_copy_nums", "dbgphmm_dbg_expand_copy_nums", "dbgphmm_dbg_to_text", "dbgphmm_dbg_to_file",
    "dbgphmm_dbg_to_model", "dbgphmm_mappings_from_map_text", "dbgphmm_mappings_from_map_file", "dbgphmm_mappings_to_map_text",
    "dbgphmm_mappings_to_map_file",
    "dbgphmm_dbg_genome_size", "dbgphmm_dbg_n_euler_circuits", "dbgphmm_euler_circuit_count", "dbgphmm_prior_normal",
    "dbgphmm_dataset_from_json_text", "dbgphmm_dataset_from_json_file", "dbgphmm_dataset_create", "dbgphmm_dataset_destroy", "dbgphmm_dataset_sizes",
    "dbgphmm_dataset_genome", "dbgphmm_dataset_reads", "dbgphmm_dataset_read_origins", "dbgphmm_dataset_params", "dbgphmm_dataset_to_json_text",
    "dbgphmm_dataset_to_json_file",
]

_lib = None


def lib():
    """Load the CUDA library.  Raises if it has not been built (python -m dbgphmm_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DbgphmmError(ERR_CUDA, f"{LIB_PATH} is missing: build it with `python -m dbgphmm_b200.build` "
                                     "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i64, dbl, ci = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_double, C.c_int
    PP = C.POINTER(Params)
    L.dbgphmm_last_error.restype = C.c_char_p
    L.dbgphmm_device_count.restype = ci
    L.dbgphmm_params_new.argtypes = [dbl, dbl, dbl, dbl, u32, u32, PP]
    L.dbgphmm_params_uniform.argtypes = [dbl, PP]
    L.dbgphmm_model_create.argtypes = [u32, u32, vp, vp, vp, vp, vp, PP, ci, u64, C.POINTER(vp)]
    L.dbgphmm_model_destroy.argtypes = [vp]
    L.dbgphmm_model_set_params.argtypes = [vp, PP]
    L.dbgphmm_model_set_probs.argtypes = [vp, vp, vp]
    L.dbgphmm_model_set_copy_nums_batch.argtypes = [vp, u32, vp, ci]
    L.dbgphmm_model_get_probs.argtypes = [vp, u32, vp, vp]
    L.dbgphmm_model_n_nodes.argtypes = [vp]; L.dbgphmm_model_n_nodes.restype = u32
    L.dbgphmm_model_n_batch.argtypes = [vp]; L.dbgphmm_model_n_batch.restype = u32
    L.dbgphmm_reads_create.argtypes = [u64, vp, vp, C.POINTER(vp)]
    L.dbgphmm_reads_destroy.argtypes = [vp]
    L.dbgphmm_reads_to_device.argtypes = [vp, vp]
    L.dbgphmm_model_wave_reads.argtypes = [vp]; L.dbgphmm_model_wave_reads.restype = u32
    L.dbgphmm_mappings_create.argtypes = [u64, vp, vp, vp, vp, C.POINTER(vp)]
    L.dbgphmm_mappings_destroy.argtypes = [vp]
    L.dbgphmm_mappings_sizes.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.dbgphmm_mappings_export.argtypes = [vp, vp, vp, vp, vp]
    L.dbgphmm_mappings_to_node_freqs.argtypes = [vp, u32, vp]
    L.dbgphmm_mappings_map_nodes.argtypes = [vp, u32, vp, vp, C.POINTER(vp)]
    L.dbgphmm_forward.argtypes = [vp, vp, u64, ci, vp, u64, C.POINTER(vp)]
    L.dbgphmm_backward.argtypes = [vp, vp, u64, ci, vp, u64, vp, C.POINTER(vp)]
    L.dbgphmm_tables_destroy.argtypes = [vp]
    L.dbgphmm_tables_len.argtypes = [vp]; L.dbgphmm_tables_len.restype = u64
    L.dbgphmm_tables_full_prob.argtypes = [vp, C.POINTER(dbl)]
    L.dbgphmm_tables_row_info.argtypes = [vp, i64, vp, vp]
    L.dbgphmm_tables_row_export.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.dbgphmm_tables_row_top_nodes.argtypes = [vp, i64, ci, u32, dbl, vp, C.POINTER(u32)]
    L.dbgphmm_output_node_freqs.argtypes = [vp, vp, vp, vp]
    L.dbgphmm_output_edge_and_init_freqs.argtypes = [vp, vp, vp, vp, vp]
    L.dbgphmm_q_score_exact.argtypes = [vp, u32, vp, vp, vp]
    L.dbgphmm_output_mapping.argtypes = [vp, vp, vp, ci, u32, dbl, C.POINTER(vp)]
    L.dbgphmm_to_full_prob_reads.argtypes = [vp, vp, vp, ci, vp, vp]
    L.dbgphmm_run_node_freqs.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp]
    L.dbgphmm_run_node_freqs_dev.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp]
    L.dbgphmm_generate_mappings.argtypes = [vp, vp, vp, ci, C.POINTER(vp)]
    L.dbgphmm_launch_count.argtypes = [ci]; L.dbgphmm_launch_count.restype = u64
    L.dbgphmm_last_timing.argtypes = [vp, C.POINTER(u64)]
    L.dbgphmm_last_dense_kernel.argtypes = [C.POINTER(dbl), C.POINTER(u64), C.POINTER(u64)]
    L.dbgphmm_dbg_from_text.argtypes = [C.c_char_p, u64, C.POINTER(vp)]
    L.dbgphmm_dbg_from_file.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.dbgphmm_dbg_destroy.argtypes = [vp]
    L.dbgphmm_dbg_sizes.argtypes = [vp, vp]
    L.dbgphmm_dbg_phmm_graph.argtypes = [vp, vp, vp, vp, vp, vp]
    L.dbgphmm_dbg_get_copy_nums.argtypes = [vp, vp]
    L.dbgphmm_dbg_set_copy_nums.argtypes = [vp, vp]
    L.dbgphmm_dbg_expand_copy_nums.argtypes = [vp, u32, vp, vp]
    L.dbgphmm_dbg_to_text.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.dbgphmm_dbg_to_file.argtypes = [vp, C.c_char_p]
    L.dbgphmm_dbg_to_model.argtypes = [vp, PP, ci, ci, u64, C.POINTER(vp)]
    L.dbgphmm_mappings_from_map_text.argtypes = [C.c_char_p, u64, C.POINTER(vp)]
    L.dbgphmm_mappings_from_map_file.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.dbgphmm_mappings_to_map_text.argtypes = [vp, vp, vp, vp, u64, C.POINTER(u64)]
    L.dbgphmm_mappings_to_map_file.argtypes = [vp, vp, vp, C.c_char_p]
    L.dbgphmm_dbg_genome_size.argtypes = [vp, u32, vp, vp]
    L.dbgphmm_dbg_n_euler_circuits.argtypes = [vp, u32, vp, vp]
    L.dbgphmm_euler_circuit_count.argtypes = [u32, u64, vp, vp, vp, ci, C.POINTER(dbl)]
    L.dbgphmm_prior_normal.argtypes = [dbl, dbl, dbl, C.POINTER(dbl)]
    L.dbgphmm_dataset_from_json_text.argtypes = [C.c_char_p, u64, C.POINTER(vp)]
    L.dbgphmm_dataset_from_json_file.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.dbgphmm_dataset_create.argtypes = [u32, vp, vp, vp, u64, u64, vp, vp, vp, vp, vp, PP, C.POINTER(vp)]
    L.dbgphmm_dataset_destroy.argtypes = [vp]
    L.dbgphmm_dataset_sizes.argtypes = [vp, vp]
    L.dbgphmm_dataset_genome.argtypes = [vp, vp, vp, vp]
    L.dbgphmm_dataset_reads.argtypes = [vp, C.POINTER(vp)]
    L.dbgphmm_dataset_read_origins.argtypes = [vp, vp, vp, vp, vp, vp]
    L.dbgphmm_dataset_params.argtypes = [vp, PP]
    L.dbgphmm_dataset_to_json_text.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.dbgphmm_dataset_to_json_file.argtypes = [vp, C.c_char_p]
    for s in SYMBOLS:
        getattr(L, s)  # fail loudly if the library does not export a declared symbol
    _lib = L
    return L


def _check(st):
    if st != OK:
        raise DbgphmmError(st, lib().dbgphmm_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count():
    return int(lib().dbgphmm_device_count())


def launch_count(reset=False):
    return int(lib().dbgphmm_launch_count(1 if reset else 0))


def last_timing():
    """(dense_ms, sparse_ms, product_ms, total_ms, dense_cells) of the last bulk call, from CUDA events."""
    ms = (C.c_double * 4)()
    cells = C.c_uint64(0)
    lib().dbgphmm_last_timing(ms, C.byref(cells))
    return ms[0], ms[1], ms[2], ms[3], int(cells.value)


def last_dense_kernel():
    """(summed ms, launches, cells) of the dense row-step kernel launches of the last bulk call."""
    ms = C.c_double(0); n = C.c_uint64(0); cells = C.c_uint64(0)
    lib().dbgphmm_last_dense_kernel(C.byref(ms), C.byref(n), C.byref(cells))
    return ms.value, int(n.value), int(cells.value)


def params_uniform(p):
    """PHMMParams::uniform (params.rs:116-124)."""
    q = Params()
    lib().dbgphmm_params_uniform(float(p), C.byref(q))
    return q


def params_new(p_mismatch, p_gap_open, p_gap_ext, p_end, n_active_nodes, n_warmup):
    """PHMMParams::new (params.rs:73-112)."""
    q = Params()
    lib().dbgphmm_params_new(p_mismatch, p_gap_open, p_gap_ext, p_end, n_active_nodes, n_warmup, C.byref(q))
    return q


def _bases(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), np.uint8).copy()
    return np.ascontiguousarray(x, np.uint8)


class Reads:
    """ReadCollection (common/collection.rs:131): CSR of uppercase ACGT reads."""

    def __init__(self, seqs):
        seqs = [_bases(s) for s in seqs]
        self.offsets = np.zeros(len(seqs) + 1, np.uint64)
        self.offsets[1:] = np.cumsum([len(s) for s in seqs])
        self.bases = np.ascontiguousarray(np.concatenate(seqs)) if seqs else np.zeros(0, np.uint8)
        h = C.c_void_p()
        _check(lib().dbgphmm_reads_create(len(seqs), _p(self.offsets), _p(self.bases), C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dbgphmm_reads_destroy(self._h)
            self._h = None

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, r):
        return self.bases[int(self.offsets[r]):int(self.offsets[r + 1])]

    def total_bases(self):
        return int(self.offsets[-1])


class Mapping:
    """hint.rs:27-30 for one read: per-base node lists and their ln probabilities."""

    def __init__(self, nodes, probs):
        self.nodes, self.probs = nodes, probs

    def __len__(self):
        return len(self.nodes)


class Mappings:
    """hint.rs:150-152 : CSR over reads -> bases -> (node, ln prob); owns a C handle."""

    def __init__(self, read_off, row_off, nodes, probs, handle=None):
        self.read_off = np.ascontiguousarray(read_off, np.uint64)
        self.row_off = np.ascontiguousarray(row_off, np.uint64)
        self.nodes = np.ascontiguousarray(nodes, np.uint32)
        self.probs = np.ascontiguousarray(probs, np.float64)
        if handle is None:
            handle = C.c_void_p()
            _check(lib().dbgphmm_mappings_create(len(self.read_off) - 1, _p(self.read_off), _p(self.row_off), _p(self.nodes),
                                                 _p(self.probs), C.byref(handle)))
        self._h = handle

    @staticmethod
    def _from_handle(h):
        nr, nrow, nent = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().dbgphmm_mappings_sizes(h, C.byref(nr), C.byref(nrow), C.byref(nent)))
        ro = np.zeros(nr.value + 1, np.uint64); rw = np.zeros(nrow.value + 1, np.uint64)
        nd = np.zeros(nent.value, np.uint32); pr = np.zeros(nent.value, np.float64)
        _check(lib().dbgphmm_mappings_export(h, _p(ro), _p(rw), _p(nd), _p(pr)))
        return Mappings(ro, rw, nd, pr, handle=h)

    @staticmethod
    def from_list(maps):
        read_off = [0]; row_off = [0]; nodes = []; probs = []
        for m in maps:
            for ns, ps in zip(m.nodes, m.probs):
                nodes.append(np.asarray(ns, np.uint32)); probs.append(np.asarray(ps, np.float64))
                row_off.append(row_off[-1] + len(ns))
            read_off.append(read_off[-1] + len(m.nodes))
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
        return Mappings(np.array(read_off, np.uint64), np.array(row_off, np.uint64), cat(nodes, np.uint32), cat(probs, np.float64))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dbgphmm_mappings_destroy(self._h)
            self._h = None

    def n_reads(self):
        return len(self.read_off) - 1

    def __getitem__(self, r):
        a, b = int(self.read_off[r]), int(self.read_off[r + 1])
        ns = [self.nodes[int(self.row_off[i]):int(self.row_off[i + 1])] for i in range(a, b)]
        ps = [self.probs[int(self.row_off[i]):int(self.row_off[i + 1])] for i in range(a, b)]
        return Mapping(ns, ps)

    def slice(self, lo, hi):
        """The Mappings of reads lo .. hi-1 (a shard of the read set keeps its own mappings)."""
        if not (0 <= lo <= hi <= self.n_reads()):
            raise DbgphmmError(ERR_INVALID, "mappings slice out of range")
        r0, r1 = int(self.read_off[lo]), int(self.read_off[hi])
        e0, e1 = int(self.row_off[r0]), int(self.row_off[r1])
        return Mappings(self.read_off[lo:hi + 1] - np.uint64(r0), self.row_off[r0:r1 + 1] - np.uint64(e0), self.nodes[e0:e1], self.probs[e0:e1])

    @staticmethod
    def from_map_str(text):
        """MultiDbg::from_map_str (multi_dbg/output.rs:589-591)."""
        b = text.encode() if isinstance(text, str) else bytes(text)
        h = C.c_void_p()
        _check(lib().dbgphmm_mappings_from_map_text(b, len(b), C.byref(h)))
        return Mappings._from_handle(h)

    @staticmethod
    def from_map_file(path):
        """MultiDbg::from_map_file_raw (multi_dbg/output.rs:612-623); .map.gz / .mpz are gzip."""
        h = C.c_void_p()
        _check(lib().dbgphmm_mappings_from_map_file(os.fsencode(path), C.byref(h)))
        return Mappings._from_handle(h)

    def to_map_string(self, reads, dbg=None):
        """MultiDbg::to_map_string (multi_dbg/output.rs:485-489)."""
        need = C.c_uint64()
        dh = dbg._h if dbg is not None else None
        _check(lib().dbgphmm_mappings_to_map_text(self._h, reads._h, dh, None, 0, C.byref(need)))
        buf = C.create_string_buffer(max(1, need.value))
        _check(lib().dbgphmm_mappings_to_map_text(self._h, reads._h, dh, buf, need.value, C.byref(need)))
        return buf.raw[:need.value].decode()

    def to_map_file(self, path, reads, dbg=None):
        """MultiDbg::to_map_file (multi_dbg/output.rs:461-472)."""
        _check(lib().dbgphmm_mappings_to_map_file(self._h, reads._h, dbg._h if dbg is not None else None, os.fsencode(path)))

    def to_node_freqs(self, n_nodes):
        """Mappings::to_node_freqs (hint.rs:161-171) == MultiDbg::mappings_to_freqs (multi_dbg/draft.rs:201-212)."""
        f = np.zeros(n_nodes)
        _check(lib().dbgphmm_mappings_to_node_freqs(self._h, n_nodes, _p(f)))
        return f

    def map_nodes(self, node_map, n_nodes_before=None):
        """Mapping::map_nodes (hint.rs:66-88) applied to every read.  node_map: callable node -> list of nodes (as in the
        reference), or a (map_off, map_to) CSR pair over the old node ids."""
        if callable(node_map):
            if n_nodes_before is None:
                n_nodes_before = int(self.nodes.max()) + 1 if len(self.nodes) else 0
            off = np.zeros(n_nodes_before + 1, np.uint64); to = []
            for v in range(n_nodes_before):
                to.extend(int(w) for w in node_map(v)); off[v + 1] = len(to)
            to = np.asarray(to, np.uint32)
        else:
            off = np.ascontiguousarray(node_map[0], np.uint64); to = np.ascontiguousarray(node_map[1], np.uint32)
            n_nodes_before = len(off) - 1
        h = C.c_void_p()
        _check(lib().dbgphmm_mappings_map_nodes(self._h, n_nodes_before, _p(off), _p(to), C.byref(h)))
        return Mappings._from_handle(h)


class Row:
    """One PHMMTable (table.rs:42-73), natural logs."""
    __slots__ = ("is_dense", "ids", "m", "i", "ids_d", "d", "mb", "ib", "e")

    def n_active_nodes(self):
        """table.rs:108-110: entries of m (all N for a dense row)."""
        return len(self.m)

    def diff(self, other, n_nodes):
        return table_diff(self, other, n_nodes)

    def log_diff(self, other, n_nodes):
        return table_log_diff(self, other, n_nodes)

    def to_nodevec(self, n_nodes):
        return self.merged(n_nodes)

    def merged(self, n_nodes):
        """PHMMTable::to_nodevec (table.rs:199-211) as a dense ln array (absent = -inf)."""
        with np.errstate(divide="ignore", invalid="ignore"):
            if self.is_dense:
                return np.logaddexp(np.logaddexp(self.m, self.i), self.d)
            v = np.full(n_nodes, -np.inf)
            v[self.ids] = np.logaddexp(self.m, self.i)
            v[self.ids_d] = np.logaddexp(v[self.ids_d], self.d)
        return v


def table_diff(a, b, n_nodes):
    """PHMMTable::diff (table.rs:187-195): sum of |p_a - p_b| over the M / I / D states of every node and mb, ib, e -- the measure
    the reference's own dense-vs-sparse tests use (forward.rs:621-638, tests/hmm.rs:122-214).  a, b: rows with is_dense / ids / m ..."""
    tot = 0.0
    with np.errstate(over="ignore"):
        for x, y in zip(_dense_states(a, n_nodes), _dense_states(b, n_nodes)):
            tot += float(np.abs(np.exp(x) - np.exp(y)).sum())
        for name in ("mb", "ib", "e"):
            tot += abs(float(np.exp(getattr(a, name))) - float(np.exp(getattr(b, name))))
    return tot


def table_log_diff(a, b, n_nodes):
    """PHMMTable::log_diff (table.rs:174-182) over Prob::log_diff (prob.rs:110-124): sum of |ln p_a - ln p_b|; two zeros differ by 0,
    a zero and a non-zero by +inf."""
    def ld(x, y):
        x = np.atleast_1d(np.asarray(x, np.float64)); y = np.atleast_1d(np.asarray(y, np.float64))
        zx, zy = np.isneginf(x), np.isneginf(y)
        with np.errstate(invalid="ignore"):
            d = np.where(zx & zy, 0.0, np.where(zx | zy, np.inf, np.abs(x - y)))
        return float(d.sum())
    tot = sum(ld(x, y) for x, y in zip(_dense_states(a, n_nodes), _dense_states(b, n_nodes)))
    return tot + sum(ld(getattr(a, name), getattr(b, name)) for name in ("mb", "ib", "e"))


class QScore:
    """QScore (hmmv2/q.rs:14-52): the three terms q_score_exact returns."""
    __slots__ = ("init", "trans", "prior")

    def __init__(self, init, trans, prior):
        self.init, self.trans, self.prior = init, trans, prior

    def total(self):
        return self.init + self.trans + self.prior

    def sub(self, other):
        return QScore(self.init - other.init, self.trans - other.trans, self.prior - other.prior)


class PHMMTables:
    """PHMMTables (table.rs:365-435) of one read, resident on the GPU."""

    def __init__(self, model, handle):
        self._model, self._h = model, handle
        self.n_nodes = model.n_nodes

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dbgphmm_tables_destroy(self._h)
            self._h = None

    def __len__(self):
        return int(lib().dbgphmm_tables_len(self._h))

    def n_emissions(self):
        return len(self)

    def full_prob(self):
        v = C.c_double()
        _check(lib().dbgphmm_tables_full_prob(self._h, C.byref(v)))
        return v.value

    def row(self, i):
        """tables[i]; i = -1 is init_table."""
        info = np.zeros(3, np.uint64); sc = np.zeros(3, np.float64)
        _check(lib().dbgphmm_tables_row_info(self._h, i, _p(info), _p(sc)))
        r = Row()
        r.is_dense = bool(info[0]); r.mb, r.ib, r.e = (float(v) for v in sc)
        if r.is_dense:
            N = self.n_nodes
            r.ids = r.ids_d = None
            r.m = np.empty(N); r.i = np.empty(N); r.d = np.empty(N)
            _check(lib().dbgphmm_tables_row_export(self._h, i, None, _p(r.m), _p(r.i), None, _p(r.d)))
        else:
            nm, nd = int(info[1]), int(info[2])
            r.ids = np.empty(nm, np.uint32); r.m = np.empty(nm); r.i = np.empty(nm)
            r.ids_d = np.empty(nd, np.uint32); r.d = np.empty(nd)
            _check(lib().dbgphmm_tables_row_export(self._h, i, _p(r.ids), _p(r.m), _p(r.i), _p(r.ids_d), _p(r.d)))
        return r

    def top_nodes(self, i, k):
        out = np.empty(MAX_ACTIVE_NODES, np.uint32); n = C.c_uint32()
        _check(lib().dbgphmm_tables_row_top_nodes(self._h, i, 0, k, 0.0, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def table(self, i):
        return self.row(i)

    def first_table(self):
        return self.row(0)

    def last_table(self):
        return self.row(len(self) - 1)

    def filled_nodes(self, i):
        """PHMMTable::filled_nodes (table.rs:117-123): None for a dense row, else the top-|m entries| of the merged row."""
        info = np.zeros(3, np.uint64); sc = np.zeros(3, np.float64)
        _check(lib().dbgphmm_tables_row_info(self._h, i, _p(info), _p(sc)))
        return None if info[0] else self.top_nodes(i, int(info[1]))

    def top_nodes_with_prob(self, i, k):
        """table.rs:153-159: [(node, ln merged prob)] of the top k nodes of row i."""
        v = self.row(i).merged(self.n_nodes)
        return [(int(n), float(v[n])) for n in self.top_nodes(i, k)]

    def top_nodes_with_prob_by_score_ratio(self, i, ratio):
        """table.rs:163-169."""
        v = self.row(i).merged(self.n_nodes)
        return [(int(n), float(v[n])) for n in self.top_nodes_by_score_ratio(i, ratio)]

    def top_nodes_by_score_ratio(self, i, ratio):
        out = np.empty(MAX_ACTIVE_NODES, np.uint32); n = C.c_uint32()
        _check(lib().dbgphmm_tables_row_top_nodes(self._h, i, 1, 0, ratio, _p(out), C.byref(n)))
        return out[:n.value].copy()


def _dense_states(row, n_nodes):
    """(m, i, d) of a Row as dense ln arrays; entries a sparse row does not store are the SparseVec default, ln 0."""
    if row.is_dense:
        return row.m, row.i, row.d
    m = np.full(n_nodes, -np.inf); i = np.full(n_nodes, -np.inf); d = np.full(n_nodes, -np.inf)
    m[row.ids] = row.m; i[row.ids] = row.i; d[row.ids_d] = row.d
    return m, i, d


def table_merged(tables, is_forward, merged_index):
    """PHMMTables::table_merged (table.rs:414-434), 0 <= merged_index <= n: Forward[0] and Backward[n] are the init tables."""
    n = len(tables)
    if is_forward:
        return tables.row(-1 if merged_index == 0 else merged_index - 1)
    return tables.row(-1 if merged_index >= n else merged_index)


def emit_probs(forward, backward, merged_index):
    """PHMMOutput::to_emit_probs (table.rs:500-505): (F[i] * B[i]) / P(x) with P from the forward tables, as a dense Row of
    natural logs.  Host-side composition over exported rows (inspection / tests; the bulk products run on the device).  A state
    one side does not store is that side's default (zero), so the product is zero there -- the values of the reference's
    SparseVec product, in dense storage.  `forward` / `backward`: anything with row(i), len() and full_prob()."""
    p = forward.full_prob()
    if np.isneginf(p):
        raise DbgphmmError(4, "P(read) == 0: to_emit_probs would be NaN (table.rs:500-505)")
    n_nodes = forward.n_nodes
    f = table_merged(forward, True, merged_index); b = table_merged(backward, False, merged_index)
    r = Row()
    r.is_dense = True; r.ids = r.ids_d = None
    (fm, fi, fd), (bm, bi, bd) = _dense_states(f, n_nodes), _dense_states(b, n_nodes)
    r.m, r.i, r.d = fm + bm - p, fi + bi - p, fd + bd - p
    r.mb, r.ib, r.e = float(f.mb + b.mb - p), float(f.ib + b.ib - p), float(f.e + b.e - p)
    return r


def state_probs(forward, backward):
    """PHMMOutput::to_state_probs (freq.rs:237-239): the emit probs of merged index 0..n added up (Prob +, prob.rs:181-197)."""
    acc = None
    for i in range(len(forward) + 1):
        t = emit_probs(forward, backward, i)
        if acc is None:
            acc = t
            continue
        with np.errstate(invalid="ignore"):
            for name in ("m", "i", "d"):
                setattr(acc, name, np.logaddexp(getattr(acc, name), getattr(t, name)))
            acc.mb, acc.ib, acc.e = (float(np.logaddexp(a, b)) for a, b in ((acc.mb, t.mb), (acc.ib, t.ib), (acc.e, t.e)))
    return acc


class PHMMOutput:
    """PHMMOutput (table.rs:450-517)."""

    def __init__(self, model, forward, backward):
        self._model, self.forward, self.backward = model, forward, backward

    def n_emissions(self):
        return len(self.forward)

    def to_full_prob_forward(self):
        return self.forward.full_prob()

    def to_full_prob_backward(self):
        return self.backward.full_prob()

    def to_emit_probs(self, merged_index):
        """table.rs:500-505, a dense Row (host-side composition over exported rows, see emit_probs)."""
        return emit_probs(self.forward, self.backward, merged_index)

    def iter_emit_probs(self):
        """freq.rs:226-229: merged index 0..=n."""
        return (self.to_emit_probs(i) for i in range(self.n_emissions() + 1))

    def to_state_probs(self):
        """freq.rs:237-239.  to_node_freqs() is exp(merged m + i + d) of this table, computed on the device."""
        return state_probs(self.forward, self.backward)

    def to_node_freqs(self):
        f = np.empty(self._model.n_nodes)
        _check(lib().dbgphmm_output_node_freqs(self._model._h, self.forward._h, self.backward._h, _p(f)))
        return f

    def to_edge_and_init_freqs(self):
        """PHMMOutput::to_edge_and_init_freqs (freq.rs:276-298): (edge_freqs[n_edges] in EdgeIndex order, init_freqs[n_nodes]).
        The reference takes (phmm, emissions) again; the tables here remember both."""
        ef = np.empty(len(self._model.src)); nf = np.empty(self._model.n_nodes)
        _check(lib().dbgphmm_output_edge_and_init_freqs(self._model._h, self.forward._h, self.backward._h, _p(ef), _p(nf)))
        return ef, nf

    def to_edge_freqs(self):
        """freq.rs:302-309."""
        return self.to_edge_and_init_freqs()[0]

    def _mapping(self, by_ratio, n_active, ratio):
        h = C.c_void_p()
        _check(lib().dbgphmm_output_mapping(self._model._h, self.forward._h, self.backward._h, by_ratio, n_active, ratio, C.byref(h)))
        return Mappings._from_handle(h)[0]

    def to_mapping(self, n_active_nodes):
        return self._mapping(0, n_active_nodes, 0.0)

    def to_mapping_by_score_ratio(self, max_ratio):
        return self._mapping(1, 0, max_ratio)


FWD_DENSE, FWD_SPARSE, FWD_SPARSE_RATIO, FWD_MAPPING = range(4)
BWD_DENSE, BWD_SPARSE, BWD_MAPPING, BWD_BY_FORWARD = range(4)
RUN_MODES = {"dense": 0, "sparse": 1, "sparse_adaptive": 2, "with_mapping": 3}


def _padd(a, b):
    """Prob + Prob (prob.rs:181-197)."""
    x, y = (a, b) if a >= b else (b, a)
    if np.isneginf(y):
        return x
    if x == y:
        return x + float(np.log(2.0))
    return x + float(np.log1p(np.exp(y - x)))


class Score:
    """Score (multi_dbg/posterior.rs:164-208): natural logs; p() = P(R|X) P(G) #circuits."""
    __slots__ = ("likelihood", "prior", "genome_size", "n_euler_circuits")

    def __init__(self, likelihood, prior, genome_size, n_euler_circuits):
        self.likelihood, self.prior, self.genome_size, self.n_euler_circuits = likelihood, prior, genome_size, n_euler_circuits

    def p(self):
        return self.likelihood + self.prior + self.n_euler_circuits

    def __repr__(self):
        return f"Score(likelihood={self.likelihood}, prior={self.prior}, genome_size={self.genome_size}, n_euler_circuits={self.n_euler_circuits})"


class Posterior:
    """Posterior (multi_dbg/posterior.rs:82-161): the distinct copy-number vectors seen so far with their scores."""

    def __init__(self):
        self.samples = []            # (copy_nums tuple, Score) in insertion order
        self._p = -np.inf

    def contains(self, copy_nums):
        return self.find(copy_nums) is not None

    def find(self, copy_nums):
        key = tuple(int(c) for c in copy_nums)
        for k, sc in self.samples:
            if k == key:
                return sc
        return None

    def add(self, copy_nums, score):
        """posterior.rs:93-98: a copy-number vector counts once."""
        if not self.contains(copy_nums):
            self._p = _padd(self._p, score.p())
            self.samples.append((tuple(int(c) for c in copy_nums), score))

    def p(self):
        return self._p

    def max_sample(self):
        """posterior.rs:113-118 (max_by_key: the LAST of equal maxima, like Iterator::max_by_key)."""
        best = None
        for k, sc in self.samples:
            if best is None or sc.p() >= best[1].p():
                best = (k, sc)
        return best

    def max_copy_nums(self):
        return np.array(self.max_sample()[0], np.uint32)

    def p_edge_x(self, edge, x):
        """P(X[edge] = x | R) (posterior.rs:141-143 over hist.rs:48-71), natural log."""
        z, px = -np.inf, -np.inf
        for k, sc in self.samples:
            w = sc.p() - self._p
            z = _padd(z, w)
            if k[edge] == x:
                px = _padd(px, w)
        return px - z if not np.isneginf(px) else -np.inf


def euler_circuit_count(n_nodes, edges, allow_multiple_component):
    """graph/euler.rs:94-123: ln of the number of Euler circuits of a multigraph given as (source, target, multiplicity) triples."""
    e = np.ascontiguousarray(edges, np.uint32).reshape(-1, 3)
    s, t, w = (np.ascontiguousarray(e[:, i]) for i in range(3))
    v = C.c_double()
    _check(lib().dbgphmm_euler_circuit_count(n_nodes, len(e), _p(s), _p(t), _p(w), int(allow_multiple_component), C.byref(v)))
    return v.value


class Dataset:
    """e2e::Dataset (e2e.rs:31-130): genome, genome_size, positioned reads and the PHMM parameters they were sampled with, in the
    reference's JSON form.  `reads()` is the ReadCollection the hot path takes."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dbgphmm_dataset_destroy(self._h)
            self._h = None

    @staticmethod
    def from_json_str(text):
        b = text.encode() if isinstance(text, str) else bytes(text)
        h = C.c_void_p()
        _check(lib().dbgphmm_dataset_from_json_text(b, len(b), C.byref(h)))
        return Dataset(h)

    @staticmethod
    def from_json_file(path):
        h = C.c_void_p()
        _check(lib().dbgphmm_dataset_from_json_file(os.fsencode(path), C.byref(h)))
        return Dataset(h)

    @staticmethod
    def new(haplotypes, styles, reads, param, genome_size=None, revcomp=None, origins=None):
        """haplotypes: sequences; styles: 'C' / 'L' / 'F' per haplotype; reads: sequences; origins: per read a list of (hap, pos) or
        None for an inserted base (GenomeGraphPos, genome_graph.rs:60-115)."""
        haps = [_bases(h) for h in haplotypes]
        hoff = np.zeros(len(haps) + 1, np.uint64); hoff[1:] = np.cumsum([len(h) for h in haps])
        hb = np.ascontiguousarray(np.concatenate(haps)) if haps else np.zeros(0, np.uint8)
        st = np.frombuffer("".join(styles).encode(), np.uint8).copy()
        rs = [_bases(r) for r in reads]
        roff = np.zeros(len(rs) + 1, np.uint64); roff[1:] = np.cumsum([len(r) for r in rs])
        rb = np.ascontiguousarray(np.concatenate(rs)) if rs else np.zeros(0, np.uint8)
        rv = None if revcomp is None else np.ascontiguousarray(revcomp, np.uint8)
        oh = op = None
        if origins is not None:
            oh = np.array([(-1 if o is None else o[0]) for r in origins for o in r], np.int64)
            op = np.array([(0 if o is None else o[1]) for r in origins for o in r], np.uint64)
            assert len(oh) == len(rb)
        gs = int(hoff[-1]) if genome_size is None else int(genome_size)
        h = C.c_void_p()
        q = param.copy()
        _check(lib().dbgphmm_dataset_create(len(haps), _p(hoff), _p(hb), _p(st), gs, len(rs), _p(roff), _p(rb), _p(rv), _p(oh), _p(op), C.byref(q), C.byref(h)))
        return Dataset(h)

    def _sizes(self):
        s = np.zeros(5, np.uint64)
        _check(lib().dbgphmm_dataset_sizes(self._h, _p(s)))
        return [int(x) for x in s]

    def genome_size(self):
        return self._sizes()[4]

    def genome(self):
        """-> list of (style, bases)"""
        n, nb = self._sizes()[:2]
        off = np.zeros(n + 1, np.uint64); b = np.zeros(nb, np.uint8); st = np.zeros(n, np.uint8)
        _check(lib().dbgphmm_dataset_genome(self._h, _p(off), _p(b), _p(st)))
        return [(chr(st[i]), b[int(off[i]):int(off[i + 1])].tobytes()) for i in range(n)]

    def reads(self):
        """-> Reads (the sequences; positions are dropped like ReadCollection::to_fasta does)"""
        _, _, n, nb, _ = self._sizes()
        off = np.zeros(n + 1, np.uint64); b = np.zeros(nb, np.uint8)
        _check(lib().dbgphmm_dataset_read_origins(self._h, _p(off), _p(b), None, None, None))
        return Reads([b[int(off[i]):int(off[i + 1])] for i in range(n)])

    def read_origins(self):
        """-> (revcomp flags, per read a list of (hap, pos) or None)"""
        _, _, n, nb, _ = self._sizes()
        off = np.zeros(n + 1, np.uint64); rv = np.zeros(n, np.uint8); oh = np.zeros(nb, np.int64); op = np.zeros(nb, np.uint64)
        _check(lib().dbgphmm_dataset_read_origins(self._h, _p(off), None, _p(rv), _p(oh), _p(op)))
        out = [[None if oh[j] < 0 else (int(oh[j]), int(op[j])) for j in range(int(off[i]), int(off[i + 1]))] for i in range(n)]
        return [bool(x) for x in rv], out

    def params(self):
        q = Params()
        _check(lib().dbgphmm_dataset_params(self._h, C.byref(q)))
        return q

    def coverage(self):
        """Dataset::coverage (e2e.rs:72-74): total bases of the reads / genome_size"""
        return self._sizes()[3] / max(1, self.genome_size())

    def to_json_string(self):
        need = C.c_uint64(0)
        _check(lib().dbgphmm_dataset_to_json_text(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(int(need.value))
        _check(lib().dbgphmm_dataset_to_json_text(self._h, buf, need.value, C.byref(need)))
        return buf.raw[:need.value].decode()

    def to_json_file(self, path):
        _check(lib().dbgphmm_dataset_to_json_file(self._h, os.fsencode(path)))


class MultiDbg:
    """The part of MultiDbg (multi_dbg.rs:170-186) a DBG file carries: compact edges with their k-mers, copy numbers and
    full-graph edge ids.  Host only (no GPU needed) except to_phmm*."""

    def __init__(self, handle):
        self._h = handle
        sz = (C.c_uint32 * 6)()
        _check(lib().dbgphmm_dbg_sizes(handle, sz))
        self._k, self.n_nodes_full, self.n_edges_full, self.n_nodes_compact, self.n_edges_compact, self.n_phmm_edges = [int(x) for x in sz]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dbgphmm_dbg_destroy(self._h)
            self._h = None

    @staticmethod
    def from_dbg_str(text):
        """MultiDbg::from_dbg_str (multi_dbg/output.rs:340-342)."""
        b = text.encode() if isinstance(text, str) else bytes(text)
        h = C.c_void_p()
        _check(lib().dbgphmm_dbg_from_text(b, len(b), C.byref(h)))
        return MultiDbg(h)

    @staticmethod
    def from_dbg_file(path):
        """MultiDbg::from_dbg_file (multi_dbg/output.rs:346-357); .dbg.gz / .dbz are gzip."""
        h = C.c_void_p()
        _check(lib().dbgphmm_dbg_from_file(os.fsencode(path), C.byref(h)))
        return MultiDbg(h)

    def k(self):
        return self._k

    def to_dbg_string(self):
        """MultiDbg::to_dbg_string (multi_dbg/output.rs:129-133)."""
        need = C.c_uint64()
        _check(lib().dbgphmm_dbg_to_text(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(max(1, need.value))
        _check(lib().dbgphmm_dbg_to_text(self._h, buf, need.value, C.byref(need)))
        return buf.raw[:need.value].decode()

    def to_dbg_file(self, path):
        _check(lib().dbgphmm_dbg_to_file(self._h, os.fsencode(path)))

    def phmm_graph(self):
        """to_seq_graph (multi_dbg.rs:1370-1390): (edge_src, edge_dst, emission, node copy numbers, compact edge of each PHMM node)."""
        src = np.zeros(self.n_phmm_edges, np.uint32); dst = np.zeros(self.n_phmm_edges, np.uint32)
        em = np.zeros(self.n_edges_full, np.uint8); cn = np.zeros(self.n_edges_full, np.uint32); ce = np.zeros(self.n_edges_full, np.uint32)
        _check(lib().dbgphmm_dbg_phmm_graph(self._h, _p(src), _p(dst), _p(em), _p(cn), _p(ce)))
        return src, dst, em, cn, ce

    def get_copy_nums(self):
        x = np.zeros(self.n_edges_compact, np.uint32)
        _check(lib().dbgphmm_dbg_get_copy_nums(self._h, _p(x)))
        return x

    def set_copy_nums(self, copy_nums):
        x = np.ascontiguousarray(copy_nums, np.uint32)
        if x.shape != (self.n_edges_compact,):
            raise DbgphmmError(ERR_INVALID, "copy_nums must have one entry per compact edge")
        _check(lib().dbgphmm_dbg_set_copy_nums(self._h, _p(x)))

    def expand_copy_nums(self, candidates):
        """[B][n_edges_compact] candidate copy numbers -> [B][n_edges_full] for PHMMModel.set_copy_nums_batch."""
        x = np.ascontiguousarray(np.atleast_2d(candidates), np.uint32)
        if x.shape[1] != self.n_edges_compact:
            raise DbgphmmError(ERR_INVALID, "candidates must be [B][n_edges_compact]")
        out = np.zeros((x.shape[0], self.n_edges_full), np.uint32)
        _check(lib().dbgphmm_dbg_expand_copy_nums(self._h, x.shape[0], _p(x), _p(out)))
        return out

    # ---- the terms of to_score beside the likelihood (multi_dbg/posterior.rs:225-277)
    def _candidates(self, candidates):
        if candidates is None:
            return None, 1
        x = np.ascontiguousarray(np.atleast_2d(candidates), np.uint32)
        if x.shape[1] != self.n_edges_compact:
            raise DbgphmmError(ERR_INVALID, "candidates must be [B][n_edges_compact]")
        return x, x.shape[0]

    def genome_size(self, candidates=None):
        """MultiDbg::genome_size (multi_dbg.rs:1018-1028) of the current copy numbers (an int) or of every candidate ([B])."""
        x, b = self._candidates(candidates)
        out = np.zeros(b, np.uint64)
        _check(lib().dbgphmm_dbg_genome_size(self._h, b, _p(x), _p(out)))
        return int(out[0]) if candidates is None else out

    def n_euler_circuits(self, candidates=None):
        """MultiDbg::n_euler_circuits (multi_dbg.rs:831-837): ln of the number of Euler circuits (-inf: none)."""
        x, b = self._candidates(candidates)
        out = np.zeros(b, np.float64)
        _check(lib().dbgphmm_dbg_n_euler_circuits(self._h, b, _p(x), _p(out)))
        return float(out[0]) if candidates is None else out

    def to_prior(self, genome_size_expected, genome_size_sigma, candidates=None):
        """MultiDbg::to_prior (posterior.rs:225-231): ln Normal(genome_size; expected, sigma)."""
        g = np.atleast_1d(self.genome_size(candidates))
        out = np.zeros(len(g)); v = C.c_double()
        for i, x in enumerate(g):
            _check(lib().dbgphmm_prior_normal(float(x), float(genome_size_expected), float(genome_size_sigma), C.byref(v)))
            out[i] = v.value
        return float(out[0]) if candidates is None else out

    def to_scores(self, phmm, reads, mappings, candidates, genome_size_expected, genome_size_sigma, mode="normal"):
        """MultiDbg::to_score (posterior.rs:259-277) for every candidate copy-number vector over compact edges at once: what
        sample_posterior_once evaluates per neighbour with dbg.clone() + set_copy_nums + to_phmm + to_full_prob_reads
        (posterior.rs:504-515).  `phmm`: a model of this graph (to_phmm); its parameter sets are replaced by the candidates'.
        One expansion to k-mer copy numbers, one on-device derivation of (init, trans) for the batch, one batched
        to_full_prob_reads (use_max_ratio = true like to_likelihood, posterior.rs:247-255), then the host-side terms."""
        x, b = self._candidates(candidates)
        phmm.set_copy_nums_batch(self.expand_copy_nums(x), mode)
        like = phmm.to_full_prob_reads(reads, mappings, True)[0]
        gs = self.genome_size(x); prior = self.to_prior(genome_size_expected, genome_size_sigma, x); ne = self.n_euler_circuits(x)
        return [Score(float(like[i]), float(prior[i]), int(gs[i]), float(ne[i])) for i in range(b)]

    def clone(self):
        return MultiDbg.from_dbg_str(self.to_dbg_string())

    def sample_posterior_once(self, phmm, reads, mappings, neighbors, posterior, genome_size_expected, genome_size_sigma, mode="normal",
                              multi_move=False):
        """MultiDbg::sample_posterior_once (posterior.rs:470-600): score every neighbour the posterior has not seen yet -- all of them
        in ONE batched to_scores where the reference clones the graph per neighbour under rayon -- add them, and return the best
        sample (copy_nums tuple, Score) if it is not the current copy-number vector, else None.
        `neighbors`: copy-number vectors over compact edges, e.g. the output of the reference's neighbour search (not built here).
        multi_move (posterior.rs:533-588, the mode of rescue-only rounds): the neighbours that improve on the current score are
        taken best first and accepted while their changes touch disjoint edges (`is_independent_update`, neighbors.rs:493-509; the
        update cycle of a neighbour is its difference from the current vector); the combined move is scored and added as well."""
        def score_new(cands):
            todo, seen = [], set()
            for c in cands:
                key = tuple(int(v) for v in c)
                if key not in seen and not posterior.contains(key):
                    seen.add(key); todo.append(key)
            if todo:
                for key, sc in zip(todo, self.to_scores(phmm, reads, mappings, np.array(todo, np.uint32), genome_size_expected, genome_size_sigma, mode)):
                    posterior.add(key, sc)
        score_new(neighbors)
        current = tuple(int(v) for v in self.get_copy_nums())
        if multi_move:
            cur_score = posterior.find(current)
            if cur_score is None:
                raise DbgphmmError(ERR_INVALID, "current copy number was not sampled")       # posterior.rs:540 (expect)
            keys = [tuple(int(v) for v in c) for c in neighbors]
            # sorted_by_key(score).rev(): descending, later neighbours first among equal scores
            order = sorted(range(len(keys)), key=lambda i: (posterior.find(keys[i]).p(), i), reverse=True)
            combined, touched = list(current), set()
            for i in order:
                if not posterior.find(keys[i]).p() > cur_score.p():
                    break                                                                     # neighbours are sorted by score
                delta = {e: keys[i][e] - current[e] for e in range(len(current)) if keys[i][e] != current[e]}
                if touched.isdisjoint(delta):
                    for e, dv in delta.items():
                        combined[e] += dv
                    touched.update(delta)
            score_new([combined])
        best = posterior.max_sample()
        return best if best[0] != current else None

    def sample_posterior(self, phmm, reads, mappings, genome_size_expected, genome_size_sigma, neighbors_fn, max_iter, mode="normal",
                         multi_move=False):
        """The greedy search of MultiDbg::sample_posterior (posterior.rs:314-420): start from the current copy numbers, score the
        neighbours, move to the best one, stop at a local optimum or after max_iter moves.  `neighbors_fn(dbg)` returns the candidate
        sets to try in order (the reference tries rescue, partial and full neighbours, posterior.rs:352-372) for the copy numbers `dbg`
        currently holds; multi_move as in sample_posterior_once (the reference switches it on for rescue-only rounds, posterior.rs:391).
        `self` is left untouched; returns the Posterior."""
        post = Posterior()
        dbg = self.clone()
        copy_nums = tuple(int(v) for v in dbg.get_copy_nums())
        post.add(copy_nums, dbg.to_scores(phmm, reads, mappings, np.array([copy_nums], np.uint32), genome_size_expected, genome_size_sigma, mode)[0])
        n_iter = 0
        while n_iter < max_iter:
            dbg.set_copy_nums(np.array(copy_nums, np.uint32))
            for cand in neighbors_fn(dbg):
                sample = dbg.sample_posterior_once(phmm, reads, mappings, cand, post, genome_size_expected, genome_size_sigma, mode, multi_move)
                if sample is not None:
                    copy_nums = sample[0]
                    n_iter += 1
                    break
            else:
                break        # no better neighbour: local optimum
        return post

    def _to_phmm(self, param, mode, device, mem_budget_bytes):
        src, dst, em, cn, _ = self.phmm_graph()
        h = C.c_void_p()
        _check(lib().dbgphmm_dbg_to_model(self._h, C.byref(param), mode, device, mem_budget_bytes, C.byref(h)))
        m = PHMMModel.__new__(PHMMModel)
        m.src, m.dst, m.emission = src, dst, em
        m.param = param.copy(); m.param.n_warmup = self._k
        m.n_nodes = self.n_edges_full; m.device = device; m._h = h
        return m

    def to_phmm(self, param, device=0, mem_budget_bytes=0):
        """MultiDbg::to_phmm (multi_dbg.rs:1394-1397)."""
        return self._to_phmm(param, 0, device, mem_budget_bytes)

    def to_non_zero_phmm(self, param, device=0, mem_budget_bytes=0):
        """MultiDbg::to_non_zero_phmm (multi_dbg.rs:1406-1409)."""
        return self._to_phmm(param, 1, device, mem_budget_bytes)

    def to_uniform_phmm(self, param, device=0, mem_budget_bytes=0):
        """MultiDbg::to_uniform_phmm (multi_dbg.rs:1400-1403)."""
        return self._to_phmm(param, 2, device, mem_budget_bytes)


class PHMMModel:
    """PHMMModel<PNode,PEdge> (hmmv2/common.rs:61-67) resident on one B200.

    edge_src/edge_dst in EdgeIndex order, emission bytes ('A','C','G','T','n'), log_init/log_trans as natural logs."""

    def __init__(self, edge_src, edge_dst, emission, log_init, log_trans, param, device=0, mem_budget_bytes=0):
        self.src = np.ascontiguousarray(edge_src, np.uint32)
        self.dst = np.ascontiguousarray(edge_dst, np.uint32)
        self.emission = _bases(emission)
        li = np.ascontiguousarray(log_init, np.float64); lt = np.ascontiguousarray(log_trans, np.float64)
        self.param = param.copy()
        self.n_nodes = len(self.emission)
        self.device = device
        h = C.c_void_p()
        _check(lib().dbgphmm_model_create(self.n_nodes, len(self.src), _p(self.src), _p(self.dst), _p(self.emission), _p(li), _p(lt),
                                          C.byref(self.param), device, mem_budget_bytes, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dbgphmm_model_destroy(self._h)
            self._h = None

    def close(self):
        """Release the device graph and every cached row buffer of this device now (also happens when the object is collected)."""
        self.__del__()

    def n_edges(self):
        return len(self.src)

    def set_params(self, param):
        self.param = param.copy()
        _check(lib().dbgphmm_model_set_params(self._h, C.byref(self.param)))

    def set_probs(self, log_init, log_trans):
        li = np.ascontiguousarray(log_init, np.float64); lt = np.ascontiguousarray(log_trans, np.float64)
        _check(lib().dbgphmm_model_set_probs(self._h, _p(li), _p(lt)))

    def set_copy_nums_batch(self, copy_nums, mode="normal"):
        """Candidate copy-number assignments X [B][N] -> B parameter sets on the device (seq_graph.rs:160-273)."""
        cn = np.ascontiguousarray(copy_nums, np.uint32)
        if cn.ndim == 1:
            cn = cn[None, :]
        assert cn.shape[1] == self.n_nodes
        _check(lib().dbgphmm_model_set_copy_nums_batch(self._h, cn.shape[0], _p(cn), {"normal": 0, "non_zero": 1, "uniform": 2}[mode]))

    def get_probs(self, x=0):
        li = np.empty(self.n_nodes); lt = np.empty(len(self.src))
        with np.errstate(divide="ignore"):
            _check(lib().dbgphmm_model_get_probs(self._h, x, _p(li), _p(lt)))
        return li, lt

    def n_batch(self):
        return int(lib().dbgphmm_model_n_batch(self._h))

    def wave_reads(self):
        """Reads whose sparse rows are resident at once (SMs x sparse jobs per SM): the efficient batch quantum of the bulk calls."""
        return int(lib().dbgphmm_model_wave_reads(self._h))

    # ---- forward.rs
    def _fwd(self, x, kind, mappings=None, read_index=0):
        x = _bases(x)
        h = C.c_void_p()
        _check(lib().dbgphmm_forward(self._h, _p(x), len(x), kind, mappings._h if mappings is not None else None, read_index, C.byref(h)))
        return PHMMTables(self, h)

    def forward(self, x):
        return self._fwd(x, FWD_DENSE)

    def forward_sparse(self, x, use_max_ratio):
        return self._fwd(x, FWD_SPARSE_RATIO if use_max_ratio else FWD_SPARSE)

    def forward_with_mapping(self, x, mappings, read_index=0):
        return self._fwd(x, FWD_MAPPING, mappings, read_index)

    # ---- backward.rs
    def _bwd(self, x, kind, mappings=None, read_index=0, fwd=None):
        x = _bases(x)
        h = C.c_void_p()
        _check(lib().dbgphmm_backward(self._h, _p(x), len(x), kind, mappings._h if mappings is not None else None, read_index,
                                      fwd._h if fwd is not None else None, C.byref(h)))
        return PHMMTables(self, h)

    def backward(self, x):
        return self._bwd(x, BWD_DENSE)

    def backward_sparse(self, x):
        return self._bwd(x, BWD_SPARSE)

    def backward_with_mapping(self, x, mappings, read_index=0):
        return self._bwd(x, BWD_MAPPING, mappings, read_index)

    def backward_by_forward(self, x, forward):
        return self._bwd(x, BWD_BY_FORWARD, fwd=forward)

    # ---- freq.rs:42-76
    def run(self, x):
        return PHMMOutput(self, self.forward(x), self.backward(x))

    def run_sparse(self, x):
        return PHMMOutput(self, self.forward_sparse(x, False), self.backward_sparse(x))

    def run_sparse_adaptive(self, x, use_max_ratio):
        f = self.forward_sparse(x, use_max_ratio)
        return PHMMOutput(self, f, self.backward_by_forward(x, f))

    def run_with_mapping(self, x, mappings, read_index=0):
        return PHMMOutput(self, self.forward_with_mapping(x, mappings, read_index), self.backward_with_mapping(x, mappings, read_index))

    def q_score_exact(self, edge_freqs, init_freqs, x=0):
        """q.rs:66-96 -> (init, trans, prior); QScore::total() is their sum."""
        ef = np.ascontiguousarray(edge_freqs, np.float64); nf = np.ascontiguousarray(init_freqs, np.float64)
        out = np.zeros(3)
        _check(lib().dbgphmm_q_score_exact(self._h, x, _p(ef), _p(nf), _p(out)))
        return tuple(out)

    # ---- bulk calls (freq.rs:87-192, hint.rs:193-220)
    def to_full_prob_reads(self, reads, mappings=None, use_max_ratio=True):
        """-> (ln P(R|X) [n_batch], per-read ln P [n_batch][n_reads]) — freq.rs:175-192 for every candidate X."""
        B = self.n_batch()
        tot = np.empty(B); per = np.empty((B, len(reads)))
        _check(lib().dbgphmm_to_full_prob_reads(self._h, reads._h, mappings._h if mappings is not None else None, int(use_max_ratio),
                                                _p(tot), _p(per)))
        return tot, per

    def run_node_freqs(self, reads, mode, use_max_ratio=True, mappings=None, want_freqs=True):
        """-> (node_freqs[N] summed over reads, ln P forward [R], ln P backward [R], (cells_fwd, cells_bwd))."""
        fr = np.zeros(self.n_nodes) if want_freqs else None
        lf = np.empty(len(reads)); lb = np.empty(len(reads)); cells = np.zeros(2, np.uint64)
        _check(lib().dbgphmm_run_node_freqs(self._h, reads._h, RUN_MODES[mode], int(use_max_ratio),
                                            mappings._h if mappings is not None else None, _p(fr), _p(lf), _p(lb), _p(cells)))
        return fr, lf, lb, (int(cells[0]), int(cells[1]))

    # ---- the remaining whole-read-set methods of freq.rs, each one bulk call (reads: Reads or a list of sequences)
    @staticmethod
    def _reads(seqs):
        return seqs if isinstance(seqs, Reads) else Reads(list(seqs))

    def to_node_freqs(self, seqs):
        """PHMMModel::to_node_freqs (freq.rs:87-102): dense forward + backward per read, node frequencies summed over reads."""
        return self.run_node_freqs(self._reads(seqs), "dense")[0]

    def to_full_prob(self, seqs):
        """to_full_prob / to_full_prob_parallel (freq.rs:105-135): ln of the product over reads of the dense forward P(read);
        the sum is taken in read order (rayon's order is unspecified)."""
        return float(self.run_node_freqs(self._reads(seqs), "dense", want_freqs=False)[1].sum())

    to_full_prob_parallel = to_full_prob

    def to_full_prob_sparse(self, seqs, use_max_ratio):
        """freq.rs:138-150: product of forward_sparse(read, use_max_ratio).full_prob() (parameter set 0)."""
        return float(self.to_full_prob_reads(self._reads(seqs), None, use_max_ratio)[0][0])

    def to_full_prob_sparse_backward(self, seqs):
        """freq.rs:153-164: product of backward_sparse(read).full_prob()."""
        return float(self.run_node_freqs(self._reads(seqs), "sparse", want_freqs=False)[2].sum())

    def forward_sparse_score_only(self, x, use_max_ratio):
        """forward.rs:158-206: forward_sparse keeping one row; returns ln P (parameter set 0)."""
        return float(self.to_full_prob_reads(Reads([x]), None, use_max_ratio)[1][0, 0])

    def forward_with_mapping_score_only(self, x, mappings, read_index=0):
        """forward.rs:79-89 with the mapping of read `read_index` of `mappings`."""
        one = Mappings.from_list([mappings[read_index]])
        return float(self.to_full_prob_reads(Reads([x]), one, False)[1][0, 0])

    def run_node_freqs_dev(self, reads, mode, node_freqs_ptr, use_max_ratio=True, mappings=None, logp_fwd_ptr=None, logp_bwd_ptr=None):
        """Device-pointer variant: node_freqs_ptr (f64[N] on this model's device) is accumulated into."""
        cells = np.zeros(2, np.uint64)
        _check(lib().dbgphmm_run_node_freqs_dev(self._h, reads._h, RUN_MODES[mode], int(use_max_ratio),
                                                mappings._h if mappings is not None else None, C.c_void_p(node_freqs_ptr),
                                                C.c_void_p(logp_fwd_ptr) if logp_fwd_ptr else None,
                                                C.c_void_p(logp_bwd_ptr) if logp_bwd_ptr else None, _p(cells)))
        return int(cells[0]), int(cells[1])

    def generate_mappings(self, reads, mappings=None, use_max_ratio=True):
        h = C.c_void_p()
        _check(lib().dbgphmm_generate_mappings(self._h, reads._h, mappings._h if mappings is not None else None, int(use_max_ratio), C.byref(h)))
        return Mappings._from_handle(h)

    def reads_to_device(self, reads):
        _check(lib().dbgphmm_reads_to_device(self._h, reads._h))
