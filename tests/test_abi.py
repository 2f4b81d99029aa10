"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every declared symbol, and refuses to
compute without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dbgphmm_b200 import build as B
from dbgphmm_b200 import graphs
from dbgphmm_b200 import hmmv2 as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return H.lib()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "dbgphmm_b200.h")).read()
    declared = set(re.findall(r"\b(dbgphmm_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(H.SYMBOLS), declared ^ set(H.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s)


def test_params_match_reference_formula(lib):
    q = H.params_uniform(0.001)  # params.rs:116-124
    assert q.n_active_nodes == 40 and q.n_warmup == 50 and q.warmup_threshold == 200 and q.n_max_gaps == 4
    assert abs(np.exp(q.p_MM) - (1 - 0.002 - 1e-5)) < 1e-15 and q.p_random == np.log(0.25)
    from oracle import oracle as O
    o = O.params_uniform(0.001)
    for name, _ in H.Params._fields_:
        assert getattr(q, name) == getattr(o, name), name


def test_reads_validation(lib):
    with pytest.raises(H.DbgphmmError):
        H.Reads([b"ACGN"])  # collection.rs:236-249 panics on non-ACGT
    r = H.Reads([b"ACGT", b"GG"])
    assert len(r) == 2 and r.total_bases() == 6


def test_mappings_roundtrip_and_freqs(lib):
    m = H.Mappings.from_list([H.Mapping([[1, 2], [3]], [[np.log(0.75), np.log(0.25)], [0.0]])])
    assert m.n_reads() == 1 and len(m[0]) == 2
    f = m.to_node_freqs(5)  # hint.rs:161-171
    assert np.allclose(f, [0, 0.75, 0.25, 1.0, 0])


def test_mapping_node_convert(lib):
    """hint.rs:234-270 (mapping_node_convert): case 1 is the reference's exact assertion; case 2 (printed only there) and the
    dropped-node case of PurgeEdgeMap::update_mapping (multi_dbg.rs:1783-1791) follow from map_nodes' definition (hint.rs:66-88)."""
    lg = np.log
    m = H.Mappings.from_list([H.Mapping([[0, 1], [2, 3]], [[lg(0.6), lg(0.4)], [lg(0.9), lg(0.1)]])])
    m1 = m.map_nodes(lambda v: [v + 1], 4)
    assert [list(x) for x in m1[0].nodes] == [[1, 2], [3, 4]]
    assert [list(x) for x in m1[0].probs] == [[lg(0.6), lg(0.4)], [lg(0.9), lg(0.1)]]   # p / 1 is exact: assert_eq in the reference
    m2 = m.map_nodes(lambda v: [v, v + 1], 4)
    assert [list(x) for x in m2[0].nodes] == [[1, 0, 2], [3, 2, 4]]
    assert np.allclose(np.exp(np.concatenate(m2[0].probs)), [0.5, 0.3, 0.2, 0.5, 0.45, 0.05], rtol=1e-14)
    m3 = m.map_nodes(lambda v: [] if v in (0, 3) else [7], 4)
    assert [list(x) for x in m3[0].nodes] == [[7], [7]]
    assert np.allclose(np.exp(np.concatenate(m3[0].probs)), [0.4, 0.9], rtol=1e-14)
    # several reads keep their row structure; more images than MAX_ACTIVE_NODES are cut to the 400 most probable
    two = H.Mappings.from_list([H.Mapping([[0]], [[0.0]]), H.Mapping([[1], [0, 1]], [[0.0], [lg(0.5), lg(0.5)]])])
    wide = two.map_nodes(lambda v: list(range(100 + 500 * v, 600 + 500 * v)), 2)
    assert wide.n_reads() == 2 and len(wide[0]) == 1 and len(wide[1]) == 2
    assert [len(x) for x in wide[1].nodes] == [400, 400] and list(wide[1].nodes[0][:3]) == [600, 601, 602]
    with pytest.raises(H.DbgphmmError):
        m.map_nodes(lambda v: [v], 2)   # node 2 / 3 outside the node map


def test_no_gpu_means_loud_failure(lib):
    if H.device_count() > 0:
        pytest.skip("a GPU is present")
    sg = graphs.mock_linear()
    li, lt = sg.to_probs()
    with pytest.raises(H.DbgphmmError) as ei:
        H.PHMMModel(sg.src, sg.dst, sg.base, li, lt, H.params_uniform(0.01))
    assert ei.value.status == H.ERR_CUDA


def test_rust_shim_stays_in_step_with_the_header_and_the_build():
    """rust/ is source only (no cargo/rustc here): check what can be checked without compiling it -- its `extern "C"`
    block binds exactly the functions include/dbgphmm_b200.h declares, each with the same number of arguments, the params struct has the
    header's fields in the header's order, and build.rs compiles the same sources with the same link libraries as
    dbgphmm_b200/build.py."""
    hdr = open(os.path.join(ROOT, "include", "dbgphmm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    rs = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    rs_nc = re.sub(r"//[^\n]*", "", rs)

    def n_args(arglist):
        a = arglist.strip()
        return 0 if a in ("", "void") else a.count(",") + 1

    c_decl = {m.group(1): n_args(m.group(2)) for m in re.finditer(r"\b(dbgphmm_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", hdr)}
    ext = re.search(r'extern "C" \{(.*?)\n\}', rs_nc, flags=re.S).group(1)
    r_decl = {m.group(1): n_args(m.group(2)) for m in re.finditer(r"pub fn (dbgphmm_[a-z_0-9]+)\s*\(([^)]*)\)", ext)}
    assert set(r_decl) == set(c_decl) == set(H.SYMBOLS), set(r_decl) ^ set(c_decl)
    for name, n in r_decl.items():
        assert n == c_decl[name], f"{name}: {n} arguments in the Rust binding, {c_decl[name]} in the header"

    c_struct = re.search(r"typedef struct dbgphmm_params\s*\{(.*?)\}", hdr, flags=re.S)
    if c_struct is None:
        c_struct = re.search(r"struct dbgphmm_params\s*\{(.*?)\}", hdr, flags=re.S)
    c_fields = []
    for stmt in c_struct.group(1).split(";"):
        toks = stmt.replace(",", " ").split()
        if len(toks) >= 2:
            c_fields += [t.lower() for t in toks[1:]]
    r_struct = re.search(r"pub struct dbgphmm_params \{(.*?)\n\}", rs_nc, flags=re.S).group(1)
    r_fields = re.findall(r"pub ([a-z_0-9]+):", r_struct)
    assert r_fields == c_fields, (r_fields, c_fields)
    assert [n.lower() for n, _ in H.Params._fields_] == c_fields

    brs = open(os.path.join(ROOT, "rust", "build.rs")).read()
    srcs = re.search(r"let srcs = \[(.*?)\];", brs).group(1)
    assert re.findall(r'"([a-z_]+\.cu)"', srcs) == B.SOURCES
    for flag in ("arch=compute_100a,code=sm_100a", "-lineinfo", "-lcudart", "-lz"):
        assert flag in brs, flag


def test_mappings_create_rejects_malformed_csr(lib):
    ro = lambda *v: np.array(v, np.uint64)
    nodes = np.array([1, 2, 3], np.uint32); lp = np.zeros(3)
    ok = H.Mappings(ro(0, 2), ro(0, 2, 3), nodes, lp)
    assert ok.n_reads() == 1 and [list(x) for x in ok[0].nodes] == [[1, 2], [3]]
    for read_off, row_off in ((ro(1, 2), ro(0, 2, 3)),        # read_off[0] != 0
                              (ro(0, 2, 1), ro(0, 2, 3)),     # read_off decreasing
                              (ro(0, 2), ro(1, 2, 3)),        # row_off[0] != 0
                              (ro(0, 2), ro(0, 3, 2))):       # row_off decreasing
        with pytest.raises(H.DbgphmmError) as ei:
            H.Mappings(read_off, row_off, nodes, lp)
        assert ei.value.status == H.ERR_INVALID
    h = C.c_void_p()
    st = lib.dbgphmm_mappings_create(1, ro(0, 2).ctypes.data_as(C.c_void_p), ro(0, 2, 3).ctypes.data_as(C.c_void_p), None, None, C.byref(h))
    assert st == H.ERR_INVALID and b"nodes is null" in lib.dbgphmm_last_error()


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header compiles as pedantic C99 and a C program links against the library (what cgo / JNI /
    a Rust `extern "C"` block would bind)."""
    import subprocess
    B.build()
    src = tmp_path / "c.c"
    src.write_text('#include <stdio.h>\n#include "dbgphmm_b200.h"\n'
                   'int main(void) { dbgphmm_params p; dbgphmm_params_uniform(0.01, &p); double lp = 1.0;\n'
                   '  if (dbgphmm_prior_normal(30.0, 28.0, 5.0, &lp) != DBGPHMM_OK || !(lp < 0.0)) return 1;\n'
                   '  printf("%u %d\\n", p.n_active_nodes, dbgphmm_device_count() >= 0); return p.n_max_gaps == 4 ? 0 : 2; }\n')
    libdir = os.path.join(ROOT, "dbgphmm_b200", "lib")
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    exe = tmp_path / "c"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-L", libdir, "-ldbgphmm_b200", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-lcudart", "-o", str(exe)], env=env)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.split()[0] == "40", (out.returncode, out.stdout, out.stderr)
