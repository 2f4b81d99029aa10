"""DBG / MAP text formats (multi_dbg/output.rs:155-345, 455-623): host-only entry points of the C ABI, no GPU needed.

Pinned on the reference's own fixtures: the DBG example of README.md:174-191 is toy::repeat (multi_dbg/toy.rs:260-305), whose
node-centric graph tests/ already hold as graphs.toy_repeat(); the dump/load round trip mirrors output.rs:831-857 (dumpload,
dbg_gz_compressed) and the MAP round trip output.rs:880-905 (map)."""
import gzip
import json
import os

import numpy as np
import pytest

from dbgphmm_b200 import graphs, hmmv2 as H

README_DBG = """# #: comment
# K section: k of DBG
K\t4
# N section: node = (k-1)-mer
#\tid\tsequence of k-1-mer
N\t0\tnnn
N\t1\tCAG
# E section: edge = simple path of k-mers
E\t0\t1\t0\tCAGGAAnnn\t1\t9,10,11,12,13,14
E\t1\t1\t1\tCAGCAG\t3\t6,7,8
E\t2\t0\t1\tnnnTCCCAG\t1\t0,1,2,3,4,5
"""


def test_readme_dbg_is_toy_repeat():
    d = H.MultiDbg.from_dbg_str(README_DBG)
    assert (d.k(), d.n_nodes_compact, d.n_edges_compact, d.n_edges_full) == (4, 2, 3, 15)
    src, dst, em, cn, ce = d.phmm_graph()
    sg, k = graphs.toy_repeat()
    assert k == 4
    assert bytes(em) == bytes(sg.base) and list(cn) == list(sg.node_copy_num)
    # same PHMM edge set; the node-centric edges are grouped by the (k-1)-mer they pass through, whose numbering differs
    # between the hand-written toy (14 named nodes) and from_dbg_reader (compact nodes first), so compare per target node
    assert sorted(zip(src.tolist(), dst.tolist())) == sorted(zip(sg.src.tolist(), sg.dst.tolist()))
    for v in range(sg.n_nodes):   # parent order of every node (it fixes the fold order of the parent sums)
        assert [s for s, t in zip(src, dst) if t == v] == [s for s, t in zip(sg.src, sg.dst) if t == v]
    assert list(ce) == [2] * 6 + [1] * 3 + [0] * 6
    assert list(d.get_copy_nums()) == [1, 3, 1]


def test_dbg_dump_load_round_trip(tmp_path):
    d = H.MultiDbg.from_dbg_str(README_DBG)
    s = d.to_dbg_string()
    d1 = H.MultiDbg.from_dbg_str(s)
    assert d1.to_dbg_string() == s
    body = [ln for ln in s.splitlines() if not ln.startswith("#")]
    assert body == [ln for ln in README_DBG.splitlines() if ln and not ln.startswith("#")]
    for name in ("hoge.dbg", "repeat.dbg.gz", "repeat.dbz"):
        path = tmp_path / name
        d.to_dbg_file(path)
        raw = open(path, "rb").read()
        assert (raw[:2] == b"\x1f\x8b") == (not name.endswith(".dbg"))
        d2 = H.MultiDbg.from_dbg_file(path)
        assert d2.to_dbg_string() == s
        for a, b in zip(d.phmm_graph(), d2.phmm_graph()):
            assert np.array_equal(a, b)


def test_copy_numbers_over_compact_edges():
    d = H.MultiDbg.from_dbg_str(README_DBG)
    X = np.array([[1, 3, 1], [1, 0, 1], [2, 5, 2]], np.uint32)
    full = d.expand_copy_nums(X)
    assert full.shape == (3, 15)
    _, _, _, _, ce = d.phmm_graph()
    assert np.array_equal(full, X[:, ce])
    d.set_copy_nums([1, 7, 1])
    assert list(d.get_copy_nums()) == [1, 7, 1] and "CAGCAG\t7\t" in d.to_dbg_string()
    with pytest.raises(H.DbgphmmError):   # flow in != flow out at a node (multi_dbg.rs:1051 asserts)
        d.set_copy_nums([1, 3, 2])


@pytest.mark.parametrize("bad", ["N\t0\tnnn\n", "K\t4\nN\t1\tnnn\n", "K\t4\nN\t0\tnnn\nE\t0\t0\t0\tnnnA\t1\t0,1\n",
                                 "K\t4\nN\t0\tnnn\nE\t1\t0\t0\tnnnA\t1\t0\n", "K\t4\nN\t0\tnnn\nE\t0\t0\t0\tnnnAC\t1\t0,0\n"])
def test_dbg_reader_rejects_what_the_reference_asserts(bad):
    with pytest.raises(H.DbgphmmError):
        H.MultiDbg.from_dbg_str(bad)


def test_map_round_trip(tmp_path):
    reads = H.Reads([b"GATCC", b"TAT"])
    vals = [0.0, -1e-7, -0.1, -2.302585092994046, -13.815510557964274, -1234.5678901234567, -np.inf, -5e-324]
    rows = [[(3, vals[0]), (2, vals[1]), (4, vals[2])], [(4, vals[3])], [(5, vals[4]), (6, vals[5])], [(6, vals[6])], [(7, vals[7]), (8, -1.0)],
            [(0, -0.5)], [(1, -0.25), (2, -3.0)], [(9, -0.125)]]
    mp = H.Mappings.from_list([H.Mapping([[n for n, _ in r] for r in rows[:5]], [[p for _, p in r] for r in rows[:5]]),
                               H.Mapping([[n for n, _ in r] for r in rows[5:]], [[p for _, p in r] for r in rows[5:]])])
    d = H.MultiDbg.from_dbg_str(README_DBG)
    s = mp.to_map_string(reads, d)
    lines = s.splitlines()
    assert "# k=4 n_edges_full=15 n_edges_compact=3" in lines and "# read\tpos\tbase\tnodes_and_probs" in lines and "# i=1" in lines
    body = [ln for ln in lines if not ln.startswith("#")]
    # Rust's `{}` for f64: shortest round-trip digits, never an exponent (output.rs:516)
    assert body[0] == "0\t0\tG\t3:0,2:-0.0000001,4:-0.1"
    assert body[1] == "0\t1\tA\t4:-2.302585092994046"
    assert body[3] == "0\t3\tC\t6:-inf"
    assert body[5] == "1\t0\tT\t0:-0.5"
    assert "e" not in body[4].split("\t")[3]
    mp2 = H.Mappings.from_map_str(s)
    assert np.array_equal(mp2.read_off, mp.read_off) and np.array_equal(mp2.row_off, mp.row_off)
    assert np.array_equal(mp2.nodes, mp.nodes) and np.array_equal(mp2.probs, mp.probs)   # bit-exact round trip
    for name in ("a.map", "a.map.gz", "a.mpz"):
        path = tmp_path / name
        mp.to_map_file(path, reads, d)
        if name != "a.map":
            assert gzip.open(path, "rt").read() == s
        mp3 = H.Mappings.from_map_file(path)
        assert np.array_equal(mp3.nodes, mp.nodes) and np.array_equal(mp3.probs, mp.probs)
    with pytest.raises(H.DbgphmmError):
        H.Mappings.from_map_str("0\t1\tA\t3:-0.5\n")   # first row of a read must be position 0 (output.rs:551-552)
    with pytest.raises(H.DbgphmmError):
        mp.to_map_string(H.Reads([b"GATCC"]), d)


@pytest.mark.gpu
def test_dbg_file_to_phmm_scores_like_the_hand_built_model():
    from oracle import oracle as O
    d = H.MultiDbg.from_dbg_str(README_DBG)
    sg, k = graphs.toy_repeat()
    par = H.params_uniform(0.01)
    g = d.to_non_zero_phmm(par)
    li, lt = sg.to_probs("non_zero")
    op = O.params_uniform(0.01); op.n_warmup = k
    o = O.PHMMModel(sg.src, sg.dst, sg.base, li, lt, op)
    reads = [b"TCCCAGCAGCAGCAGGAA", b"CCAGCAGG"]
    tot, per = g.to_full_prob_reads(H.Reads(reads), None, False)
    s, p = o.to_full_prob_reads(O.Reads(reads), None, False)
    assert np.allclose(per[0], p, rtol=1e-9, atol=0)
    # a batch of candidates over compact edges
    X = np.array([[1, 3, 1], [1, 2, 1], [1, 6, 1]], np.uint32)
    g.set_copy_nums_batch(d.expand_copy_nums(X), "normal")
    tot, per = g.to_full_prob_reads(H.Reads(reads), None, False)
    for b in range(3):
        li, lt = sg.to_probs("normal", d.expand_copy_nums(X[b])[0])
        o.set_probs(li, lt)
        s, p = o.to_full_prob_reads(O.Reads(reads), None, False)
        assert np.allclose(per[b], p, rtol=1e-9, atol=0)


def test_no_cpp_exception_crosses_the_c_abi():
    """Every status-returning entry point is a function-try-block (ABI_CATCH, csrc/model.h).  A second K line that contradicts
    the E lines already read makes std::string::substr throw inside dbg_finish: the caller gets a status, not std::terminate."""
    with pytest.raises(H.DbgphmmError) as ei:
        H.MultiDbg.from_dbg_str("K\t4\nN\t0\tnnn\nE\t0\t0\t0\tnnnA\t1\t0\nK\t10\n")
    assert ei.value.status == H.ERR_INVALID and "C++ exception" in str(ei.value)


def test_mutated_dbg_and_map_texts_are_parsed_or_rejected_never_crash():
    """1500 seeded random mutations of the README DBG text and of a MAP text: every one either parses (and can then be walked,
    written back, scored for Euler circuits) or is rejected with a status.  (12,000 mutations were run once without a crash.)"""
    import random
    rnd = random.Random(1)
    map_text = "# c\n0\t0\tA\t1:-0.5,2:-1.25\n0\t1\tC\t3:-inf\n1\t0\tG\t\n1\t1\tT\t7:0\n"
    parsed = rejected = 0
    for _ in range(1500):
        base = rnd.choice([README_DBG, map_text])
        b = bytearray(base.encode())
        for _ in range(rnd.randint(1, 6)):
            op, pos = rnd.random(), rnd.randrange(len(b))
            if op < 0.4:
                b[pos] = rnd.choice(b"0123456789,\t\n:-KNEnACGT. ")
            elif op < 0.7:
                del b[pos:pos + rnd.randint(1, 5)]
            else:
                b[pos:pos] = bytes(rnd.choice(b"0123456789,\t\n:-KNE") for _ in range(rnd.randint(1, 4)))
            if not b:
                b = bytearray(b"K")
        text = bytes(b).decode("latin1")
        try:
            if base is map_text:
                m = H.Mappings.from_map_str(text)
                [m[r] for r in range(m.n_reads())]
            else:
                d = H.MultiDbg.from_dbg_str(text)
                d.phmm_graph(); d.genome_size()
                assert H.MultiDbg.from_dbg_str(d.to_dbg_string()).to_dbg_string() == d.to_dbg_string()
                try:
                    d.n_euler_circuits()
                except H.DbgphmmError:
                    pass                       # copy numbers that no longer balance
            parsed += 1
        except H.DbgphmmError:
            rejected += 1
    assert parsed > 50 and rejected > 50


# ---- dataset JSON (Dataset::to_json_file / from_json_file, e2e.rs:123-130)
def _ref_prob(lnp):
    """Prob's Display (prob.rs:158-162): '{}({:.4})' of (ln p, p) the way Rust prints f64."""
    import math
    if lnp == -math.inf:
        return "-inf(0.0000)"
    r = repr(float(lnp))
    if "e" in r or "E" in r:
        from decimal import Decimal
        r = format(Decimal(r), "f")
    if r.endswith(".0"):
        r = r[:-2]
    return f"{r}({math.exp(lnp):.4f})"


def test_dataset_json_matches_the_reference_serde_form(tmp_path):
    from dbgphmm_b200 import hmmv2 as H
    par = H.params_uniform(0.001)
    haps = [b"ATCGATTTAGC", b"GGGC"]
    reads = [b"ATCGT", b"TTAG"]
    origins = [[(0, 0), (0, 1), (0, 2), None, (0, 3)], [(0, 5), (0, 6), (0, 7), (0, 9)]]   # the fixture of collection.rs:847-861 + one with a deletion
    d = H.Dataset.new(haps, ["L", "C"], reads, par, revcomp=[1, 0], origins=origins)
    text = d.to_json_string()
    doc = json.loads(text)
    # the literal forms the reference's own tests pin: styled sequences (collection.rs:836-842), positioned reads (:711-725)
    assert doc["genome"] == ["L:ATCGATTTAGC", "C:GGGC"]
    assert doc["genome_size"] == 15
    assert doc["reads"] == {"reads": ["ATCGT:-:0-0,0-1,0-2,I,0-3", "TTAG:+:0-5,0-6,0-7,0-9"]}
    pp = doc["phmm_params"]
    assert list(pp) == ["p_mismatch", "p_match", "p_random", "p_gap_open", "p_gap_ext", "p_end", "p_MM", "p_IM", "p_DM", "p_MI", "p_II", "p_DI", "p_MD", "p_ID",
                        "p_DD", "n_active_nodes", "active_node_max_ratio", "n_warmup", "warmup_threshold", "n_max_gaps"]      # field order of params.rs:16-66
    for name in list(pp)[:15]:
        assert pp[name] == _ref_prob(getattr(par, name)), name
    assert (pp["n_active_nodes"], pp["active_node_max_ratio"], pp["n_warmup"], pp["warmup_threshold"], pp["n_max_gaps"]) == (40, 30.0, 50, 200, 4)
    assert '"active_node_max_ratio":30.0' in text       # serde_json prints an f64 with its fraction
    # round trip through text and through a gzip file
    for e in (H.Dataset.from_json_str(text), None):
        if e is None:
            p = str(tmp_path / "d.json.gz"); d.to_json_file(p); e = H.Dataset.from_json_file(p)
        assert e.to_json_string() == text
        assert e.genome() == [("L", haps[0]), ("C", haps[1])] and e.genome_size() == 15
        assert [bytes(r) for r in [e.reads()[0], e.reads()[1]]] == reads
        assert e.read_origins() == ([True, False], origins)
        q = e.params()
        assert all(getattr(q, n) == getattr(par, n) for n, _ in H.Params._fields_)
        assert abs(e.coverage() - 9 / 15) < 1e-15


def test_dataset_json_reader_accepts_reference_output_and_rejects_malformed_text():
    from dbgphmm_b200 import hmmv2 as H
    # as serde_json writes it: no spaces, zero probability as "-inf(0.0000)", warmup_threshold absent in old files (params.rs:60 default)
    probs = ",".join(f'"{n}":"{v}"' for n, v in [("p_mismatch", "-inf(0.0000)"), ("p_match", "0(1.0000)"), ("p_random", "-1.3862943611198906(0.2500)"),
                                                   ("p_gap_open", "-inf(0.0000)"), ("p_gap_ext", "-inf(0.0000)"), ("p_end", "-11.512925464970229(0.0000)"),
                                                   ("p_MM", "-0.000010000050000287824(1.0000)"), ("p_IM", "-1e-5(1.0000)"), ("p_DM", "-0.00001(1.0000)"),
                                                   ("p_MI", "-inf(0.0000)"), ("p_II", "-inf(0.0000)"), ("p_DI", "-inf(0.0000)"), ("p_MD", "-inf(0.0000)"),
                                                   ("p_ID", "-inf(0.0000)"), ("p_DD", "-inf(0.0000)")])
    text = ('{"genome":["L:ACGT"],"genome_size":4,"reads":{"reads":["ACG:+:0-0,0-1,0-2","CG"]},"phmm_params":{' + probs +
            ',"n_active_nodes":40,"active_node_max_ratio":30.0,"n_warmup":50,"n_max_gaps":4}}')
    d = H.Dataset.from_json_str(text)
    q = d.params()
    assert q.p_mismatch == -np.inf and q.p_match == 0.0 and q.p_MM == -0.000010000050000287824 and q.p_IM == -1e-5 and q.warmup_threshold == 200
    assert [bytes(d.reads()[i]) for i in range(2)] == [b"ACG", b"CG"]
    assert d.read_origins()[1][1] == [None, None]      # a plain sequence carries no origins
    for bad in ['{"genome":["X:ACGT"]', text.replace('"L:ACGT"', '"Q:ACGT"'), text.replace("0-0,0-1,0-2", "0-0,0-1"), text.replace('"ACG:+:', '"AcG:+:'),
                text.replace('"genome_size":4', '"genome_size":-4'), text[:-1], text + "x", text.replace('"n_warmup":50,', "")]:
        with pytest.raises(H.DbgphmmError):
            H.Dataset.from_json_str(bad)


def test_text_readers_under_address_and_ub_sanitizers(tmp_path):
    """tests/native/fuzz_formats.cpp: csrc/formats.cu and csrc/score.cu compiled as plain C++ with -fsanitize=address,undefined (leak
    check included); 60,000 seeded mutations of a DBG text, a MAP text and a dataset JSON are parsed, whatever parses is walked, written
    back and parsed again.  (3,000,000 mutations were run once: no finding.)"""
    import shutil
    import subprocess
    from dbgphmm_b200 import hmmv2 as H
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fuzz_formats")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I", cuda_inc, "-x", "c++",
           os.path.join(root, "dbgphmm_b200/csrc/formats.cu"), os.path.join(root, "dbgphmm_b200/csrc/score.cu"), os.path.join(root, "tests/native/fuzz_formats.cpp"),
           "-lz", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    (tmp_path / "seed.dbg").write_text(README_DBG)
    (tmp_path / "seed.map").write_text("# c\n0\t0\tA\t1:-0.5,2:-1.25\n0\t1\tC\t3:-inf\n1\t0\tG\t\n1\t1\tT\t7:0\n")
    d = H.Dataset.new([b"ATCGATTTAGC", b"GGGC"], ["L", "C"], [b"ATCGT", b"TTAG"], H.params_uniform(0.001), revcomp=[1, 0],
                      origins=[[(0, 0), (0, 1), (0, 2), None, (0, 3)], [(0, 5), (0, 6), (0, 7), (0, 9)]])
    (tmp_path / "seed.json").write_text(d.to_json_string())
    r = subprocess.run([exe, str(tmp_path / "seed.dbg"), str(tmp_path / "seed.map"), str(tmp_path / "seed.json"), "60000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    counts = [int(x) for x in r.stdout.replace(";", " ").split() if x.isdigit()]
    assert len(counts) == 6 and all(c > 100 for c in counts), r.stdout      # every reader both accepted and rejected mutants
