"""The library's HOST logic under AddressSanitizer + UBSan on a machine without a GPU: csrc/model.cu (relabelling, adjacency, the
tiling plans of the dense kernels, the recompute cones) and csrc/engine.cu (the stream-ordered block cache, the memory budget) are
compiled by nvcc with the sanitizers on the host side and linked with tests/native/cuda_stub.cpp -- device memory is malloc'ed, so
every upload and index computation is checked -- and tests/native/host_logic.cpp, which validates every table the dense kernels
index with against the graph.  First run found two bugs of the planner (a halo node left out of a tile's layout when its first
upstream neighbour sits on the other branch of a fork: out-of-bounds write; an endless climb on a cycle of halo nodes)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from dbgphmm_b200 import graphs, synth, hmmv2 as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAN = ["-fsanitize=address", "-fsanitize=undefined", "-fno-sanitize-recover=undefined"]


@pytest.fixture(scope="module")
def host_logic(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc) or shutil.which("g++") is None:
        pytest.skip("needs nvcc and g++")
    d = tmp_path_factory.mktemp("host_logic")
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "include")
    objs = []
    for name in ("model", "engine"):
        o = str(d / f"{name}.o")
        cmd = [nvcc, "-std=c++17", "-O1", "-g", "-gencode", "arch=compute_100a,code=sm_100a"] + [x for f in SAN for x in ("-Xcompiler", f)] + \
              ["-c", os.path.join(ROOT, "dbgphmm_b200/csrc", name + ".cu"), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        objs.append(o)
    for name in ("cuda_stub", "host_logic"):
        o = str(d / f"{name}.o")
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", *SAN, "-I", cuda_inc, "-c", os.path.join(ROOT, "tests/native", name + ".cpp"), "-o", o],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        objs.append(o)
    exe = str(d / "host_logic")
    r = subprocess.run(["g++", *SAN, *objs, "-o", exe, "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe, d


def dump(path, src, dst, base, li, lt, par):
    with open(path, "wb") as f:
        f.write(np.array([len(base), len(src)], np.uint32).tobytes())
        f.write(np.asarray(src, np.uint32).tobytes()); f.write(np.asarray(dst, np.uint32).tobytes())
        f.write(np.asarray(base, np.uint8).tobytes()); f.write(np.asarray(li, np.float64).tobytes()); f.write(np.asarray(lt, np.float64).tobytes())
        f.write(bytes(par))


def test_block_cache_and_budget(host_logic):
    exe, _ = host_logic
    r = subprocess.run([exe, "cache"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "cache ok" in r.stdout, r.stdout + r.stderr[-3000:]


def test_tiling_plans_of_many_graph_shapes(host_logic):
    exe, d = host_logic
    par = H.params_uniform(0.001); par.n_warmup = 40
    files = []

    def add(name, g_or_arrays, p=par, mode="normal"):
        path = str(d / (name + ".bin"))
        if isinstance(g_or_arrays, tuple):
            dump(path, *g_or_arrays, p)
        else:
            li, lt = g_or_arrays.to_probs(mode)
            dump(path, g_or_arrays.src, g_or_arrays.dst, g_or_arrays.base, li, lt, p)
        files.append(path)

    h0 = synth.random_genome(30000, 0); h1 = synth.mutate_substitutions(h0, 0.01, 1)
    add("diploid", graphs.build_dbg([h0.tobytes(), h1.tobytes()], 40, seed=100)[0])                 # the C3 graph, scaled down
    hap = synth.tandem_repeat_genome(1000, 6, 2000, seed=3, divergence=0.005); hap2 = synth.mutate_substitutions(hap, 0.002, 77)
    add("tandem", graphs.build_dbg([hap.tobytes(), hap2.tobytes()], 40, seed=9)[0], mode="non_zero")  # the C4 graph, scaled down
    a = synth.random_genome(3000, 11).tobytes()                                                      # microsatellites and a homopolymer longer than k
    g1 = a[:1000] + b"AT" * 60 + a[1000:2000] + b"A" * 100 + a[2000:2500] + b"CAG" * 40 + a[2500:]
    g2 = a[:1000] + b"AT" * 55 + a[1000:2000] + b"A" * 90 + a[2000:2500] + b"CAG" * 44 + a[2500:]
    add("micro", graphs.build_dbg([g1, g2], 40, seed=3)[0])
    par8 = H.params_uniform(0.01); par8.n_warmup = 8
    add("k8", graphs.build_dbg([synth.random_genome(3000, 5).tobytes()], 8, seed=1)[0], p=par8)      # branchy, cycles of halo nodes (used to hang)
    rng = np.random.default_rng(5)                                                                    # chain + cross edges (used to write out of bounds)
    n = 20000; src = list(range(n - 1)); dst = list(range(1, n))
    for _ in range(600):
        v = int(rng.integers(0, n)); src.append(v); dst.append(int(np.clip(v + rng.integers(-30, 30), 0, n - 1)))
    add("cross", (src, dst, rng.choice(np.frombuffer(b"ACGT", np.uint8), n), np.log(np.full(n, 1.0 / n)), np.log(rng.random(len(src)) * 0.9 + 0.05)))
    rng = np.random.default_rng(7)                                                                    # too dense for a tile: must be rejected, not crash
    n = 5000; src = []; dst = []
    for v in range(n):
        for c in rng.choice(n, size=rng.integers(0, 4), replace=False):
            src.append(v); dst.append(int(c))
    add("random", (src, dst, rng.choice(np.frombuffer(b"ACGTn", np.uint8), n), np.log(np.full(n, 1.0 / n)), np.log(rng.random(len(src)) + 1e-3)))
    add("loop", ([0, 0, 1], [0, 1, 0], np.frombuffer(b"AC", np.uint8), [np.log(.5)] * 2, [np.log(.5)] * 3))
    add("one", ([], [], np.frombuffer(b"A", np.uint8), [0.0], []))
    r = subprocess.run([exe, "plan", *files], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = r.stdout
    for name in ("diploid", "tandem", "micro", "k8", "cross", "loop", "one"):
        assert f"{name}.bin: N=" in out, out
    assert "random.bin: rejected (graph too dense" in out
    assert out.count("fwd2 tiles=") >= 6 and "not available" not in out.split("random.bin")[0]
    # tiling efficiency of the C3-like graph: with the branches of a bubble labelled next to each other a two-row tile owns ~138 of
    # its 160 positions (122 when every alternative branch was labelled at the end of the traversal)
    import re
    dip = out.split("diploid.bin")[1].split(".bin")[0]
    core = {m.group(1): float(m.group(2)) for m in re.finditer(r"(fwd2|bwd2|fwd|bwd)\s+tiles=\d+ core/tile=([0-9.]+)", dip)}
    assert core["fwd2"] > 132 and core["bwd2"] > 132 and core["fwd"] > 140 and core["bwd"] > 140, core
