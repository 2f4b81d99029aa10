"""Sparse-kernel residency sweep on the C3 workload: entry capacity x threads per job (one process, inputs built once).

    python tools/sweep_sparse.py [--reads 1480]        -> gpurun_out/sweep_sparse.log
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbgphmm_b200 import hmmv2 as H
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1480)
ap.add_argument("--genome-len", type=int, default=1_000_000)
ap.add_argument("--read-len", type=int, default=10_000)
ap.add_argument("--k", type=int, default=40)
ap.add_argument("--configs", default="256:64,128:64,128:32,96:64,96:32")
a = ap.parse_args()
a.reads_per_gpu = a.reads
g, li, lt, reads = bench.make_inputs(a, 0, a.reads)
par = H.params_uniform(0.001); par.n_warmup = a.k
m = H.PHMMModel(g.src, g.dst, g.base, li, lt, par)
rd = H.Reads(reads)
os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/sweep_sparse.log", "w")
ref = None
for cfg in a.configs.split(","):
    cap, thr = cfg.split(":")
    os.environ["DBGPHMM_SPARSE_CAP"] = cap
    os.environ["DBGPHMM_SPARSE_THREADS"] = thr
    for rep in range(2):
        t0 = time.time()
        fr, lf, lb, cells = m.run_node_freqs(rd, "sparse")
        wall = time.time() - t0
        d, s, p, tot, dc = H.last_timing()
    if ref is None:
        ref = (fr.copy(), lf.copy())
    same = bool((fr == ref[0]).all() and (lf == ref[1]).all())
    line = (f"cap {cap} threads {thr}: wall {wall*1e3:.0f} ms total {tot:.0f} dense {d:.0f} sparse {s:.0f} product {p:.0f} "
            f"-> {sum(cells)/wall/1e9:.2f} GCUPS  identical_to_first={same}")
    print(line, flush=True); out.write(line + "\n"); out.flush()
