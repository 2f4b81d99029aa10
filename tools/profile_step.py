"""One bulk run_sparse step on the C3 workload with the library's phase timings (used for ncu captures too)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dbgphmm_b200 import hmmv2 as H
if os.environ.get("DBGPHMM_LIB_PATH"):   # tuning sweeps: a library built by tools/build_variant.py
    H.LIB_PATH = os.environ["DBGPHMM_LIB_PATH"]
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=32)
ap.add_argument("--genome-len", type=int, default=1_000_000)
ap.add_argument("--read-len", type=int, default=10_000)
ap.add_argument("--k", type=int, default=40)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--mode", default="sparse")
ap.add_argument("--budget-gb", type=float, default=0)
a = ap.parse_args()
a.reads_per_gpu = a.reads
t0 = time.time()
g, li, lt, reads = bench.make_inputs(a, 0, a.reads)
print(f"inputs: N={g.n_nodes} E={g.n_edges} reads={len(reads)} built in {time.time()-t0:.1f}s", flush=True)
par = H.params_uniform(0.001); par.n_warmup = a.k
t0 = time.time()
m = H.PHMMModel(g.src, g.dst, g.base, li, lt, par, mem_budget_bytes=int(a.budget_gb * 2**30))
print(f"model_create {time.time()-t0:.2f}s", flush=True)
rd = H.Reads(reads)
for rep in range(a.reps):
    t0 = time.time()
    fr, lf, lb, cells = m.run_node_freqs(rd, a.mode)
    wall = time.time() - t0
    d, s, p, tot, dc = H.last_timing()
    km, kl, kc = H.last_dense_kernel()
    print(f"rep {rep}: wall {wall*1e3:.0f} ms | total {tot:.0f} dense {d:.0f} sparse {s:.0f} product {p:.0f} | dense kernel {km:.0f} ms / {kl} launches "
          f"({kc/ max(km,1e-9)/1e6:.2f} GCUPS) | cells fwd {cells[0]:.3e} bwd {cells[1]:.3e} -> {sum(cells)/wall/1e9:.2f} GCUPS | launches {H.launch_count()}", flush=True)
print("logp", lf[:3], lb[:3], "freq sum", fr.sum())
