"""Reduced bench shard for compute-sanitizer (memcheck / racecheck / initcheck): run_sparse + node freqs in the stream strategy.
No torch, no oracle: only the C ABI.   python tools/san_case.py [genome_len] [n_reads] [read_len] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dbgphmm_b200 import hmmv2 as H, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 96
RL = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
STEPS = int(sys.argv[4]) if len(sys.argv) > 4 else 2
w = synth.make_workload("san", L, 40, 4, RL, 0.001, ploidy=2, het=0.01, seed=7)
reads = w.reads
while len(reads) < R:
    reads = reads + synth.sample_reads(w.haplotypes, 4, RL, 0.001, 99 + len(reads))
reads = reads[:R]
li, lt = w.graph.to_probs("normal")
par = H.params_uniform(0.001); par.n_warmup = w.k
m = H.PHMMModel(w.graph.src, w.graph.dst, w.graph.base, li, lt, par)
rd = H.Reads(reads)
ref = None
for s in range(STEPS):
    fr, lf, lb, cells = m.run_node_freqs(rd, "sparse")
    print("step", s, "N", w.graph.n_nodes, "reads", len(reads), "cells", cells, "sum lf", float(lf.sum()), "sum freq", float(fr.sum()), flush=True)
    if ref is None:
        ref = (fr, lf, lb)
    else:
        assert np.allclose(fr, ref[0], rtol=1e-9, atol=1e-12) and np.array_equal(lf, ref[1]) and np.array_equal(lb, ref[2])
print("ok")
