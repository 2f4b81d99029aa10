"""Ad-hoc check of the fast dense kernels against the exact one on a scaled C3 graph (test infrastructure)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dbgphmm_b200 import hmmv2 as H
from dbgphmm_b200 import synth
from tests.common import gpu_model, oracle_params
w = synth.make_workload("C3m", int(sys.argv[1]), 40, 1, 3_000, 0.001, ploidy=2, het=0.01, seed=5, n_reads=2)
par = oracle_params(0.001, n_warmup=w.k)
g = gpu_model(w.graph, par)
read = w.reads[0][:6]
out = {}
for mode in ("1", "0"):
    os.environ["DBGPHMM_FORCE_EXACT"] = mode
    for name, fn in (("fwd", g.forward), ("bwd", g.backward)):
        t = fn(read)
        rows = [t.row(i) for i in range(len(read))]
        out[(mode, name)] = rows
import collections
a, b = out[("1", "fwd")][0], out[("0", "fwd")][0]
bad = np.where(np.isnan(b.m) | np.isnan(b.i) | np.isnan(b.d))[0]
print("n_bad", len(bad))
sg = w.graph
indeg = np.bincount(sg.dst, minlength=sg.n_nodes); outdeg = np.bincount(sg.src, minlength=sg.n_nodes)
for v in bad[:12]:
    print(v, "m", b.m[v], "i", b.i[v], "d", b.d[v], "| exact", a.m[v], a.i[v], a.d[v], "| indeg", indeg[v], "outdeg", outdeg[v])
print("indeg hist of bad", collections.Counter(indeg[bad].tolist()), "outdeg", collections.Counter(outdeg[bad].tolist()))
print("graph indeg hist", collections.Counter(indeg.tolist()))
