"""Ad-hoc GPU-vs-oracle diagnostics (prints the first divergence in detail).  Test infrastructure: lives under tests/ because it uses the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dbgphmm_b200 import hmmv2 as H, synth, graphs
from oracle import oracle as O
from tests.common import *


def first_diff(gt, ot, name):
    for r in range(len(ot)):
        a, b = gt.row(r), ot.row(r)
        if a.is_dense != b.is_dense:
            print(name, "row", r, "density", a.is_dense, b.is_dense); return r
        if not b.is_dense:
            if list(a.ids) != list(b.ids) or list(a.ids_d) != list(b.ids_d):
                print(name, "row", r, "ids differ")
                print(" gpu ids ", list(map(int, a.ids))); print(" ref ids ", list(map(int, b.ids)))
                print(" gpu idsd", list(map(int, a.ids_d))); print(" ref idsd", list(map(int, b.ids_d)))
                p = ot.row(r - 1) if r > 0 else None
                q = gt.row(r - 1) if r > 0 else None
                if p is not None and not p.is_dense:
                    mo = p.merged(gt.n_nodes); mg = q.merged(gt.n_nodes)
                    order = np.argsort(-mo, kind="stable")[:100]
                    print(" prev merged (ref) top:", [(int(k), float(mo[k]), float(mg[k])) for k in order if np.isfinite(mo[k]) or np.isfinite(mg[k])][:90])
                return r
    print(name, "all rows equal"); return -1


def case1():
    w = synth.make_workload("t", 600, 12, 4, 150, 0.01, ploidy=2, het=0.02, seed=3, n_reads=4)
    par = oracle_params(0.01, n_warmup=w.k, n_active=80)
    g, o = gpu_model(w.graph, par), oracle_model(w.graph, par)
    for i, read in enumerate(w.reads[:3]):
        first_diff(g.backward_sparse(read), o.backward_sparse(read), f"case1 bwd read{i}")


def case2():
    w = synth.make_workload("t", 600, 16, 4, 150, 0.003, ploidy=2, het=0.02, seed=5, n_reads=4)
    par = oracle_params(0.001, n_warmup=w.k, warmup_threshold=30)
    g, o = gpu_model(w.graph, par, "non_zero"), oracle_model(w.graph, par, "non_zero")
    reads = w.reads[:3]
    gm2 = g.generate_mappings(H.Reads(reads), None, False)
    om2 = o.generate_mappings(O.Reads(reads), None, False)
    bad = np.nonzero(gm2.nodes != om2.nodes)[0]
    print("case2 n diff", len(bad), "of", len(om2.nodes))
    if len(bad):
        e = bad[0]
        row = np.searchsorted(om2.row_off, e, side="right") - 1
        a, b = int(om2.row_off[row]), int(om2.row_off[row + 1])
        print(" row", row, "entry", e - a)
        print(" gpu", list(zip(gm2.nodes[a:b].tolist(), np.round(gm2.probs[a:b], 3).tolist())))
        print(" ref", list(zip(om2.nodes[a:b].tolist(), np.round(om2.probs[a:b], 3).tolist())))
        rd = np.searchsorted(om2.read_off, row, side="right") - 1
        r = row - int(om2.read_off[rd])
        read = reads[rd]
        gf = g.forward_sparse(read, False); of = o.forward_sparse(read, False)
        gb = g.backward_by_forward(read, gf); ob = o.backward_by_forward(read, of)
        first_diff(gf, of, "case2 fwd"); first_diff(gb, ob, "case2 bwd")
        fr, br = of.row(r), ob.row(r + 1) if r + 1 < len(ob) else None
        print(" F row", r, "dense", fr.is_dense, "n", None if fr.is_dense else (len(fr.ids), len(fr.ids_d)))
        if br is not None:
            print(" B row", r + 1, "dense", br.is_dense, "n", None if br.is_dense else (len(br.ids), len(br.ids_d)))


def case3():
    w = synth.make_workload("t", 600, 16, 4, 150, 0.003, ploidy=2, het=0.02, seed=8, n_reads=5)
    sg = w.graph
    par = oracle_params(0.001, n_warmup=w.k, warmup_threshold=30)
    o = oracle_model(sg, par, "non_zero")
    rng = np.random.default_rng(0)
    X = np.stack([sg.node_copy_num] + [np.maximum(0, sg.node_copy_num + rng.integers(-1, 2, sg.n_nodes)) for _ in range(5)])
    for x in range(len(X)):
        li, lt = sg.to_probs("normal", X[x]); o.set_probs(li, lt)
        try:
            s, p = o.to_full_prob_reads(O.Reads(w.reads), None, True)
            print("case3 oracle x", x, s)
        except Exception as ex:
            print("case3 oracle x", x, "raised", ex)
        print("   returned", s)


def case4():
    sg, seq = random_linear_graph(120, 3)
    par = oracle_params(0.01, n_warmup=40)
    g, o = gpu_model(sg, par), oracle_model(sg, par)
    for read in (seq[5:6], seq[10:25]):
        try:
            gf = g.forward_sparse(read, True)
            first_diff(gf, o.forward_sparse(read, True), "case4 fwd")
            first_diff(g.backward_by_forward(read, gf), o.backward_by_forward(read, o.forward_sparse(read, True)), "case4 bwd")
        except Exception as ex:
            print("case4", len(read), "raised", ex)


if __name__ == "__main__":
    for c in sys.argv[1:] or ["1", "2", "3", "4"]:
        try:
            globals()["case" + c]()
        except Exception as ex:
            import traceback; traceback.print_exc()
