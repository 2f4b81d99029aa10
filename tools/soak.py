"""Soak of the bench shard through the C ABI (no torch): STEPS passes of run_sparse + node freqs over the C3 shard, every pass
compared with the first one.  On a failure prints the library's error (with CUDA_LAUNCH_BLOCKING=1: the failing launch).
    python tools/soak.py [steps] [reads] [seed] [genome_len] [read_len]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dbgphmm_b200 import hmmv2 as H, synth, graphs

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 100
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
SEED = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
GL = int(sys.argv[4]) if len(sys.argv) > 4 else 1_000_000
RL = int(sys.argv[5]) if len(sys.argv) > 5 else 10_000
h0 = synth.random_genome(GL, 0); h1 = synth.mutate_substitutions(h0, 0.01, 1)
g, _ = graphs.build_dbg([h0.tobytes(), h1.tobytes()], 40, seed=100)
cov = R * RL / (2.0 * GL)
reads = synth.sample_reads([h0, h1], cov, RL, 0.001, SEED)[:R]
while len(reads) < R:
    reads += synth.sample_reads([h0, h1], cov, RL, 0.001, 5000 + SEED + len(reads))[:R - len(reads)]
li, lt = g.to_probs("normal")
par = H.params_uniform(0.001); par.n_warmup = 40
m = H.PHMMModel(g.src, g.dst, g.base, li, lt, par)
rd = H.Reads(reads)
ref = None
t0 = time.time()
for s in range(STEPS):
    try:
        fr, lf, lb, cells = m.run_node_freqs(rd, "sparse")
    except Exception as e:
        print(f"SOAK FAILED at step {s}: {e!r}", flush=True)
        sys.exit(3)
    if ref is None:
        ref = (fr, lf, lb, cells)
        print("step 0: cells", cells, "sum lf", float(lf.sum()), "sum freq", float(fr.sum()), flush=True)
    else:
        ok = np.allclose(fr, ref[0], rtol=1e-9, atol=1e-9) and np.array_equal(lf, ref[1]) and np.array_equal(lb, ref[2]) and cells == ref[3]
        if not ok:
            bad = int(np.sum(lf != ref[1])); badb = int(np.sum(lb != ref[2]))
            print(f"SOAK MISMATCH at step {s}: {bad} lf / {badb} lb differ, max freq diff {np.abs(fr - ref[0]).max()}, cells {cells} vs {ref[3]}", flush=True)
    if s % 10 == 9:
        print(f"step {s} ok ({(time.time() - t0) / (s + 1):.2f} s/step)", flush=True)
print("soak done")
