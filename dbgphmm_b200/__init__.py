"""dbgphmm_b200 — B200-native (sm_100a) implementation of dbgphmm's hmmv2 read-likelihood hot path.

Layout: csrc/ (CUDA kernels + C ABI, built to lib/libdbgphmm_b200.so), hmmv2.py (host-side mirror of the
reference's `impl PHMMModel` surface over the C ABI), graphs.py / synth.py (callers' side: graph builders
and synthetic workloads).  See DESIGN.md.
"""
