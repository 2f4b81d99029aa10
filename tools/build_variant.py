#!/usr/bin/env python
"""Developer tool: build the library with extra -D flags into dbgphmm_b200/lib/variants/<name>/ (tuning sweeps).

usage: python tools/build_variant.py NAME -DWT_WARPS=7 ...     then     DBGPHMM_LIB_PATH=<that .so> python tools/profile_step.py
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from dbgphmm_b200 import build as B  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(B.LIBDIR, "variants", name)
os.makedirs(out, exist_ok=True)
env = dict(os.environ)
env.pop("CXX", None); env.pop("CC", None)
procs, objs = [], []
for src in B.SOURCES:
    o = os.path.join(out, src.replace(".cu", ".o"))
    objs.append(o)
    procs.append(subprocess.Popen([B._nvcc()] + B.NVCC_FLAGS + flags + ["-ccbin", "/usr/bin/g++", "-c", os.path.join(B.CSRC, src), "-o", o], env=env))
if any(p.wait() for p in procs):
    sys.exit("nvcc failed")
lib = os.path.join(out, "libdbgphmm_b200.so")
subprocess.check_call([B._nvcc(), "-shared", "-ccbin", "/usr/bin/g++", "-o", lib] + objs + ["-lcudart"], env=env)
print(lib)
