"""Static resource report of every kernel of the library: registers, spills, shared memory, as `ptxas -v` prints them.

    python tools/ptxas_summary.py > profiles/rN_ptxas.txt

Compiles dbgphmm_b200/csrc/*.cu with the flags of dbgphmm_b200/build.py plus `-Xptxas -v` into build/ptxas/ (git-ignored;
needs no GPU) and prints one line per kernel.  Registers per thread bound the resident warps per SM
(64 K registers / (threads x registers)); a non-zero spill count in a hot kernel is the first thing to look at before
spending GPU time on it (B200_PROFILING.md).
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dbgphmm_b200 import build as B  # noqa: E402


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    out_dir = os.path.join(ROOT, "build", "ptxas")
    os.makedirs(out_dir, exist_ok=True)
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)
    procs = []
    for src in B.SOURCES:
        log = os.path.join(out_dir, src.replace(".cu", ".log"))
        if "--reuse" in sys.argv and os.path.exists(log):
            continue
        cmd = [B._nvcc()] + B.NVCC_FLAGS + ["-ccbin", "/usr/bin/g++", "-Xptxas", "-v", "-c", os.path.join(B.CSRC, src),
                                             "-o", os.path.join(out_dir, src.replace(".cu", ".o"))]
        procs.append((log, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for log, p in procs:
        text, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(text)
            raise SystemExit(f"nvcc failed, see {log}")
        open(log, "w").write(text)
    rows = []
    for src in B.SOURCES:
        text = open(os.path.join(out_dir, src.replace(".cu", ".log"))).read()
        for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'\n(.*?)(?=ptxas info\s+: Compiling|ptxas info\s+: Function properties for (?!_Z)|\Z)",
                             text, flags=re.S):
            name, body = m.group(1), m.group(2)
            regs = re.search(r"Used (\d+) registers", body)
            spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", body)
            smem = re.search(r"(\d+) bytes smem", body)
            bars = re.search(r"used (\d+) barriers", body)
            sp = tuple(int(x) for x in spill.groups()) if spill else (0, 0, 0)
            rows.append((src, name, int(regs.group(1)) if regs else -1) + sp +
                        (int(smem.group(1)) if smem else 0, int(bars.group(1)) if bars else 0))
    names = demangle([r[1] for r in rows])
    print(f"# ptxas -v, sm_100a, flags of dbgphmm_b200/build.py ({' '.join(B.NVCC_FLAGS)})")
    print(f"# {'file':<12} {'regs':>4} {'stack':>6} {'spill st':>8} {'spill ld':>8} {'static smem':>11} {'barriers':>8}  kernel")
    for src, name, regs, stack, sst, sld, smem, bars in rows:
        d = names[name]
        d = re.sub(r"\(.*$", "", d)          # drop the argument list
        d = re.sub(r"^void ", "", d)
        print(f"  {src:<12} {regs:>4} {stack:>6} {sst:>8} {sld:>8} {smem:>11} {bars:>8}  {d}")
    spilled = [r for r in rows if r[4] or r[5]]
    print(f"# {len(rows)} kernels, {len(spilled)} with spills" + (": " + ", ".join(re.sub(r'^void ', '', re.sub(r'\(.*$', '', names[r[1]])) for r in spilled) if spilled else ""))


if __name__ == "__main__":
    main()
