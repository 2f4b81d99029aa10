"""The C++ host mirror (include/dbgphmm_b200.hpp) compiles against the C ABI and links with the built library."""
import os
import subprocess

from dbgphmm_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_compiles_and_links(tmp_path):
    B.build()
    src = tmp_path / "t.cpp"
    src.write_text('#include "dbgphmm_b200.hpp"\nint main() { auto p = dbgphmm::uniform(0.01); return p.n_max_gaps == 4 && p.n_active_nodes == 40 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    libdir = os.path.join(ROOT, "dbgphmm_b200", "lib")
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir, "-ldbgphmm_b200",
                           f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-lcudart", "-o", str(exe)], env=env)
    assert subprocess.call([str(exe)]) == 0
