"""The C++ host mirror (include/dbgphmm_b200.hpp) compiles against the C ABI and links with the built library."""
import os
import subprocess

from dbgphmm_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_compiles_and_links(tmp_path):
    B.build()
    src = tmp_path / "t.cpp"
    src.write_text(r'''#include "dbgphmm_b200.hpp"
#include <cmath>
int main() {
    auto p = dbgphmm::uniform(0.01);
    if (!(p.n_max_gaps == 4 && p.n_active_nodes == 40)) return 1;
    // README.md:174-191 (toy::repeat) through the host-only format entry points
    auto d = dbgphmm::MultiDbg::from_dbg_str("K\t4\nN\t0\tnnn\nN\t1\tCAG\nE\t0\t1\t0\tCAGGAAnnn\t1\t9,10,11,12,13,14\n"
                                            "E\t1\t1\t1\tCAGCAG\t3\t6,7,8\nE\t2\t0\t1\tnnnTCCCAG\t1\t0,1,2,3,4,5\n");
    if (d->k() != 4 || d->n_edges_full() != 15 || d->n_edges_compact() != 3) return 2;
    auto full = d->expand_copy_nums(1, {1, 2, 1});
    if (full.size() != 15 || full[6] != 2 || full[0] != 1 || full[14] != 1) return 3;
    auto d2 = dbgphmm::MultiDbg::from_dbg_str(d->to_dbg_string());
    if (d2->to_dbg_string() != d->to_dbg_string()) return 4;
    try { d->set_copy_nums({1, 3, 2}); return 5; } catch (const dbgphmm::Error&) {}
    // Mapping::map_nodes (hint.rs:234-270, case 1: v -> [v + 1]) through the mirror
    {
        const uint64_t read_off[2] = {0, 2}, row_off[3] = {0, 2, 4};
        const uint32_t nodes[4] = {0, 1, 2, 3};
        const double logp[4] = {std::log(0.6), std::log(0.4), std::log(0.9), std::log(0.1)};
        dbgphmm_mappings* h = nullptr;
        if (dbgphmm_mappings_create(1, read_off, row_off, nodes, logp, &h) != DBGPHMM_OK) return 6;
        dbgphmm::Mappings mp(h);
        dbgphmm::Mappings m1 = mp.map_nodes({{1}, {2}, {3}, {4}});
        uint64_t nr = 0, nrow = 0, nent = 0;
        if (dbgphmm_mappings_sizes(m1.handle(), &nr, &nrow, &nent) != DBGPHMM_OK || nr != 1 || nrow != 2 || nent != 4) return 7;
        uint64_t ro[2], rw[3]; uint32_t nd[4]; double lp[4];
        if (dbgphmm_mappings_export(m1.handle(), ro, rw, nd, lp) != DBGPHMM_OK) return 8;
        if (!(nd[0] == 1 && nd[1] == 2 && nd[2] == 3 && nd[3] == 4 && lp[0] == logp[0] && lp[3] == logp[3])) return 9;
    }
    return 0;
}
''')
    exe = tmp_path / "t"
    libdir = os.path.join(ROOT, "dbgphmm_b200", "lib")
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir, "-ldbgphmm_b200",
                           f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-lcudart", "-o", str(exe)], env=env)
    assert subprocess.call([str(exe)]) == 0
