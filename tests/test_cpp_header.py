"""The C++ host mirror (include/dbgphmm_b200.hpp) — the compiled-language host layer above the C ABI (the reference is Rust; no
Rust toolchain exists here).  One C++ program, written like the reference's own unit tests of the path
(hmm_forward_mock_linear_high_error forward.rs:599-619, hmm_backward_mock_linear_high_error backward.rs:607-628,
hmm_forward_with_hint_mock_linear_high_error forward.rs:640-669, hmm_backward_with_hint backward.rs:630-652, the hmm_freq_* invariants
freq.rs:434-610), with the graph of graph/mocks.rs:8-12 and the expected values of tests/golden/reference_kat.json embedded:

  * CPU (`not gpu`): it compiles, links against the built library, runs the host-only part (DBG text format, Mapping::map_nodes) and,
    on a box without a GPU, fails loudly in the PHMMModel constructor (status DBGPHMM_ERR_CUDA): there is no CPU fallback.
  * GPU (`-m gpu`): the whole program passes on the device."""
import os
import subprocess

import numpy as np
import pytest

from dbgphmm_b200 import build as B
from dbgphmm_b200 import graphs
from tests.common import kat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RC_NO_DEVICE = 100


def _arr(v, fmt):
    return "{" + ", ".join(fmt(x) for x in v) + "}"


def _f64(x):
    return "-INFINITY" if np.isneginf(x) else repr(float(x))


def _source():
    K = kat()
    sg = graphs.mock_linear()
    li, lt = sg.to_probs()
    fh, bh, hint, bhint = K["forward_high_error"], K["backward_high_error"], K["hint_mock_linear_high_error"], K["backward_with_hint"]
    assert fh["read"] == bh["read"] == hint["read"] == bhint["read"] and fh["p"] == bh["p"] == hint["p"] == bhint["p"]
    nodes = "{" + ", ".join(_arr(r, str) for r in hint["nodes"]) + "}"
    return r'''#include "dbgphmm_b200.hpp"
#include <cmath>
#include <cstdio>
#define CHECK(cond, code) do { if (!(cond)) { std::fprintf(stderr, "check failed (%d): %s\n", code, #cond); return code; } } while (0)
static bool near(double a, double b, double eps) { return std::fabs(a - b) < eps; }

static int host_only() {
    auto p = dbgphmm::uniform(0.01);
    CHECK(p.n_max_gaps == 4 && p.n_active_nodes == 40, 1);
    // README.md:174-191 (toy::repeat) through the host-only format entry points
    auto d = dbgphmm::MultiDbg::from_dbg_str("K\t4\nN\t0\tnnn\nN\t1\tCAG\nE\t0\t1\t0\tCAGGAAnnn\t1\t9,10,11,12,13,14\n"
                                            "E\t1\t1\t1\tCAGCAG\t3\t6,7,8\nE\t2\t0\t1\tnnnTCCCAG\t1\t0,1,2,3,4,5\n");
    CHECK(d->k() == 4 && d->n_edges_full() == 15 && d->n_edges_compact() == 3, 2);
    auto full = d->expand_copy_nums(1, {1, 2, 1});
    CHECK(full.size() == 15 && full[6] == 2 && full[0] == 1 && full[14] == 1, 3);
    auto d2 = dbgphmm::MultiDbg::from_dbg_str(d->to_dbg_string());
    CHECK(d2->to_dbg_string() == d->to_dbg_string(), 4);
    try { d->set_copy_nums({1, 3, 2}); return 5; } catch (const dbgphmm::Error&) {}
    // Mapping::map_nodes (hint.rs:234-270, case 1: v -> [v + 1]) through the mirror
    dbgphmm::Mapping m0;
    m0.nodes = {{0, 1}, {2, 3}};
    m0.probs = {{std::log(0.6), std::log(0.4)}, {std::log(0.9), std::log(0.1)}};
    dbgphmm::Mappings mp(std::vector<dbgphmm::Mapping>{m0});
    dbgphmm::Mappings m1 = mp.map_nodes({{1}, {2}, {3}, {4}});
    CHECK(m1.n_reads() == 1, 6);
    dbgphmm::Mapping a = m1.at(0);
    CHECK(a.nodes.size() == 2 && a.nodes[0] == std::vector<uint32_t>({1, 2}) && a.nodes[1] == std::vector<uint32_t>({3, 4}), 7);
    CHECK(a.probs[0][0] == m0.probs[0][0] && a.probs[1][1] == m0.probs[1][1], 8);
    // n_euler_circuits_test_toy (multi_dbg.rs:2320-2328), genome size, prior
    CHECK(near(std::exp(d->n_euler_circuits()), 1.0, 1e-4), 40);
    CHECK(d->genome_size() == 18 && near(d->to_prior(18, 3), -0.5 * std::log(2.0 * M_PI * 9.0), 1e-15), 41);
    auto ne = d->n_euler_circuits(2, {1, 3, 1, 0, 0, 0});
    CHECK(near(std::exp(ne[0]), 1.0, 1e-4) && std::exp(ne[1]) == 0.0, 42);
    dbgphmm::Score sc; sc.likelihood = -10; sc.prior = -2; sc.n_euler_circuits = std::log(5.0);
    CHECK(near(sc.p(), -12 + std::log(5.0), 1e-15), 43);
    // Posterior bookkeeping (posterior.rs:82-161)
    dbgphmm::Posterior post;
    CHECK(post.p() == -INFINITY, 44);
    dbgphmm::Score s2; s2.likelihood = -11; s2.prior = -1.5;
    post.add({{1, 5, 4}, sc}); post.add({{1, 4, 4}, s2}); post.add({{1, 5, 4}, s2});   // the third is a duplicate: ignored
    CHECK(post.samples().size() == 2 && post.contains({1, 4, 4}) && !post.contains({0, 0, 0}), 45);
    CHECK(near(post.p(), dbgphmm::prob_add(sc.p(), s2.p()), 1e-12) && post.max_copy_nums() == std::vector<uint32_t>({1, 5, 4}), 46);
    CHECK(near(post.p_edge_x(1, 5), sc.p() - post.p(), 1e-12) && near(post.p_edge_x(0, 1), 0.0, 1e-12) && post.p_edge_x(2, 7) == -INFINITY, 47);
    auto f = mp.to_node_freqs(5);   // hint.rs:161-171
    CHECK(near(f[0], 0.6, 1e-15) && near(f[2], 0.9, 1e-15) && f[4] == 0.0, 9);
    // Dataset JSON (e2e.rs:123-130): styled sequences and positioned reads in the reference's serde forms (collection.rs:836-861)
    const std::string dj = "{\"genome\":[\"C:ATCGAT\",\"L:GGGC\"],\"genome_size\":10,\"reads\":{\"reads\":[\"ATCGT:-:0-0,0-1,0-2,I,0-3\"]},"
        "\"phmm_params\":{\"p_mismatch\":\"-4.605170185988091(0.0100)\",\"p_match\":\"-0.01005033585350145(0.9900)\",\"p_random\":\"-1.3862943611198906(0.2500)\","
        "\"p_gap_open\":\"-4.605170185988091(0.0100)\",\"p_gap_ext\":\"-4.605170185988091(0.0100)\",\"p_end\":\"-11.512925464970229(0.0000)\","
        "\"p_MM\":\"-0.02021270866322344(0.9800)\",\"p_IM\":\"-0.02021270866322344(0.9800)\",\"p_DM\":\"-0.02021270866322344(0.9800)\","
        "\"p_MI\":\"-4.605170185988091(0.0100)\",\"p_II\":\"-4.605170185988091(0.0100)\",\"p_DI\":\"-4.605170185988091(0.0100)\","
        "\"p_MD\":\"-4.605170185988091(0.0100)\",\"p_ID\":\"-4.605170185988091(0.0100)\",\"p_DD\":\"-4.605170185988091(0.0100)\","
        "\"n_active_nodes\":40,\"active_node_max_ratio\":30.0,\"n_warmup\":50,\"warmup_threshold\":200,\"n_max_gaps\":4}}";
    auto ds = dbgphmm::Dataset::from_json_str(dj);
    CHECK(ds.genome_size() == 10 && ds.n_reads() == 1 && ds.genome() == std::vector<std::string>({"C:ATCGAT", "L:GGGC"}) && ds.reads()[0] == "ATCGT", 48);
    CHECK(ds.params().n_warmup == 50 && near(ds.params().p_mismatch, std::log(0.01), 1e-15) && near(ds.coverage(), 0.5, 1e-15), 49);
    CHECK(dbgphmm::Dataset::from_json_str(ds.to_json_string()).to_json_string() == ds.to_json_string(), 50);
    return 0;
}

int main() {
    if (int rc = host_only()) return rc;
    const std::vector<uint32_t> src = ''' + _arr(sg.src, str) + r''', dst = ''' + _arr(sg.dst, str) + r''';
    const std::vector<uint8_t> emission = ''' + _arr(sg.base, str) + r''';   // "''' + K["mock_linear"]["seq"] + r'''" (graph/mocks.rs:8-12)
    const std::vector<double> log_init = ''' + _arr(li, _f64) + r''', log_trans = ''' + _arr(lt, _f64) + r''';
    const double eps = ''' + repr(fh["eps"]) + r''';
    const std::string read = "''' + fh["read"] + r'''", read2 = "''' + fh["read2"] + r'''";
    std::unique_ptr<dbgphmm::PHMMModel> phmm;
    try {
        phmm = std::make_unique<dbgphmm::PHMMModel>(src, dst, emission, log_init, log_trans, dbgphmm::uniform(''' + repr(fh["p"]) + r'''));
    } catch (const dbgphmm::Error& e) {
        std::fprintf(stderr, "PHMMModel: status %d: %s\n", e.status, e.what());
        return e.status == DBGPHMM_ERR_CUDA && dbgphmm_device_count() == 0 ? ''' + str(RC_NO_DEVICE) + r''' : 10;
    }
    CHECK(phmm->n_nodes() == 10 && phmm->n_edges() == 9, 11);

    // hmm_forward_mock_linear_high_error (forward.rs:599-619)
    auto r1 = phmm->forward(read);
    CHECK(r1->n_emissions() == 5, 12);
    CHECK(near(r1->table(4).e, ''' + repr(fh["e"][0][1]) + r''', eps) && near(r1->full_prob(), ''' + repr(fh["e"][0][1]) + r''', eps), 13);
    CHECK(near(r1->table(4).m[7], ''' + repr(fh["m"][0][2]) + r''', eps), 14);
    auto r2 = phmm->forward(read2);
    CHECK(near(r2->table(4).e, ''' + repr(fh["e2"][0][1]) + r''', eps), 15);
    CHECK(near(r1->table(3).e, r2->table(3).e, 1e-12), 16);                       // the common prefix gives the same rows
    CHECK(r1->table(0).is_dense && r1->table(0).m.size() == 10 && r1->init_table().mb == 0.0, 17);

    // hmm_backward_mock_linear_high_error (backward.rs:607-628)
    auto b1 = phmm->backward(read);
    CHECK(near(b1->table(0).m[2], ''' + repr(bh["m"][0][2]) + r''', eps), 18);
    CHECK(near(b1->table(0).mb, ''' + repr(bh["mb"][0][1]) + r''', eps) && near(b1->full_prob(), ''' + repr(bh["mb"][0][1]) + r''', eps), 19);
    CHECK(near(phmm->backward(read2)->table(0).mb, ''' + repr(bh["mb2"][0][1]) + r''', eps), 20);

    // hmm_forward_with_hint_mock_linear_high_error (forward.rs:640-669): exact node lists, then the mapping as a hint
    dbgphmm::PHMMOutput o = phmm->run(read);
    dbgphmm::Mapping hint = o.to_mapping(''' + str(hint["n_active"]) + r''');
    const std::vector<std::vector<uint32_t>> want = ''' + nodes + r''';
    CHECK(hint.nodes == want, 21);
    dbgphmm::Mappings hints(std::vector<dbgphmm::Mapping>{hint});
    auto r3 = phmm->forward_with_mapping(read, hints, 0);
    CHECK(!r3->table(2).is_dense && r3->table(2).ids.size() == 3, 22);
    CHECK(near(r1->full_prob(), r3->full_prob(), ''' + repr(hint["max_log_diff_dense_vs_hint"]) + r'''), 23);
    CHECK(near(phmm->forward_with_mapping_score_only(read, hint), r3->full_prob(), 1e-9), 24);
    // hmm_backward_with_hint_mock_linear_high_error (backward.rs:630-652)
    dbgphmm::Mappings hints5(std::vector<dbgphmm::Mapping>{o.to_mapping(''' + str(bhint["n_active"]) + r''')});
    CHECK(near(b1->full_prob(), phmm->backward_with_mapping(read, hints5, 0)->full_prob(), ''' + repr(bhint["max_log_diff"]) + r'''), 25);

    // freq.rs:434-610: forward and backward totals agree, node frequencies follow the true path, the reads' products add up
    CHECK(near(o.to_full_prob_forward(), o.to_full_prob_backward(), 1e-2), 26);
    std::vector<double> nf = o.to_node_freqs();
    double tot = 0; for (double v : nf) tot += v;
    CHECK(nf.size() == 10 && nf[5] > 0.9 && nf[0] < 0.1 && tot > 4.5 && tot < 5.6, 27);
    auto ei = o.to_edge_and_init_freqs();
    CHECK(ei.first.size() == 9 && ei.second.size() == 10 && ei.first[4] > 0.85 && ei.first[0] < 0.01 && ei.second[3] > 0.8 && ei.second[9] < 0.01, 28);   // 4 -> 5 used, read starts at node 3
    std::vector<double> q = phmm->q_score_exact(ei.first, ei.second);
    CHECK(q.size() == 3 && std::isfinite(q[0]) && q[1] <= 0.0 && q[2] == 0.0, 29);
    dbgphmm::PHMMOutput os = phmm->run_sparse(read);   // shorter than n_warmup: every row dense, identical to run()
    CHECK(near(os.to_full_prob_forward(), o.to_full_prob_forward(), 1e-9), 30);
    CHECK(near(phmm->run_sparse_adaptive(read, true).to_full_prob_backward(), o.to_full_prob_backward(), 1e-9), 31);
    CHECK(phmm->run_with_mapping(read, hints, 0).n_emissions() == 5, 32);
    CHECK(r1->top_nodes(4, 1) == std::vector<uint32_t>({7}), 33);
    dbgphmm::Reads reads(std::vector<std::string>{read, read2});
    const double both = r1->full_prob() + r2->full_prob();
    CHECK(near(phmm->to_full_prob(reads), both, 1e-9) && near(phmm->to_full_prob_parallel(reads), both, 1e-9), 34);
    CHECK(near(phmm->to_full_prob_sparse(reads, false), both, 1e-9), 35);
    CHECK(near(phmm->to_full_prob_sparse_backward(reads), b1->full_prob() + phmm->backward(read2)->full_prob(), 1e-9), 36);
    CHECK(near(phmm->forward_sparse_score_only(read2, false), r2->full_prob(), 1e-9), 37);
    std::vector<double> nf2 = phmm->to_node_freqs(reads), nfb = phmm->run(read2).to_node_freqs();
    for (int v = 0; v < 10; v++) CHECK(near(nf2[v], nf[v] + nfb[v], 1e-9), 38);
    dbgphmm::Mappings gm = phmm->generate_mappings(reads, nullptr, false);   // hint.rs:193-220
    CHECK(gm.n_reads() == 2 && gm.at(0).nodes.size() == 5 && gm.at(0).nodes[0][0] == 3, 39);
    std::puts("cpp mirror ok");
    return 0;
}
'''


def _build_and_run(tmp_path, source=None):
    B.build()
    src = tmp_path / "t.cpp"
    src.write_text(source if source is not None else _source())
    exe = tmp_path / "t"
    libdir = os.path.join(ROOT, "dbgphmm_b200", "lib")
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir,
                           "-ldbgphmm_b200", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-lcudart", "-o", str(exe)], env=env)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    return p.returncode, p.stdout + p.stderr


def test_cpp_mirror_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    from dbgphmm_b200 import hmmv2 as H
    rc, out = _build_and_run(tmp_path)
    if H.device_count() == 0:
        assert rc == RC_NO_DEVICE, (rc, out)      # host-only part passed, then DBGPHMM_ERR_CUDA from the constructor
    else:
        assert rc == 0, (rc, out)


@pytest.mark.gpu
def test_cpp_mirror_reference_unit_tests_on_the_gpu(tmp_path):
    rc, out = _build_and_run(tmp_path)
    assert rc == 0 and "cpp mirror ok" in out, (rc, out)
