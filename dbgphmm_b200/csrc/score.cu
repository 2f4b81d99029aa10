// score.cu — the terms of MultiDbg::to_score beside the likelihood (multi_dbg/posterior.rs:225-277), host side only (no kernels):
// what the posterior sampler needs per candidate copy-number vector X once P(R|X) comes from the device.
//   n_euler_circuits : MultiDbg::n_euler_circuits (multi_dbg.rs:831-837) -> euler_circuit_count (graph/euler.rs:22-123):
//                      BEST theorem over the compact graph with copy numbers as edge multiplicities, in log space
//   genome_size      : MultiDbg::genome_size (multi_dbg.rs:1018-1028)
//   prior            : MultiDbg::to_prior (posterior.rs:225-231) = distribution::normal (distribution.rs:22-25)
// The reference takes the log-determinant from LAPACK (ndarray_linalg sln_det); here: LU with partial pivoting in f64.
#include <cmath>
#include <cstring>
#include <vector>
#include "model.h"
#include "../../include/dbgphmm_b200.h"

// log(n!) as the reference adds it up: ln(n) + ln(n-1) + ... + ln(1) (utils.rs:105-111)
static double log_factorial(uint64_t n) {
    double r = 0.0;
    for (uint64_t i = n; i >= 1; i--) r += std::log((double)i);
    return r;
}

// (sign, ln |det A|) of a dense n x n matrix (row-major, destroyed): what sln_det returns (euler.rs:56)
static void sln_det(std::vector<double>& a, size_t n, double* sign, double* ln) {
    double s = 1.0, l = 0.0;
    for (size_t c = 0; c < n; c++) {
        size_t piv = c;
        double best = std::fabs(a[c * n + c]);
        for (size_t r = c + 1; r < n; r++) { const double v = std::fabs(a[r * n + c]); if (v > best) { best = v; piv = r; } }
        if (best == 0.0) { *sign = 0.0; *ln = -INFINITY; return; }
        if (piv != c) { for (size_t k = c; k < n; k++) std::swap(a[c * n + k], a[piv * n + k]); s = -s; }
        const double d = a[c * n + c];
        if (d < 0) s = -s;
        l += std::log(std::fabs(d));
        for (size_t r = c + 1; r < n; r++) {
            const double f = a[r * n + c] / d;
            if (f == 0.0) continue;
            double* rr = &a[r * n]; const double* rc = &a[c * n];
            for (size_t k = c + 1; k < n; k++) rr[k] -= f * rc[k];
        }
    }
    *sign = s; *ln = l;
}

struct MultiGraph {   // DiGraph<(), usize> of euler.rs: parallel edges and self loops allowed
    uint32_t n = 0;
    std::vector<uint32_t> s, t;
    std::vector<uint64_t> w;
};

// euler_circuit_count_in_connected (euler.rs:22-84) over the nodes `nodes` (a strongly connected set) of g
static double count_in_connected(const MultiGraph& g, const std::vector<uint32_t>& nodes, std::vector<int64_t>& local) {
    const size_t n = nodes.size();
    if (n == 0) return 0.0;
    for (size_t i = 0; i < n; i++) local[nodes[i]] = (int64_t)i;
    std::vector<double> L(n * n, 0.0);
    std::vector<uint64_t> out(n, 0);
    for (size_t e = 0; e < g.s.size(); e++) {
        const int64_t i = local[g.s[e]], j = local[g.t[e]];
        if (i < 0 || j < 0) continue;
        out[i] += g.w[e];
        L[i * n + i] += (double)g.w[e];    // degree matrix: copies leaving node i
        L[i * n + j] -= (double)g.w[e];    // minus the adjacency matrix (a self loop cancels on the diagonal)
    }
    L[0] += 1.0;                            // "starting point is arbitrary" (euler.rs:52-54)
    double sign, ln;
    sln_det(L, n, &sign, &ln);
    double count = ln == -INFINITY ? 0.0 : sign * ln;     // (as the reference combines them, euler.rs:62-66)
    for (size_t i = 0; i < n; i++) if (out[i] > 0) count += log_factorial(out[i] - 1);
    for (size_t e = 0; e < g.s.size(); e++) if (local[g.s[e]] >= 0 && local[g.t[e]] >= 0) count -= log_factorial(g.w[e]);
    for (size_t i = 0; i < n; i++) local[nodes[i]] = -1;
    return count;
}

// strongly connected components (iterative Tarjan) of the nodes with keep[v] != 0
static void scc(const MultiGraph& g, const std::vector<uint8_t>& keep, std::vector<std::vector<uint32_t>>* comps) {
    const uint32_t n = g.n;
    std::vector<uint32_t> off(n + 1, 0), adj(g.s.size());
    for (size_t e = 0; e < g.s.size(); e++) off[g.s[e] + 1]++;
    for (uint32_t v = 0; v < n; v++) off[v + 1] += off[v];
    { std::vector<uint32_t> cur(off.begin(), off.end() - 1); for (size_t e = 0; e < g.s.size(); e++) adj[cur[g.s[e]]++] = g.t[e]; }
    const uint32_t NONE = 0xffffffffu;
    std::vector<uint32_t> index(n, NONE), low(n, 0), it(n, 0), stack, call;
    std::vector<uint8_t> on(n, 0);
    uint32_t next = 0;
    for (uint32_t root = 0; root < n; root++) {
        if (!keep[root] || index[root] != NONE) continue;
        call.push_back(root);
        while (!call.empty()) {
            const uint32_t v = call.back();
            if (index[v] == NONE) { index[v] = low[v] = next++; stack.push_back(v); on[v] = 1; it[v] = off[v]; }
            bool descended = false;
            while (it[v] < off[v + 1]) {
                const uint32_t u = adj[it[v]++];
                if (!keep[u]) continue;
                if (index[u] == NONE) { call.push_back(u); descended = true; break; }
                if (on[u]) low[v] = std::min(low[v], index[u]);
            }
            if (descended) continue;
            call.pop_back();
            if (!call.empty()) low[call.back()] = std::min(low[call.back()], low[v]);
            if (low[v] == index[v]) {
                comps->emplace_back();
                uint32_t u;
                do { u = stack.back(); stack.pop_back(); on[u] = 0; comps->back().push_back(u); } while (u != v);
            }
        }
    }
}

// euler_circuit_count (euler.rs:94-123).  The BEST theorem counts circuits of EULERIAN graphs; a multigraph whose copies do not
// balance at some node has none, and -inf is returned for it (the reference evaluates the same formula on whatever is left after
// `retain_nodes`, which is -inf on its own unbalanced test case, euler.rs:153-155; MultiDbg only ever passes balanced copy
// numbers, multi_dbg.rs:1041-1052).
static double euler_circuit_count(const MultiGraph& g0, bool allow_multiple_component) {
    MultiGraph g; g.n = g0.n;
    for (size_t e = 0; e < g0.s.size(); e++) if (g0.w[e] > 0) { g.s.push_back(g0.s[e]); g.t.push_back(g0.t[e]); g.w.push_back(g0.w[e]); }   // remove zero edges
    std::vector<long long> bal(g.n, 0);
    std::vector<uint8_t> keep(g.n, 0);
    for (size_t e = 0; e < g.s.size(); e++) { bal[g.s[e]] -= (long long)g.w[e]; bal[g.t[e]] += (long long)g.w[e]; keep[g.s[e]] = 1; }        // isolated nodes drop out
    uint32_t n = 0;
    for (uint32_t v = 0; v < g.n; v++) n += keep[v];
    if (n == 0) return -INFINITY;
    for (uint32_t v = 0; v < g.n; v++) if (bal[v] != 0) return -INFINITY;
    std::vector<std::vector<uint32_t>> comps;
    scc(g, keep, &comps);
    std::vector<int64_t> local(g.n, -1);
    if (!allow_multiple_component) {
        if (comps.size() > 1) return -INFINITY;
        return count_in_connected(g, comps[0], local);
    }
    double ret = 0.0;
    for (auto& c : comps) ret += count_in_connected(g, c, local);
    return ret;
}

extern "C" int dbgphmm_euler_circuit_count(uint32_t n_nodes, uint64_t n_edges, const uint32_t* edge_src, const uint32_t* edge_dst,
                                           const uint32_t* multiplicity, int allow_multiple_component, double* out_ln_count) try {
    if (!out_ln_count || (n_edges && (!edge_src || !edge_dst || !multiplicity))) { dbg_set_error("euler_circuit_count: bad argument"); return DBGPHMM_ERR_INVALID; }
    MultiGraph g; g.n = n_nodes;
    for (uint64_t e = 0; e < n_edges; e++) {
        if (edge_src[e] >= n_nodes || edge_dst[e] >= n_nodes) { dbg_set_error("euler_circuit_count: edge endpoint out of range"); return DBGPHMM_ERR_INVALID; }
        g.s.push_back(edge_src[e]); g.t.push_back(edge_dst[e]); g.w.push_back(multiplicity[e]);
    }
    *out_ln_count = euler_circuit_count(g, allow_multiple_component != 0);
    return DBGPHMM_OK;
} ABI_CATCH

extern "C" int dbgphmm_prior_normal(double x, double mu, double sigma, double* out_ln_p) try {
    if (!out_ln_p) { dbg_set_error("prior_normal: bad argument"); return DBGPHMM_ERR_INVALID; }
    const double s2 = sigma * sigma;   // distribution.rs:22-25
    *out_ln_p = (-0.5 * std::log(2.0 * M_PI * s2)) - ((x - mu) * (x - mu) / (2.0 * s2));
    return DBGPHMM_OK;
} ABI_CATCH
