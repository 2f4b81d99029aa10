// dbgphmm_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A CPU restatement of the hmmv2 profile-HMM hot path of ryought/dbgphmm, used
//   * by tests/ as the checker for the CUDA path,
//   * by __graft_entry__.smoke() as the checker,
//   * by bench.py's cpu_baseline / `--impl reference` leg as the timed CPU baseline
//     ("port": the Rust reference cannot be built here, no cargo/rustc).
// Nothing under dbgphmm_b200/ may import, link or execute this file.
//
// Parity status: PINNED against the reference's own known-answer tests
//   (forward.rs:576-669, backward.rs:577-652, freq.rs:434-610, posterior/test.rs:545-576),
//   see tests/test_oracle_golden.py.  Unpinned (third-party crate semantics not in
//   /root/reference): tie-breaking of sparsevec::to_top_k_indexes (we use: higher value
//   first, then earlier insertion position / lower node index), behaviour at the
//   400-entry capacity (we raise an error), density of mixed dense*sparse products
//   (we keep the sparse operand's index set), petgraph adjacency order (we follow
//   petgraph 0.6: most recently added edge first).
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src).  Arithmetic is log-space f64 exactly as prob.rs.
//
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

static const double NEG_INF = -std::numeric_limits<double>::infinity();
static const size_t MAX_ACTIVE_NODES = 400;  // hmmv2/table.rs:22
static const uint8_t NULL_BASE = 'n';         // common.rs:21

// ---------------------------------------------------------------- prob.rs
// prob.rs:181-197  impl Add for Prob
static inline double padd(double a, double b) {
    double x, y;
    if (a >= b) { x = a; y = b; } else { x = b; y = a; }
    if (y == NEG_INF) return x;
    if (x == y) return x + std::log(2.0);
    return x + std::log1p(std::exp(y - x));
}
// prob.rs:204-221  Mul / Div
static inline double pmul(double a, double b) { return a + b; }
static inline double pdiv(double a, double b) { return a - b; }

// ---------------------------------------------------------------- params.rs
struct Params {  // hmmv2/params.rs:16-66 ; all p_* are natural-log probabilities
    double p_mismatch, p_match, p_random, p_gap_open, p_gap_ext, p_end;
    double p_MM, p_IM, p_DM, p_MI, p_II, p_DI, p_MD, p_ID, p_DD;
    uint32_t n_active_nodes;
    uint32_t n_warmup;
    uint32_t warmup_threshold;
    uint32_t n_max_gaps;
    double active_node_max_ratio;
};

// hmmv2/params.rs:73-124  PHMMParams::new + uniform
static Params params_new(double p_mismatch, double p_gap_open, double p_gap_ext, double p_end,
                         uint32_t n_active_nodes, uint32_t n_warmup) {
    Params q;
    // Prob::from_prob(v) = ln(v); to_value() = exp(ln(v))  (prob.rs:58-72)
    double l_mis = std::log(p_mismatch), l_open = std::log(p_gap_open), l_ext = std::log(p_gap_ext),
           l_end = std::log(p_end);
    q.p_mismatch = l_mis; q.p_gap_open = l_open; q.p_gap_ext = l_ext; q.p_end = l_end;
    q.p_DD = l_ext; q.p_II = l_ext; q.p_MI = l_open; q.p_MD = l_open; q.p_ID = l_open; q.p_DI = l_open;
    q.p_MM = std::log(1.0 - 2.0 * std::exp(l_open) - std::exp(l_end));
    q.p_DM = std::log(1.0 - std::exp(l_open) - std::exp(l_ext) - std::exp(l_end));
    q.p_IM = std::log(1.0 - std::exp(l_open) - std::exp(l_ext) - std::exp(l_end));
    q.p_match = std::log(1.0 - std::exp(l_mis));
    q.p_random = std::log(0.25);
    q.n_active_nodes = n_active_nodes;
    q.active_node_max_ratio = 30.0;
    q.n_warmup = n_warmup;
    q.n_max_gaps = 4;
    q.warmup_threshold = MAX_ACTIVE_NODES / 2;  // params.rs:68-70
    return q;
}

// ---------------------------------------------------------------- sparsevec (re-creation)
// table.rs:28  NodeVec = SparseVec<Prob, NodeIndex, 400>.  Semantics inferred from call
// sites: absent sparse index reads the default; IndexMut on an absent sparse
// index appends (capacity 400); iter() yields stored entries in insertion order (dense: all).
struct NodeVec {
    bool dense = true;
    uint32_t n = 0;
    double dflt = NEG_INF;
    std::vector<double> dv;
    std::vector<uint32_t> ids;
    std::vector<double> vals;

    NodeVec() {}
    NodeVec(uint32_t n_, double d, bool is_dense) : dense(is_dense), n(n_), dflt(d) {
        if (dense) dv.assign(n, d);
    }
    size_t n_elements() const { return dense ? n : ids.size(); }
    double get(uint32_t i) const {
        if (dense) return dv[i];
        for (size_t j = 0; j < ids.size(); j++) if (ids[j] == i) return vals[j];
        return dflt;
    }
    double& ref(uint32_t i) {
        if (dense) return dv[i];
        for (size_t j = 0; j < ids.size(); j++) if (ids[j] == i) return vals[j];
        if (ids.size() >= MAX_ACTIVE_NODES) throw std::runtime_error("SparseVec: insufficient capacity");
        ids.push_back(i); vals.push_back(dflt);
        return vals.back();
    }
    template <class F> void for_each(F f) const {
        if (dense) { for (uint32_t i = 0; i < n; i++) f(i, dv[i]); }
        else { for (size_t j = 0; j < ids.size(); j++) f(ids[j], vals[j]); }
    }
    // `a += &b`  (forward.rs:448): for every stored entry of b, a[idx] += val
    void add_assign(const NodeVec& o) {
        o.for_each([&](uint32_t i, double v) { double& r = ref(i); r = padd(r, v); });
    }
    // descending by value; ties: earlier stored position first (UNPINNED, see header)
    std::vector<std::pair<uint32_t, double>> sorted_desc(size_t k) const {
        std::vector<std::pair<uint32_t, double>> e;
        e.reserve(n_elements());
        for_each([&](uint32_t i, double v) {
            if (v != v) throw std::runtime_error("NaN in Prob ordering (prob.rs:296-300 panics)");
            e.push_back({i, v});
        });
        // positions are implicit in e's order -> use stable ordering on value only
        std::vector<uint32_t> pos(e.size());
        for (size_t j = 0; j < pos.size(); j++) pos[j] = (uint32_t)j;
        auto cmp = [&](uint32_t a, uint32_t b) {
            if (e[a].second != e[b].second) return e[a].second > e[b].second;
            return a < b;
        };
        if (k < pos.size()) {
            std::nth_element(pos.begin(), pos.begin() + k, pos.end(), cmp);
            pos.resize(k);
        }
        std::sort(pos.begin(), pos.end(), cmp);
        std::vector<std::pair<uint32_t, double>> out;
        out.reserve(pos.size());
        for (uint32_t p : pos) out.push_back(e[p]);
        return out;
    }
    std::vector<uint32_t> to_top_k_indexes(size_t k) const {  // table.rs:121,128
        std::vector<uint32_t> r;
        for (auto& p : sorted_desc(std::min(k, MAX_ACTIVE_NODES))) r.push_back(p.first);
        return r;
    }
};

// ---------------------------------------------------------------- table.rs
struct Table {  // table.rs:42-73
    NodeVec m, i, d;
    double mb = NEG_INF, ib = NEG_INF, e = NEG_INF;
    Table() {}
    Table(bool dense, uint32_t n, double vm, double vi, double vd, double mb_, double ib_, double e_)
        : m(n, vm, dense), i(n, vi, dense), d(n, vd, dense), mb(mb_), ib(ib_), e(e_) {}
    static Table zero(bool dense, uint32_t n) { return Table(dense, n, NEG_INF, NEG_INF, NEG_INF, NEG_INF, NEG_INF, NEG_INF); }
    bool is_dense() const { return m.dense; }
    uint32_t n_nodes() const { return m.n; }
    // table.rs:199-211
    NodeVec to_nodevec() const {
        NodeVec v(n_nodes(), NEG_INF, is_dense());
        m.for_each([&](uint32_t k, double p) { double& r = v.ref(k); r = padd(r, p); });
        i.for_each([&](uint32_t k, double p) { double& r = v.ref(k); r = padd(r, p); });
        d.for_each([&](uint32_t k, double p) { double& r = v.ref(k); r = padd(r, p); });
        return v;
    }
    std::vector<uint32_t> top_nodes(size_t k) const { return to_nodevec().to_top_k_indexes(k); }  // table.rs:127
    // table.rs:134-149
    std::vector<uint32_t> top_nodes_by_score_ratio(double max_ratio) const {
        std::vector<uint32_t> ret;
        auto v = to_nodevec().sorted_desc(MAX_ACTIVE_NODES);
        if (!v.empty()) {
            double p0 = v[0].second;
            for (auto& pr : v) if (p0 - pr.second < max_ratio) ret.push_back(pr.first);
        }
        return ret;
    }
    // table.rs:117-123
    std::vector<uint32_t> filled_nodes() const { return to_nodevec().to_top_k_indexes(m.n_elements()); }
};

enum Kind { FORWARD = 0, BACKWARD = 1 };
struct Tables {  // table.rs:365-435
    Table init_table;
    std::vector<Table> tables;
    Kind kind;
    size_t n_emissions() const { return tables.size(); }
    const Table& table_merged(size_t mi) const {  // table.rs:414-434
        if (kind == FORWARD) return mi == 0 ? init_table : tables[mi - 1];
        return mi >= tables.size() ? init_table : tables[mi];
    }
    double full_prob() const {  // table.rs:395-401
        if (tables.empty()) throw std::runtime_error("empty PHMMTables");
        return kind == FORWARD ? tables.back().e : tables.front().mb;
    }
};

// ---------------------------------------------------------------- common.rs : model
struct Model {  // hmmv2/common.rs:61-67
    Params param;
    uint32_t n_nodes = 0;
    std::vector<uint8_t> emission;
    std::vector<double> init;  // log
    std::vector<uint32_t> esrc, edst;
    std::vector<double> etrans;  // log
    // petgraph 0.6 adjacency: per node, edges in most-recently-added-first order
    std::vector<std::vector<uint32_t>> out_e, in_e;

    void build_adj() {
        out_e.assign(n_nodes, {}); in_e.assign(n_nodes, {});
        for (size_t e = esrc.size(); e-- > 0;) { out_e[esrc[e]].push_back((uint32_t)e); in_e[edst[e]].push_back((uint32_t)e); }
    }
    // common.rs:168-174
    double p_match_emit(uint32_t k, uint8_t x) const { return emission[k] == x ? param.p_match : param.p_mismatch; }
    double p_ins_emit() const { return param.p_random; }  // common.rs:180-182

    // active_nodes.rs:15-56 ; itertools unique = first occurrence, then take(400)
    std::vector<uint32_t> expand(const std::vector<uint32_t>& nodes, bool children, bool and_us) const {
        std::vector<uint32_t> r;
        std::vector<uint8_t> seen_small;
        auto push = [&](uint32_t v) {
            if (r.size() >= MAX_ACTIVE_NODES) return;
            for (uint32_t u : r) if (u == v) return;
            r.push_back(v);
        };
        if (and_us) for (uint32_t v : nodes) push(v);
        for (uint32_t v : nodes) {
            const auto& es = children ? out_e[v] : in_e[v];
            for (uint32_t e : es) push(children ? edst[e] : esrc[e]);
        }
        return r;
    }
    std::vector<uint32_t> to_childs(const std::vector<uint32_t>& n) const { return expand(n, true, false); }
    std::vector<uint32_t> to_childs_and_us(const std::vector<uint32_t>& n) const { return expand(n, true, true); }
    std::vector<uint32_t> to_parents(const std::vector<uint32_t>& n) const { return expand(n, false, false); }
    std::vector<uint32_t> to_parents_and_us(const std::vector<uint32_t>& n) const { return expand(n, false, true); }
    std::vector<uint32_t> to_all_nodes() const { std::vector<uint32_t> r(n_nodes); for (uint32_t i = 0; i < n_nodes; i++) r[i] = i; return r; }

    // ------------------------------------------------------------ forward.rs
    Table f_init() const { return Table(true, n_nodes, NEG_INF, NEG_INF, NEG_INF, 0.0, NEG_INF, NEG_INF); }  // forward.rs:255-266

    // forward.rs:337-359
    void fm(Table& t0, const Table& t1, uint8_t x, const std::vector<uint32_t>& nodes) const {
        const Params& p = param;
        for (uint32_t k : nodes) {
            double p_emit = p_match_emit(k, x);
            double p_init = init[k];
            double from_normal = NEG_INF;
            for (uint32_t e : in_e[k]) {
                uint32_t l = esrc[e];
                double v = pmul(etrans[e], padd(padd(pmul(p.p_MM, t1.m.get(l)), pmul(p.p_IM, t1.i.get(l))), pmul(p.p_DM, t1.d.get(l))));
                from_normal = padd(from_normal, v);
            }
            double from_begin = pmul(p_init, padd(pmul(p.p_MM, t1.mb), pmul(p.p_IM, t1.ib)));
            t0.m.ref(k) = pmul(p_emit, padd(from_normal, from_begin));
        }
    }
    // forward.rs:378-388
    void fi(Table& t0, const Table& t1, const std::vector<uint32_t>& nodes) const {
        const Params& p = param;
        for (uint32_t k : nodes) {
            double from_me = padd(padd(pmul(p.p_MI, t1.m.get(k)), pmul(p.p_II, t1.i.get(k))), pmul(p.p_DI, t1.d.get(k)));
            t0.i.ref(k) = pmul(p_ins_emit(), from_me);
        }
    }
    // forward.rs:480-501
    Table fd0(const Table& t0, const std::vector<uint32_t>& nodes, bool is_dense) const {
        const Params& p = param;
        Table r = Table::zero(is_dense, n_nodes);
        for (uint32_t k : nodes) {
            double from_normal = NEG_INF;
            for (uint32_t e : in_e[k]) {
                uint32_t l = esrc[e];
                double v = pmul(etrans[e], padd(pmul(p.p_MD, t0.m.get(l)), pmul(p.p_ID, t0.i.get(l))));
                from_normal = padd(from_normal, v);
            }
            double from_begin = pmul(init[k], padd(pmul(p.p_MD, t0.mb), pmul(p.p_ID, t0.ib)));
            r.d.ref(k) = padd(from_normal, from_begin);
        }
        return r;
    }
    // forward.rs:510-524
    Table fdt(const Table& fdt1, const std::vector<uint32_t>& nodes, bool is_dense) const {
        const Params& p = param;
        Table r = Table::zero(is_dense, n_nodes);
        for (uint32_t k : nodes) {
            double s = NEG_INF;
            for (uint32_t e : in_e[k]) s = padd(s, pmul(etrans[e], pmul(p.p_DD, fdt1.d.get(esrc[e]))));
            r.d.ref(k) = s;
        }
        return r;
    }
    // forward.rs:423-466
    void fd(Table& t0, const std::vector<uint32_t>& nodes, bool is_adaptive) const {
        bool is_dense = t0.is_dense();
        std::vector<uint32_t> act;
        if (is_adaptive) act = to_childs(nodes);
        Table fdt0 = fd0(t0, is_adaptive ? act : nodes, is_dense);
        t0.d.add_assign(fdt0.d);
        for (uint32_t t = 0; t < param.n_max_gaps; t++) {
            if (is_adaptive) act = to_childs(act);
            fdt0 = fdt(fdt0, is_adaptive ? act : nodes, is_dense);
            t0.d.add_assign(fdt0.d);
        }
    }
    // forward.rs:276-306  order: fm, fi, fmb, fib, fd, fe
    Table f_step(uint8_t x, const Table& prev, const std::vector<uint32_t>& nodes, bool is_dense, bool is_adaptive) const {
        Table t(is_dense, n_nodes, NEG_INF, NEG_INF, NEG_INF, NEG_INF, NEG_INF, NEG_INF);
        fm(t, prev, x, nodes);
        fi(t, prev, nodes);
        t.mb = NEG_INF;                                                                   // forward.rs:531-533
        t.ib = pmul(p_ins_emit(), padd(pmul(param.p_MI, prev.mb), pmul(param.p_II, prev.ib)));  // forward.rs:541-545
        fd(t, nodes, is_adaptive);
        double s = NEG_INF;                                                               // forward.rs:554-558
        for (uint32_t k : nodes) s = padd(s, padd(padd(t.m.get(k), t.i.get(k)), t.d.get(k)));
        t.e = pmul(param.p_end, s);
        return t;
    }
    // forward.rs:24-45
    Tables forward(const uint8_t* x, size_t n) const {
        Tables r; r.kind = FORWARD; r.init_table = f_init();
        auto all = to_all_nodes();
        for (size_t i = 0; i < n; i++) {
            const Table& prev = i == 0 ? r.init_table : r.tables.back();
            Table t = f_step(x[i], prev, all, true, false);
            r.tables.push_back(std::move(t));
        }
        return r;
    }
    // forward.rs:51-75
    Tables forward_with_mapping(const uint8_t* x, size_t n, const std::vector<std::vector<uint32_t>>& map_nodes) const {
        Tables r; r.kind = FORWARD; r.init_table = f_init();
        for (size_t i = 0; i < n; i++) {
            const Table& prev = i == 0 ? r.init_table : r.tables.back();
            Table t = f_step(x[i], prev, map_nodes[i], false, false);
            r.tables.push_back(std::move(t));
        }
        return r;
    }
    // forward.rs:79-89
    double forward_with_mapping_score_only(const uint8_t* x, size_t n, const std::vector<std::vector<uint32_t>>& map_nodes) const {
        Table t = f_init();
        for (size_t i = 0; i < n; i++) t = f_step(x[i], t, map_nodes[i], false, false);
        return t.e;
    }
    // decision logic shared by forward.rs:93-154 and :158-206
    Table sparse_step(size_t i, uint8_t x, const Table& prev, bool use_max_ratio, const std::vector<uint32_t>& all) const {
        std::vector<uint32_t> top = use_max_ratio ? prev.top_nodes_by_score_ratio(param.active_node_max_ratio)
                                                  : prev.top_nodes(param.n_active_nodes);
        bool use_dense;
        if (use_max_ratio) {
            if (prev.is_dense()) {
                if (i == 0) use_dense = true;
                else if (i < param.n_warmup) use_dense = top.size() > param.warmup_threshold;
                else use_dense = false;
            } else use_dense = false;
        } else use_dense = i < param.n_warmup;
        if (use_dense) return f_step(x, prev, all, true, false);
        auto active = to_childs_and_us(top);
        return f_step(x, prev, active, false, true);
    }
    Tables forward_sparse(const uint8_t* x, size_t n, bool use_max_ratio) const {
        Tables r; r.kind = FORWARD; r.init_table = f_init();
        auto all = to_all_nodes();
        for (size_t i = 0; i < n; i++) {
            const Table& prev = i == 0 ? r.init_table : r.tables.back();
            Table t = sparse_step(i, x[i], prev, use_max_ratio, all);
            r.tables.push_back(std::move(t));
        }
        return r;
    }
    double forward_sparse_score_only(const uint8_t* x, size_t n, bool use_max_ratio) const {
        Table t = f_init();
        auto all = to_all_nodes();
        for (size_t i = 0; i < n; i++) t = sparse_step(i, x[i], t, use_max_ratio, all);
        return t.e;
    }

    // ------------------------------------------------------------ backward.rs
    Table b_init() const {  // backward.rs:197-211
        double pe = param.p_end;
        return Table(true, n_nodes, pe, pe, pe, NEG_INF, NEG_INF, NEG_INF);
    }
    // backward.rs:354-377
    Table bd0(const Table& t1, uint8_t x, const std::vector<uint32_t>& nodes, bool is_dense) const {
        const Params& p = param;
        Table r = Table::zero(is_dense, n_nodes);
        for (uint32_t k : nodes) {
            double to_match = NEG_INF;
            for (uint32_t e : out_e[k]) {
                uint32_t l = edst[e];
                // p_trans * p_DM * p_emit * t0.m[l]  (left-assoc)
                double v = pmul(pmul(pmul(etrans[e], p.p_DM), p_match_emit(l, x)), t1.m.get(l));
                to_match = padd(to_match, v);
            }
            double to_ins = pmul(pmul(p.p_DI, p_ins_emit()), t1.i.get(k));
            r.d.ref(k) = padd(to_match, to_ins);
        }
        return r;
    }
    // backward.rs:387-404
    Table bdt(const Table& bdt1, const std::vector<uint32_t>& nodes, bool is_dense) const {
        const Params& p = param;
        Table r = Table::zero(is_dense, n_nodes);
        for (uint32_t k : nodes) {
            double s = NEG_INF;
            for (uint32_t e : out_e[k]) s = padd(s, pmul(pmul(etrans[e], p.p_DD), bdt1.d.get(edst[e])));
            r.d.ref(k) = s;
        }
        return r;
    }
    // backward.rs:299-343
    void bd(Table& t0, const Table& t1, uint8_t x, const std::vector<uint32_t>& nodes, bool is_adaptive) const {
        bool is_dense = t0.is_dense();
        std::vector<uint32_t> act;
        if (is_adaptive) act = to_parents_and_us(nodes);
        Table b0 = bd0(t1, x, is_adaptive ? act : nodes, is_dense);
        t0.d.add_assign(b0.d);
        for (uint32_t t = 0; t < param.n_max_gaps; t++) {
            if (is_adaptive) act = to_parents_and_us(act);
            b0 = bdt(b0, is_adaptive ? act : nodes, is_dense);
            t0.d.add_assign(b0.d);
        }
    }
    // backward.rs:423-444 (bm) and :462-483 (bi)
    void bmi(Table& t0, const Table& t1, uint8_t x, const std::vector<uint32_t>& nodes) const {
        const Params& p = param;
        for (uint32_t k : nodes) {
            double s = NEG_INF;
            for (uint32_t e : out_e[k]) {
                uint32_t l = edst[e];
                double pe = p_match_emit(l, x);
                double v = pmul(etrans[e], padd(pmul(pmul(p.p_MM, pe), t1.m.get(l)), pmul(p.p_MD, t0.d.get(l))));
                s = padd(s, v);
            }
            double to_ins = pmul(pmul(p.p_MI, p_ins_emit()), t1.i.get(k));
            t0.m.ref(k) = padd(s, to_ins);
        }
        for (uint32_t k : nodes) {
            double s = NEG_INF;
            for (uint32_t e : out_e[k]) {
                uint32_t l = edst[e];
                double pe = p_match_emit(l, x);
                double v = pmul(etrans[e], padd(pmul(pmul(p.p_IM, pe), t1.m.get(l)), pmul(p.p_ID, t0.d.get(l))));
                s = padd(s, v);
            }
            double to_ins = pmul(pmul(p.p_II, p_ins_emit()), t1.i.get(k));
            t0.i.ref(k) = padd(s, to_ins);
        }
    }
    // backward.rs:499-555 (bib then bmb; each a fold over `nodes`)
    void bbegin(Table& t0, const Table& t1, uint8_t x, const std::vector<uint32_t>& nodes) const {
        const Params& p = param;
        double si = NEG_INF, sm = NEG_INF;
        for (uint32_t l : nodes) {
            double pe = p_match_emit(l, x);
            si = padd(si, pmul(init[l], padd(pmul(pmul(p.p_IM, pe), t1.m.get(l)), pmul(p.p_ID, t0.d.get(l)))));
        }
        t0.ib = padd(si, pmul(pmul(p.p_II, p_ins_emit()), t1.ib));
        for (uint32_t l : nodes) {
            double pe = p_match_emit(l, x);
            sm = padd(sm, pmul(init[l], padd(pmul(pmul(p.p_MM, pe), t1.m.get(l)), pmul(p.p_MD, t0.d.get(l)))));
        }
        t0.mb = padd(sm, pmul(pmul(p.p_MI, p_ins_emit()), t1.ib));
    }
    // backward.rs:216-261  order: bd, be, (bm, bi, bib, bmb)
    Table b_step(uint8_t x, const Table& prev, const std::vector<uint32_t>& nodes, bool is_dense, bool is_adaptive) const {
        Table t(is_dense, n_nodes, NEG_INF, NEG_INF, NEG_INF, NEG_INF, NEG_INF, NEG_INF);
        bd(t, prev, x, nodes, is_adaptive);
        t.e = NEG_INF;
        if (is_adaptive) {
            auto pu = to_parents_and_us(nodes);
            bmi(t, prev, x, pu);
            bbegin(t, prev, x, pu);
        } else {
            bmi(t, prev, x, nodes);
            bbegin(t, prev, x, nodes);
        }
        return t;
    }
    // backward.rs:24-53
    Tables backward(const uint8_t* x, size_t n) const {
        Tables r; r.kind = BACKWARD; r.init_table = b_init();
        auto all = to_all_nodes();
        for (size_t i = n; i-- > 0;) {
            const Table& prev = (i == n - 1) ? r.init_table : r.tables.back();
            Table t = b_step(x[i], prev, all, true, false);
            r.tables.push_back(std::move(t));
        }
        std::reverse(r.tables.begin(), r.tables.end());
        return r;
    }
    // backward.rs:59-93
    Tables backward_with_mapping(const uint8_t* x, size_t n, const std::vector<std::vector<uint32_t>>& map_nodes) const {
        Tables r; r.kind = BACKWARD; r.init_table = b_init();
        for (size_t i = n; i-- > 0;) {
            const Table& prev = (i == n - 1) ? r.init_table : r.tables.back();
            Table t = b_step(x[i], prev, map_nodes[i], false, false);
            r.tables.push_back(std::move(t));
        }
        std::reverse(r.tables.begin(), r.tables.end());
        return r;
    }
    // backward.rs:101-142
    Tables backward_by_forward(const uint8_t* x, size_t n, const Tables& fwd) const {
        Tables r; r.kind = BACKWARD; r.init_table = b_init();
        auto all = to_all_nodes();
        for (size_t i = n; i-- > 0;) {
            const Table& prev = (i == n - 1) ? r.init_table : r.tables.back();
            Table t;
            if (i == 0 || fwd.tables[i - 1].is_dense()) t = b_step(x[i], prev, all, true, false);
            else { auto act = fwd.tables[i - 1].filled_nodes(); t = b_step(x[i], prev, act, false, false); }
            r.tables.push_back(std::move(t));
        }
        std::reverse(r.tables.begin(), r.tables.end());
        return r;
    }
    // backward.rs:146-185
    Tables backward_sparse(const uint8_t* x, size_t n) const {
        Tables r; r.kind = BACKWARD; r.init_table = b_init();
        auto all = to_all_nodes();
        for (size_t i = n; i-- > 0;) {
            Table t;
            if ((n - i - 1) < param.n_warmup) {
                const Table& prev = (i == n - 1) ? r.init_table : r.tables.back();
                t = b_step(x[i], prev, all, true, false);
            } else {
                const Table& prev = r.tables.back();
                auto act = prev.top_nodes(param.n_active_nodes);
                t = b_step(x[i], prev, act, false, true);
            }
            r.tables.push_back(std::move(t));
        }
        std::reverse(r.tables.begin(), r.tables.end());
        return r;
    }
};

// ---------------------------------------------------------------- PHMMOutput (table.rs:450-517, freq.rs:198-255, hint.rs:120-142)
// &a * &b on NodeVecs (table.rs:318-331): index set of the sparse operand (lhs if both sparse / both dense)
static NodeVec nv_mul(const NodeVec& a, const NodeVec& b) {
    if (a.dense && b.dense) {
        NodeVec r(a.n, NEG_INF, true);
        for (uint32_t i = 0; i < a.n; i++) r.dv[i] = pmul(a.dv[i], b.dv[i]);
        return r;
    }
    const NodeVec& s = a.dense ? b : a;
    const NodeVec& o = a.dense ? a : b;
    NodeVec r(a.n, NEG_INF, false);
    for (size_t j = 0; j < s.ids.size(); j++) { r.ids.push_back(s.ids[j]); r.vals.push_back(pmul(s.vals[j], o.get(s.ids[j]))); }
    return r;
}
static void nv_div(NodeVec& a, double p) {
    if (a.dense) for (auto& v : a.dv) v = pdiv(v, p); else for (auto& v : a.vals) v = pdiv(v, p);
    a.dflt = pdiv(a.dflt, p);
}
struct Output {
    Tables f, b;
    size_t n() const { return f.n_emissions(); }
    // table.rs:500-505
    Table to_emit_probs(size_t mi) const {
        double p = f.full_prob();
        const Table& ft = f.table_merged(mi);
        const Table& bt = b.table_merged(mi);
        Table r;
        r.m = nv_mul(ft.m, bt.m); r.i = nv_mul(ft.i, bt.i); r.d = nv_mul(ft.d, bt.d);
        nv_div(r.m, p); nv_div(r.i, p); nv_div(r.d, p);
        r.mb = pdiv(pmul(ft.mb, bt.mb), p); r.ib = pdiv(pmul(ft.ib, bt.ib), p); r.e = pdiv(pmul(ft.e, bt.e), p);
        return r;
    }
    // freq.rs:237-255 : sum_{i=0..n} emit_probs(i), merge m+i+d, exp.  Accumulated densely (see header).
    std::vector<double> to_node_freqs() const {
        uint32_t N = f.init_table.n_nodes();
        NodeVec am(N, NEG_INF, true), ai(N, NEG_INF, true), ad(N, NEG_INF, true);
        for (size_t mi = 0; mi <= n(); mi++) {
            Table t = to_emit_probs(mi);
            // entries that are NaN (-inf - -inf when P==0) propagate like the reference
            am.add_assign(t.m); ai.add_assign(t.i); ad.add_assign(t.d);
        }
        std::vector<double> fr(N);
        for (uint32_t k = 0; k < N; k++) fr[k] = std::exp(padd(padd(padd(NEG_INF, am.dv[k]), ai.dv[k]), ad.dv[k]));
        return fr;
    }
    // hint.rs:124-142 ; rows i=1..n
    void to_mapping(bool by_ratio, size_t n_active, double max_ratio,
                    std::vector<std::vector<uint32_t>>& nodes, std::vector<std::vector<double>>& probs) const {
        nodes.clear(); probs.clear();
        for (size_t mi = 1; mi <= n(); mi++) {
            Table t = to_emit_probs(mi);
            NodeVec v = t.to_nodevec();
            std::vector<uint32_t> ids = by_ratio ? t.top_nodes_by_score_ratio(max_ratio) : v.to_top_k_indexes(n_active);
            std::vector<double> ps;
            for (uint32_t id : ids) ps.push_back(v.get(id));
            nodes.push_back(ids); probs.push_back(ps);
        }
    }
};

// ---------------------------------------------------------------- graph/seq_graph.rs:110-273
// copy numbers -> init / trans (log).  mode: 0 = to_phmm (min_copy_num 0), 1 = to_non_zero_phmm
// (min_copy_num 1), 2 = to_uniform_phmm.  edge_copy_num[e] < 0 means None (the branch MultiDbg uses).
static void seqgraph_to_phmm(uint32_t N, uint32_t E, const uint32_t* src, const uint32_t* dst, const uint8_t* base,
                             const int64_t* node_cn, const int64_t* edge_cn, int mode, double* log_init, double* log_trans) {
    auto emittable = [&](uint32_t v) { return base[v] != NULL_BASE; };
    if (mode == 2) {
        size_t n_emit = 0;
        for (uint32_t v = 0; v < N; v++) n_emit += emittable(v);
        for (uint32_t v = 0; v < N; v++) log_init[v] = emittable(v) ? pdiv(0.0, std::log((double)n_emit)) : NEG_INF;
        std::vector<size_t> n_child(N, 0);
        for (uint32_t e = 0; e < E; e++) if (emittable(dst[e])) n_child[src[e]]++;
        for (uint32_t e = 0; e < E; e++) log_trans[e] = emittable(dst[e]) ? pdiv(0.0, std::log((double)n_child[src[e]])) : NEG_INF;
        return;
    }
    int64_t minc = mode == 1 ? 1 : 0;
    auto cn = [&](uint32_t v) { return std::max(node_cn[v], minc); };
    int64_t total = 0;
    for (uint32_t v = 0; v < N; v++) if (emittable(v)) total += cn(v);
    for (uint32_t v = 0; v < N; v++)
        log_init[v] = emittable(v) ? pdiv(std::log((double)cn(v)), std::log((double)total)) : NEG_INF;  // seq_graph.rs:166-171
    std::vector<int64_t> child_total(N, 0);
    for (uint32_t e = 0; e < E; e++) if (emittable(dst[e])) child_total[src[e]] += cn(dst[e]);
    for (uint32_t e = 0; e < E; e++) {
        uint32_t parent = src[e], child = dst[e];
        if (edge_cn && edge_cn[e] >= 0) {  // seq_graph.rs:185-196
            int64_t c = edge_cn[e];
            log_trans[e] = (emittable(child) && c > 0) ? std::log((double)c / (double)node_cn[parent]) : NEG_INF;
        } else {  // seq_graph.rs:198-209
            int64_t tot = child_total[parent];
            log_trans[e] = (emittable(child) && tot > 0) ? std::log((double)cn(child) / (double)tot) : NEG_INF;
        }
    }
}

struct Mappings {  // hint.rs:27-30,150-152 ; CSR over reads -> bases -> (node, logp)
    std::vector<std::vector<std::vector<uint32_t>>> nodes;
    std::vector<std::vector<std::vector<double>>> probs;
};

}  // namespace orc

// ===================================================================== C API (ctypes)
using namespace orc;
static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(ret) } catch (const std::exception& ex) { g_err = ex.what(); return ret; }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void orc_params_uniform(double p, Params* out) { *out = params_new(p, p, p, 0.00001, 40, 50); }  // params.rs:116-124
void orc_params_new(double p_mismatch, double p_gap_open, double p_gap_ext, double p_end, uint32_t n_active, uint32_t n_warmup, Params* out) {
    *out = params_new(p_mismatch, p_gap_open, p_gap_ext, p_end, n_active, n_warmup);
}
double orc_padd(double a, double b) { return padd(a, b); }

void orc_seqgraph_to_phmm(uint32_t N, uint32_t E, const uint32_t* src, const uint32_t* dst, const uint8_t* base,
                          const int64_t* node_cn, const int64_t* edge_cn, int mode, double* log_init, double* log_trans) {
    seqgraph_to_phmm(N, E, src, dst, base, node_cn, edge_cn, mode, log_init, log_trans);
}

Model* orc_model_create(uint32_t n_nodes, uint32_t n_edges, const uint32_t* src, const uint32_t* dst, const uint8_t* emission,
                        const double* log_init, const double* log_trans, const Params* p) {
    Model* m = new Model();
    m->param = *p; m->n_nodes = n_nodes;
    m->emission.assign(emission, emission + n_nodes);
    m->init.assign(log_init, log_init + n_nodes);
    m->esrc.assign(src, src + n_edges); m->edst.assign(dst, dst + n_edges);
    m->etrans.assign(log_trans, log_trans + n_edges);
    m->build_adj();
    return m;
}
void orc_model_set_probs(Model* m, const double* log_init, const double* log_trans) {
    m->init.assign(log_init, log_init + m->n_nodes);
    m->etrans.assign(log_trans, log_trans + m->etrans.size());
}
void orc_model_set_params(Model* m, const Params* p) { m->param = *p; }
void orc_model_destroy(Model* m) { delete m; }

static std::vector<std::vector<uint32_t>> unpack_mapping(size_t n, const uint64_t* row_off, const uint32_t* nodes) {
    std::vector<std::vector<uint32_t>> r(n);
    for (size_t i = 0; i < n; i++) r[i].assign(nodes + row_off[i], nodes + row_off[i + 1]);
    return r;
}

// kind: 0 forward (dense), 1 forward_sparse(use_max_ratio=false), 2 forward_sparse(true), 3 forward_with_mapping
Tables* orc_forward(const Model* m, const uint8_t* x, uint64_t n, int kind, const uint64_t* map_row_off, const uint32_t* map_nodes) {
    ORC_TRY
    Tables* t = new Tables();
    if (kind == 0) *t = m->forward(x, n);
    else if (kind == 1) *t = m->forward_sparse(x, n, false);
    else if (kind == 2) *t = m->forward_sparse(x, n, true);
    else *t = m->forward_with_mapping(x, n, unpack_mapping(n, map_row_off, map_nodes));
    return t;
    ORC_CATCH(nullptr)
}
// kind: 0 backward (dense), 1 backward_sparse, 2 backward_with_mapping, 3 backward_by_forward(fwd)
Tables* orc_backward(const Model* m, const uint8_t* x, uint64_t n, int kind, const uint64_t* map_row_off, const uint32_t* map_nodes, const Tables* fwd) {
    ORC_TRY
    Tables* t = new Tables();
    if (kind == 0) *t = m->backward(x, n);
    else if (kind == 1) *t = m->backward_sparse(x, n);
    else if (kind == 2) *t = m->backward_with_mapping(x, n, unpack_mapping(n, map_row_off, map_nodes));
    else *t = m->backward_by_forward(x, n, *fwd);
    return t;
    ORC_CATCH(nullptr)
}
void orc_tables_destroy(Tables* t) { delete t; }
uint64_t orc_tables_len(const Tables* t) { return t->tables.size(); }
double orc_tables_full_prob(const Tables* t) { return t->full_prob(); }
// row = -1 -> init_table
static const Table& row_of(const Tables* t, int64_t row) { return row < 0 ? t->init_table : t->tables[row]; }
// info[0]=is_dense, info[1]=n entries of m (== of i), info[2]=n entries of d ; scalars = {mb, ib, e}
void orc_tables_row_info(const Tables* t, int64_t row, uint64_t* info, double* scalars) {
    const Table& r = row_of(t, row);
    info[0] = r.is_dense(); info[1] = r.m.n_elements(); info[2] = r.d.n_elements();
    scalars[0] = r.mb; scalars[1] = r.ib; scalars[2] = r.e;
}
// dense rows: m,i,d each n_nodes, ids unused.  sparse: ids_mi[n_m], m[n_m], i[n_m], ids_d[n_d], d[n_d]
void orc_tables_row_export(const Tables* t, int64_t row, uint32_t* ids_mi, double* m, double* i, uint32_t* ids_d, double* d) {
    const Table& r = row_of(t, row);
    if (r.is_dense()) {
        std::memcpy(m, r.m.dv.data(), 8 * r.m.n); std::memcpy(i, r.i.dv.data(), 8 * r.m.n); std::memcpy(d, r.d.dv.data(), 8 * r.m.n);
    } else {
        for (size_t j = 0; j < r.m.ids.size(); j++) { ids_mi[j] = r.m.ids[j]; m[j] = r.m.vals[j]; i[j] = r.i.get(r.m.ids[j]); }
        for (size_t j = 0; j < r.d.ids.size(); j++) { ids_d[j] = r.d.ids[j]; d[j] = r.d.vals[j]; }
    }
}
// top_nodes(k) (by_ratio=0) or top_nodes_by_score_ratio(ratio) of a row; returns count
uint64_t orc_tables_row_top_nodes(const Tables* t, int64_t row, int by_ratio, uint64_t k, double ratio, uint32_t* out) {
    const Table& r = row_of(t, row);
    auto v = by_ratio ? r.top_nodes_by_score_ratio(ratio) : r.top_nodes(k);
    for (size_t j = 0; j < v.size(); j++) out[j] = v[j];
    return v.size();
}

// PHMMOutput::to_node_freqs (freq.rs:245-255) for one read
int orc_output_node_freqs(const Tables* f, const Tables* b, double* freqs) {
    ORC_TRY
    Output o{*f, *b};
    auto fr = o.to_node_freqs();
    std::memcpy(freqs, fr.data(), 8 * fr.size());
    return 0;
    ORC_CATCH(1)
}
// PHMMOutput::to_edge_and_init_freqs (freq.rs:276-298) over to_trans_and_init_probs (freq.rs:332-389) for one read:
// edge[e] = sum_{i=0..n} (mm + md + im + id + dm + dd of edge e at i).to_value(), init[v] likewise for the Begin -> v transitions.
int orc_output_edge_init_freqs(const Model* mdl, const Tables* f, const Tables* b, const uint8_t* x, uint64_t n, double* edge, double* init) {
    ORC_TRY
    const Params& pa = mdl->param;
    const size_t E = mdl->esrc.size(), N = mdl->n_nodes;
    if (n != f->n_emissions() || n != b->n_emissions()) throw std::runtime_error("emissions / tables length mismatch (freq.rs:281-282)");
    std::fill(edge, edge + E, 0.0); std::fill(init, init + N, 0.0);
    const double p = f->full_prob();
    for (size_t i = 0; i <= n; i++) {
        const Table& fi0 = f->table_merged(i);
        const Table& bi2 = b->table_merged(i + 1);
        const Table& bi1 = b->table_merged(i);
        for (size_t e = 0; e < E; e++) {
            const uint32_t k = mdl->esrc[e], l = mdl->edst[e];
            const double pt = mdl->etrans[e];
            double mm = NEG_INF, im = NEG_INF, dm = NEG_INF;   // TransProb::zero()
            if (i < n) {
                const double pe = mdl->p_match_emit(l, x[i]);
                mm = pdiv(pmul(pmul(pmul(pmul(fi0.m.get(k), pt), pa.p_MM), pe), bi2.m.get(l)), p);
                im = pdiv(pmul(pmul(pmul(pmul(fi0.i.get(k), pt), pa.p_IM), pe), bi2.m.get(l)), p);
                dm = pdiv(pmul(pmul(pmul(pmul(fi0.d.get(k), pt), pa.p_DM), pe), bi2.m.get(l)), p);
            }
            const double md = pdiv(pmul(pmul(pmul(fi0.m.get(k), pt), pa.p_MD), bi1.d.get(l)), p);
            const double id = pdiv(pmul(pmul(pmul(fi0.i.get(k), pt), pa.p_ID), bi1.d.get(l)), p);
            const double dd = pdiv(pmul(pmul(pmul(fi0.d.get(k), pt), pa.p_DD), bi1.d.get(l)), p);
            edge[e] += std::exp(padd(padd(padd(padd(padd(mm, md), im), id), dm), dd));   // trans_table.rs:40-42
        }
        for (size_t v = 0; v < N; v++) {
            const double pi_ = mdl->init[v];
            double mm = NEG_INF, im = NEG_INF;
            if (i < n) {
                const double pe = mdl->p_match_emit((uint32_t)v, x[i]);
                mm = pdiv(pmul(pmul(pmul(pmul(fi0.mb, pi_), pa.p_MM), pe), bi2.m.get((uint32_t)v)), p);
                im = pdiv(pmul(pmul(pmul(pmul(fi0.ib, pi_), pa.p_IM), pe), bi2.m.get((uint32_t)v)), p);
            }
            const double md = pdiv(pmul(pmul(pmul(fi0.mb, pi_), pa.p_MD), bi1.d.get((uint32_t)v)), p);
            const double id = pdiv(pmul(pmul(pmul(fi0.ib, pi_), pa.p_ID), bi1.d.get((uint32_t)v)), p);
            init[v] += std::exp(padd(padd(padd(padd(padd(mm, md), im), id), NEG_INF), NEG_INF));
        }
    }
    return 0;
    ORC_CATCH(1)
}
// PHMMOutput::to_mapping / to_mapping_by_score_ratio (hint.rs:124-142) for one read.
// Two-call protocol: row_counts[n] always written; nodes/probs written if non-null.
int orc_output_mapping(const Tables* f, const Tables* b, int by_ratio, uint64_t n_active, double ratio,
                       uint64_t* row_counts, uint32_t* nodes, double* probs) {
    ORC_TRY
    Output o{*f, *b};
    std::vector<std::vector<uint32_t>> ns; std::vector<std::vector<double>> ps;
    o.to_mapping(by_ratio, n_active, ratio, ns, ps);
    size_t off = 0;
    for (size_t i = 0; i < ns.size(); i++) {
        row_counts[i] = ns[i].size();
        if (nodes) for (size_t j = 0; j < ns[i].size(); j++) { nodes[off + j] = ns[i][j]; probs[off + j] = ps[i][j]; }
        off += ns[i].size();
    }
    return 0;
    ORC_CATCH(1)
}

// ---- bulk calls over a read set (the rayon-parallel entry points; OpenMP over reads) ----
// reads: offsets[R+1] into bases.  mappings (nullable): map_read_off[R+1] rows, map_row_off[rows+1], map_nodes.

// freq.rs:175-192 to_full_prob_reads ; per_read[R] written, returns sum (fixed read order).
double orc_full_prob_reads(const Model* m, uint64_t R, const uint64_t* off, const uint8_t* bases,
                           const uint64_t* map_read_off, const uint64_t* map_row_off, const uint32_t* map_nodes,
                           int use_max_ratio, double* per_read, int n_threads) {
    std::vector<double> local(R);
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int64_t r = 0; r < (int64_t)R; r++) {
        try {
            const uint8_t* x = bases + off[r]; size_t n = off[r + 1] - off[r];
            if (map_read_off) {
                auto mp = unpack_mapping(n, map_row_off + map_read_off[r], map_nodes);
                local[r] = m->forward_with_mapping_score_only(x, n, mp);
            } else local[r] = m->forward_sparse_score_only(x, n, use_max_ratio);
        } catch (...) { failed = 1; }
    }
    if (failed) { g_err = "oracle error in full_prob_reads"; return NAN; }
    double s = 0.0;
    for (uint64_t r = 0; r < R; r++) { s = pmul(s, local[r]); if (per_read) per_read[r] = local[r]; }
    return s;
}

// mode: 0 run (dense), 1 run_sparse, 2 run_sparse_adaptive(use_max_ratio), 3 run_with_mapping   (freq.rs:42-76)
// Accumulates node_freqs[N] over reads (freq.rs:87-102 semantics with the chosen run mode),
// writes logp_fwd[R] = forward.full_prob, logp_bwd[R] = backward.full_prob.
int orc_run_node_freqs(const Model* m, uint64_t R, const uint64_t* off, const uint8_t* bases, int mode, int use_max_ratio,
                       const uint64_t* map_read_off, const uint64_t* map_row_off, const uint32_t* map_nodes,
                       double* node_freqs, double* logp_fwd, double* logp_bwd, int n_threads) {
    uint32_t N = m->n_nodes;
    if (node_freqs) std::fill(node_freqs, node_freqs + N, 0.0);
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int64_t r = 0; r < (int64_t)R; r++) {
        try {
            const uint8_t* x = bases + off[r]; size_t n = off[r + 1] - off[r];
            Output o;
            if (mode == 0) { o.f = m->forward(x, n); o.b = m->backward(x, n); }
            else if (mode == 1) { o.f = m->forward_sparse(x, n, false); o.b = m->backward_sparse(x, n); }
            else if (mode == 2) { o.f = m->forward_sparse(x, n, use_max_ratio); o.b = m->backward_by_forward(x, n, o.f); }
            else { auto mp = unpack_mapping(n, map_row_off + map_read_off[r], map_nodes);
                   o.f = m->forward_with_mapping(x, n, mp); o.b = m->backward_with_mapping(x, n, mp); }
            if (logp_fwd) logp_fwd[r] = o.f.full_prob();
            if (logp_bwd) logp_bwd[r] = o.b.full_prob();
            if (node_freqs) {
                auto fr = o.to_node_freqs();
#pragma omp critical
                for (uint32_t k = 0; k < N; k++) node_freqs[k] += fr[k];
            }
        } catch (...) { failed = 1; }
    }
    if (failed) { g_err = "oracle error in run_node_freqs"; return 1; }
    return 0;
}

// Cells evaluated by f_step/b_step for a run mode (sum over rows of |nodes|, dense rows count N).
// Forward only / backward only / both selected by `dir` (1 fwd, 2 bwd, 3 both).  Used by bench accounting tests.
uint64_t orc_count_cells(const Model* m, const uint8_t* x, uint64_t n, int mode, int use_max_ratio, int dir) {
    uint64_t c = 0; uint32_t N = m->n_nodes;
    Output o;
    if (mode == 0) { return (uint64_t)N * n * ((dir & 1 ? 1 : 0) + (dir & 2 ? 1 : 0)); }
    if (mode == 1) { o.f = m->forward_sparse(x, n, false); o.b = m->backward_sparse(x, n); }
    else { o.f = m->forward_sparse(x, n, use_max_ratio); o.b = m->backward_by_forward(x, n, o.f); }
    if (dir & 1) for (auto& t : o.f.tables) c += t.m.n_elements();
    if (dir & 2) for (auto& t : o.b.tables) c += t.m.n_elements();
    return c;
}

// hint.rs:193-220 generate_mappings for a read set.  Result handle + export.
Mappings* orc_generate_mappings(const Model* m, uint64_t R, const uint64_t* off, const uint8_t* bases,
                                const uint64_t* map_read_off, const uint64_t* map_row_off, const uint32_t* map_nodes,
                                int use_max_ratio, int n_threads) {
    Mappings* out = new Mappings();
    out->nodes.resize(R); out->probs.resize(R);
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int64_t r = 0; r < (int64_t)R; r++) {
        try {
            const uint8_t* x = bases + off[r]; size_t n = off[r + 1] - off[r];
            Output o;
            if (map_read_off) { auto mp = unpack_mapping(n, map_row_off + map_read_off[r], map_nodes);
                                o.f = m->forward_with_mapping(x, n, mp); o.b = m->backward_with_mapping(x, n, mp); }
            else { o.f = m->forward_sparse(x, n, use_max_ratio); o.b = m->backward_by_forward(x, n, o.f); }
            o.to_mapping(use_max_ratio, m->param.n_active_nodes, m->param.active_node_max_ratio, out->nodes[r], out->probs[r]);
        } catch (const std::exception& ex) { failed = 1; }
    }
    if (failed) { delete out; g_err = "oracle error in generate_mappings"; return nullptr; }
    return out;
}
void orc_mappings_destroy(Mappings* mp) { delete mp; }
uint64_t orc_mappings_n_entries(const Mappings* mp) {
    uint64_t c = 0; for (auto& r : mp->nodes) for (auto& row : r) c += row.size(); return c;
}
uint64_t orc_mappings_n_rows(const Mappings* mp) { uint64_t c = 0; for (auto& r : mp->nodes) c += r.size(); return c; }
// read_off[R+1] (rows), row_off[rows+1] (entries), nodes[], probs[]
void orc_mappings_export(const Mappings* mp, uint64_t* read_off, uint64_t* row_off, uint32_t* nodes, double* probs) {
    uint64_t row = 0, ent = 0; read_off[0] = 0; row_off[0] = 0;
    for (size_t r = 0; r < mp->nodes.size(); r++) {
        for (size_t i = 0; i < mp->nodes[r].size(); i++) {
            for (size_t j = 0; j < mp->nodes[r][i].size(); j++) { nodes[ent] = mp->nodes[r][i][j]; probs[ent] = mp->probs[r][i][j]; ent++; }
            row++; row_off[row] = ent;
        }
        read_off[r + 1] = row;
    }
}
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
