// dbgphmm_b200.hpp — C++ host-side mirror of the reference's `impl PHMMModel` / `impl PHMMOutput` (src/hmmv2) over the
// C ABI of dbgphmm_b200.h.  Header only; same method names and argument meaning as the Rust methods; where the reference
// panics this throws dbgphmm::Error.  The reference is compiled Rust and no Rust toolchain exists in this image, so this
// is the compiled-language host layer (see INTEGRATION.md for the Rust shim a maintainer would add).
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "dbgphmm_b200.h"

namespace dbgphmm {

struct Error : std::runtime_error {
    int status;
    Error(int st) : std::runtime_error(dbgphmm_last_error()), status(st) {}
};
inline void check(int st) { if (st != DBGPHMM_OK) throw Error(st); }

using PHMMParams = dbgphmm_params;
inline PHMMParams uniform(double p) { PHMMParams q; dbgphmm_params_uniform(p, &q); return q; }   // params.rs:116

class Reads {  // ReadCollection (common/collection.rs:131)
public:
    explicit Reads(const std::vector<std::string>& seqs) : n_(seqs.size()) {
        std::vector<uint64_t> off(1, 0); std::string all;
        for (auto& s : seqs) { all += s; off.push_back(all.size()); }
        check(dbgphmm_reads_create(seqs.size(), off.data(), (const uint8_t*)all.data(), &h_));
    }
    ~Reads() { dbgphmm_reads_destroy(h_); }
    Reads(const Reads&) = delete; Reads& operator=(const Reads&) = delete;
    uint64_t len() const { return n_; }
    dbgphmm_reads* handle() const { return h_; }
private:
    dbgphmm_reads* h_ = nullptr;
    uint64_t n_;
};

// e2e::Dataset (e2e.rs:31-130): genome, genome_size, positioned reads and the PHMM parameters they were sampled with, in the reference's
// JSON form (Dataset::to_json_file / from_json_file).  reads() hands the read set to the hot path.
class Dataset {
public:
    explicit Dataset(dbgphmm_dataset* h) : h_(h) {}
    static Dataset from_json_file(const std::string& path) { dbgphmm_dataset* h = nullptr; check(dbgphmm_dataset_from_json_file(path.c_str(), &h)); return Dataset(h); }
    static Dataset from_json_str(const std::string& text) { dbgphmm_dataset* h = nullptr; check(dbgphmm_dataset_from_json_text(text.data(), text.size(), &h)); return Dataset(h); }
    ~Dataset() { if (h_) dbgphmm_dataset_destroy(h_); }
    Dataset(const Dataset&) = delete; Dataset& operator=(const Dataset&) = delete;
    Dataset(Dataset&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    uint64_t genome_size() const { uint64_t s[5]; check(dbgphmm_dataset_sizes(h_, s)); return s[4]; }
    uint64_t n_reads() const { uint64_t s[5]; check(dbgphmm_dataset_sizes(h_, s)); return s[2]; }
    double coverage() const { uint64_t s[5]; check(dbgphmm_dataset_sizes(h_, s)); return s[4] ? (double)s[3] / (double)s[4] : 0.0; }   // e2e.rs:72-74
    std::vector<std::string> genome() const {   // "C|L|F:bases" per haplotype (StyledSequence's Display, collection.rs:460-464)
        uint64_t s[5]; check(dbgphmm_dataset_sizes(h_, s));
        std::vector<uint64_t> off(s[0] + 1); std::string bases(s[1], ' '); std::string style(s[0], ' ');
        check(dbgphmm_dataset_genome(h_, off.data(), (uint8_t*)&bases[0], (uint8_t*)&style[0]));
        std::vector<std::string> out;
        for (uint64_t i = 0; i < s[0]; i++) out.push_back(std::string(1, style[i]) + ":" + bases.substr(off[i], off[i + 1] - off[i]));
        return out;
    }
    std::vector<std::string> reads() const {
        uint64_t s[5]; check(dbgphmm_dataset_sizes(h_, s));
        std::vector<uint64_t> off(s[2] + 1); std::string bases(s[3], ' ');
        check(dbgphmm_dataset_read_origins(h_, off.data(), (uint8_t*)&bases[0], nullptr, nullptr, nullptr));
        std::vector<std::string> out;
        for (uint64_t i = 0; i < s[2]; i++) out.push_back(bases.substr(off[i], off[i + 1] - off[i]));
        return out;
    }
    dbgphmm_params params() const { dbgphmm_params p; check(dbgphmm_dataset_params(h_, &p)); return p; }
    std::string to_json_string() const {
        uint64_t need = 0; check(dbgphmm_dataset_to_json_text(h_, nullptr, 0, &need));
        std::string s(need, ' '); check(dbgphmm_dataset_to_json_text(h_, &s[0], need, &need)); return s;
    }
    void to_json_file(const std::string& path) const { check(dbgphmm_dataset_to_json_file(h_, path.c_str())); }
    dbgphmm_dataset* handle() const { return h_; }
private:
    dbgphmm_dataset* h_ = nullptr;
};

// Mapping (hint.rs:27-30) of one read, copied to the host: per base the candidate nodes and their ln probabilities
struct Mapping {
    std::vector<std::vector<uint32_t>> nodes;
    std::vector<std::vector<double>> probs;
};

class MultiDbg;

class Mappings {  // hint.rs:150-152
public:
    explicit Mappings(dbgphmm_mappings* h) : h_(h) {}
    explicit Mappings(const std::vector<Mapping>& maps) {
        std::vector<uint64_t> read_off(1, 0), row_off(1, 0); std::vector<uint32_t> nodes; std::vector<double> logp;
        for (auto& m : maps) {
            if (m.nodes.size() != m.probs.size()) throw Error(DBGPHMM_ERR_INVALID);
            for (size_t i = 0; i < m.nodes.size(); i++) {
                if (m.nodes[i].size() != m.probs[i].size()) throw Error(DBGPHMM_ERR_INVALID);
                nodes.insert(nodes.end(), m.nodes[i].begin(), m.nodes[i].end());
                logp.insert(logp.end(), m.probs[i].begin(), m.probs[i].end());
                row_off.push_back(nodes.size());
            }
            read_off.push_back(row_off.size() - 1);
        }
        check(dbgphmm_mappings_create(maps.size(), read_off.data(), row_off.data(), nodes.data(), logp.data(), &h_));
    }
    static Mappings from_map_file(const std::string& path) {   // MultiDbg::from_map_file_raw, multi_dbg/output.rs:604-623 (.map, .gz, .mpz)
        dbgphmm_mappings* h = nullptr;
        check(dbgphmm_mappings_from_map_file(path.c_str(), &h));
        return Mappings(h);
    }
    static Mappings from_map_str(const std::string& text) {    // output.rs:589-591
        dbgphmm_mappings* h = nullptr;
        check(dbgphmm_mappings_from_map_text(text.data(), text.size(), &h));
        return Mappings(h);
    }
    ~Mappings() { if (h_) dbgphmm_mappings_destroy(h_); }
    Mappings(const Mappings&) = delete; Mappings& operator=(const Mappings&) = delete;
    Mappings(Mappings&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    dbgphmm_mappings* handle() const { return h_; }
    uint64_t n_reads() const { uint64_t n = 0; check(dbgphmm_mappings_sizes(h_, &n, nullptr, nullptr)); return n; }
    // `mappings[r]`
    Mapping at(uint64_t r) const {
        uint64_t nr = 0, nrow = 0, nent = 0;
        check(dbgphmm_mappings_sizes(h_, &nr, &nrow, &nent));
        if (r >= nr) throw Error(DBGPHMM_ERR_INVALID);
        std::vector<uint64_t> ro(nr + 1), rw(nrow + 1); std::vector<uint32_t> nd(nent); std::vector<double> lp(nent);
        check(dbgphmm_mappings_export(h_, ro.data(), rw.data(), nd.data(), lp.data()));
        Mapping m;
        for (uint64_t i = ro[r]; i < ro[r + 1]; i++) {
            m.nodes.emplace_back(nd.begin() + rw[i], nd.begin() + rw[i + 1]);
            m.probs.emplace_back(lp.begin() + rw[i], lp.begin() + rw[i + 1]);
        }
        return m;
    }
    std::vector<double> to_node_freqs(uint32_t n_nodes) const {  // hint.rs:161
        std::vector<double> f(n_nodes); check(dbgphmm_mappings_to_node_freqs(h_, n_nodes, f.data())); return f;
    }
    // Mapping::map_nodes (hint.rs:66-88) over every read; node_map[v] = the nodes v turns into (may be empty)
    Mappings map_nodes(const std::vector<std::vector<uint32_t>>& node_map) const {
        std::vector<uint64_t> off(1, 0); std::vector<uint32_t> to;
        for (auto& ws : node_map) { to.insert(to.end(), ws.begin(), ws.end()); off.push_back(to.size()); }
        dbgphmm_mappings* o = nullptr;
        check(dbgphmm_mappings_map_nodes(h_, (uint32_t)node_map.size(), off.data(), to.data(), &o));
        return Mappings(o);
    }
    // MultiDbg::to_map_file (output.rs:455-527); dbg (nullable) only feeds the header comment
    inline void to_map_file(const std::string& path, const Reads& reads, const MultiDbg* dbg = nullptr) const;
private:
    dbgphmm_mappings* h_ = nullptr;
};

// PHMMTable (table.rs:42-73) copied to the host, natural logs.  Dense: m, i, d hold all N nodes and ids / ids_d are empty.
// Sparse: (ids[j], m[j], i[j]) and (ids_d[j], d[j]) in the insertion order of the reference's SparseVec.
struct PHMMTable {
    bool is_dense = true;
    std::vector<uint32_t> ids, ids_d;
    std::vector<double> m, i, d;
    double mb = 0, ib = 0, e = 0;
};

class PHMMTables {  // table.rs:365-435, resident on the device
public:
    explicit PHMMTables(dbgphmm_tables* h) : h_(h) {}
    ~PHMMTables() { dbgphmm_tables_destroy(h_); }
    PHMMTables(const PHMMTables&) = delete; PHMMTables& operator=(const PHMMTables&) = delete;
    uint64_t n_emissions() const { return dbgphmm_tables_len(h_); }
    double full_prob() const { double v; check(dbgphmm_tables_full_prob(h_, &v)); return v; }   // table.rs:395
    // tables[row]; row = -1 is init_table (table.rs:365-380)
    PHMMTable table(int64_t row) const {
        uint64_t info[3]; double sc[3];
        check(dbgphmm_tables_row_info(h_, row, info, sc));
        PHMMTable t;
        t.is_dense = info[0] != 0; t.mb = sc[0]; t.ib = sc[1]; t.e = sc[2];
        t.m.resize(info[1]); t.i.resize(info[1]); t.d.resize(info[2]);
        if (!t.is_dense) { t.ids.resize(info[1]); t.ids_d.resize(info[2]); }
        check(dbgphmm_tables_row_export(h_, row, t.is_dense ? nullptr : t.ids.data(), t.m.data(), t.i.data(),
                                        t.is_dense ? nullptr : t.ids_d.data(), t.d.data()));
        return t;
    }
    PHMMTable init_table() const { return table(-1); }
    // PHMMTables::table_merged (table.rs:414-434); the direction is the caller's (Forward[0] / Backward[n] = init_table)
    PHMMTable table_merged(bool is_forward, uint64_t merged_index) const {
        if (is_forward) return table(merged_index == 0 ? -1 : (int64_t)merged_index - 1);
        return table(merged_index >= n_emissions() ? -1 : (int64_t)merged_index);
    }
    std::vector<uint32_t> top_nodes(int64_t row, uint32_t k) const { return top(row, 0, k, 0.0); }                        // table.rs:127
    std::vector<uint32_t> top_nodes_by_score_ratio(int64_t row, double ratio) const { return top(row, 1, 0, ratio); }    // table.rs:134
    dbgphmm_tables* handle() const { return h_; }
private:
    std::vector<uint32_t> top(int64_t row, int by_ratio, uint32_t k, double ratio) const {
        std::vector<uint32_t> out(DBGPHMM_MAX_ACTIVE_NODES); uint32_t n = 0;
        check(dbgphmm_tables_row_top_nodes(h_, row, by_ratio, k, ratio, out.data(), &n));
        out.resize(n);
        return out;
    }
    dbgphmm_tables* h_;
};

// PHMMOutput (table.rs:450-517): the forward and backward tables of one read and the products over both
class PHMMOutput {
public:
    PHMMOutput(dbgphmm_model* m, uint32_t n_nodes, uint32_t n_edges, std::unique_ptr<PHMMTables> f, std::unique_ptr<PHMMTables> b)
        : forward(std::move(f)), backward(std::move(b)), m_(m), n_nodes_(n_nodes), n_edges_(n_edges) {
        if (forward->n_emissions() != backward->n_emissions()) throw Error(DBGPHMM_ERR_INVALID);   // table.rs:473
    }
    std::unique_ptr<PHMMTables> forward, backward;
    uint64_t n_emissions() const { return forward->n_emissions(); }
    double to_full_prob_forward() const { return forward->full_prob(); }     // table.rs:482
    double to_full_prob_backward() const { return backward->full_prob(); }   // table.rs:492
    std::vector<double> to_node_freqs() const {                              // freq.rs:245-255
        std::vector<double> f(n_nodes_);
        check(dbgphmm_output_node_freqs(m_, forward->handle(), backward->handle(), f.data()));
        return f;
    }
    // freq.rs:276-298: edge freqs in EdgeIndex order + the Begin -> node freqs
    std::pair<std::vector<double>, std::vector<double>> to_edge_and_init_freqs() const {
        std::vector<double> e(n_edges_), i(n_nodes_);
        check(dbgphmm_output_edge_and_init_freqs(m_, forward->handle(), backward->handle(), e.data(), i.data()));
        return {std::move(e), std::move(i)};
    }
    std::vector<double> to_edge_freqs() const { return to_edge_and_init_freqs().first; }   // freq.rs:302-309
    Mapping to_mapping(uint32_t n_active_nodes) const { return mapping(0, n_active_nodes, 0.0); }        // hint.rs:124-133
    Mapping to_mapping_by_score_ratio(double max_ratio) const { return mapping(1, 0, max_ratio); }        // hint.rs:134-142
private:
    Mapping mapping(int by_ratio, uint32_t n_active, double ratio) const {
        dbgphmm_mappings* h = nullptr;
        check(dbgphmm_output_mapping(m_, forward->handle(), backward->handle(), by_ratio, n_active, ratio, &h));
        return Mappings(h).at(0);
    }
    dbgphmm_model* m_;
    uint32_t n_nodes_, n_edges_;
};

class PHMMModel {  // hmmv2/common.rs:61-67
public:
    PHMMModel(const std::vector<uint32_t>& edge_src, const std::vector<uint32_t>& edge_dst, const std::vector<uint8_t>& emission,
              const std::vector<double>& log_init, const std::vector<double>& log_trans, const PHMMParams& param, int device = 0,
              uint64_t mem_budget = 0) : n_nodes_((uint32_t)emission.size()), n_edges_((uint32_t)edge_src.size()) {
        if (edge_dst.size() != edge_src.size() || log_trans.size() != edge_src.size() || log_init.size() != emission.size()) throw Error(DBGPHMM_ERR_INVALID);
        check(dbgphmm_model_create(n_nodes_, n_edges_, edge_src.data(), edge_dst.data(), emission.data(), log_init.data(),
                                   log_trans.data(), &param, device, mem_budget, &h_));
    }
    // adopts a handle made by dbgphmm_dbg_to_model (MultiDbg::to_phmm below)
    PHMMModel(dbgphmm_model* h, uint32_t n_edges) : h_(h), n_nodes_(dbgphmm_model_n_nodes(h)), n_edges_(n_edges) {}
    ~PHMMModel() { dbgphmm_model_destroy(h_); }
    PHMMModel(const PHMMModel&) = delete; PHMMModel& operator=(const PHMMModel&) = delete;
    uint32_t n_nodes() const { return n_nodes_; }
    uint32_t n_edges() const { return n_edges_; }
    void set_params(const PHMMParams& param) { check(dbgphmm_model_set_params(h_, &param)); }
    void set_probs(const std::vector<double>& log_init, const std::vector<double>& log_trans) {
        if (log_init.size() != n_nodes_ || log_trans.size() != n_edges_) throw Error(DBGPHMM_ERR_INVALID);
        check(dbgphmm_model_set_probs(h_, log_init.data(), log_trans.data()));
    }
    // candidate copy-number assignments X -> parameter sets on the device (seq_graph.rs:160-273)
    void set_copy_nums_batch(uint32_t n_batch, const uint32_t* copy_nums, int mode = 0) { check(dbgphmm_model_set_copy_nums_batch(h_, n_batch, copy_nums, mode)); }
    uint32_t n_batch() const { return dbgphmm_model_n_batch(h_); }
    uint32_t wave_reads() const { return dbgphmm_model_wave_reads(h_); }   // reads whose sparse rows are resident at once: the efficient batch quantum

    std::unique_ptr<PHMMTables> forward(const std::string& x) { return fwd(x, DBGPHMM_FWD_DENSE); }                           // forward.rs:24
    std::unique_ptr<PHMMTables> forward_sparse(const std::string& x, bool use_max_ratio) { return fwd(x, use_max_ratio ? DBGPHMM_FWD_SPARSE_RATIO : DBGPHMM_FWD_SPARSE); }  // :93
    std::unique_ptr<PHMMTables> forward_with_mapping(const std::string& x, const Mappings& m, uint64_t i) { return fwd(x, DBGPHMM_FWD_MAPPING, &m, i); }  // :51
    std::unique_ptr<PHMMTables> backward(const std::string& x) { return bwd(x, DBGPHMM_BWD_DENSE); }                          // backward.rs:24
    std::unique_ptr<PHMMTables> backward_sparse(const std::string& x) { return bwd(x, DBGPHMM_BWD_SPARSE); }                  // :146
    std::unique_ptr<PHMMTables> backward_with_mapping(const std::string& x, const Mappings& m, uint64_t i) { return bwd(x, DBGPHMM_BWD_MAPPING, &m, i); }  // :59
    std::unique_ptr<PHMMTables> backward_by_forward(const std::string& x, const PHMMTables& f) { return bwd(x, DBGPHMM_BWD_BY_FORWARD, nullptr, 0, &f); }  // :101

    // freq.rs:42-76
    PHMMOutput run(const std::string& x) { return out(forward(x), backward(x)); }
    PHMMOutput run_sparse(const std::string& x) { return out(forward_sparse(x, false), backward_sparse(x)); }
    PHMMOutput run_sparse_adaptive(const std::string& x, bool use_max_ratio) {
        auto f = forward_sparse(x, use_max_ratio);
        auto b = backward_by_forward(x, *f);
        return out(std::move(f), std::move(b));
    }
    PHMMOutput run_with_mapping(const std::string& x, const Mappings& m, uint64_t i) { return out(forward_with_mapping(x, m, i), backward_with_mapping(x, m, i)); }

    // freq.rs:175-192 for every candidate X: returns ln P(R|X) [n_batch]
    std::vector<double> to_full_prob_reads(const Reads& reads, const Mappings* mappings, bool use_max_ratio) {
        std::vector<double> out(dbgphmm_model_n_batch(h_));
        check(dbgphmm_to_full_prob_reads(h_, reads.handle(), mappings ? mappings->handle() : nullptr, use_max_ratio, out.data(), nullptr));
        return out;
    }
    // run / run_sparse / run_sparse_adaptive / run_with_mapping (freq.rs:42-76) + to_node_freqs (freq.rs:245), summed over reads;
    // logp_fwd / logp_bwd (nullable) receive to_full_prob_forward / to_full_prob_backward of every read
    std::vector<double> to_node_freqs(const Reads& reads, int run_mode, bool use_max_ratio = true, const Mappings* mappings = nullptr,
                                      std::vector<double>* logp_fwd = nullptr, std::vector<double>* logp_bwd = nullptr) {
        std::vector<double> f(n_nodes_);
        if (logp_fwd) logp_fwd->resize(reads.len());
        if (logp_bwd) logp_bwd->resize(reads.len());
        check(dbgphmm_run_node_freqs(h_, reads.handle(), run_mode, use_max_ratio, mappings ? mappings->handle() : nullptr, f.data(),
                                     logp_fwd ? logp_fwd->data() : nullptr, logp_bwd ? logp_bwd->data() : nullptr, nullptr));
        return f;
    }
    // PHMMModel::to_node_freqs (freq.rs:87-102): dense forward + backward per read
    std::vector<double> to_node_freqs(const Reads& reads) { return to_node_freqs(reads, DBGPHMM_RUN_DENSE); }
    // to_full_prob / to_full_prob_parallel (freq.rs:105-135): ln of the product over reads of the dense forward P(read), in read order
    double to_full_prob(const Reads& reads) { std::vector<double> lf; to_node_freqs(reads, DBGPHMM_RUN_DENSE, true, nullptr, &lf); return sum(lf); }
    double to_full_prob_parallel(const Reads& reads) { return to_full_prob(reads); }
    // freq.rs:138-150 (parameter set 0)
    double to_full_prob_sparse(const Reads& reads, bool use_max_ratio) { return to_full_prob_reads(reads, nullptr, use_max_ratio)[0]; }
    // freq.rs:153-164
    double to_full_prob_sparse_backward(const Reads& reads) {
        std::vector<double> lb; to_node_freqs(reads, DBGPHMM_RUN_SPARSE, true, nullptr, nullptr, &lb); return sum(lb);
    }
    // forward.rs:158-206 / 79-89
    double forward_sparse_score_only(const std::string& x, bool use_max_ratio) {
        Reads one(std::vector<std::string>{x});
        return to_full_prob_reads(one, nullptr, use_max_ratio)[0];
    }
    double forward_with_mapping_score_only(const std::string& x, const Mapping& m) {
        Reads one(std::vector<std::string>{x});
        Mappings hint(std::vector<Mapping>{m});
        return to_full_prob_reads(one, &hint, false)[0];
    }
    Mappings generate_mappings(const Reads& reads, const Mappings* mappings, bool use_max_ratio) {  // hint.rs:193
        dbgphmm_mappings* out = nullptr;
        check(dbgphmm_generate_mappings(h_, reads.handle(), mappings ? mappings->handle() : nullptr, use_max_ratio, &out));
        return Mappings(out);
    }
    // q_score_exact (q.rs:66-96) of parameter set x: {init, trans, prior}
    std::vector<double> q_score_exact(const std::vector<double>& edge_freqs, const std::vector<double>& init_freqs, uint32_t x = 0) const {
        if (edge_freqs.size() != n_edges_ || init_freqs.size() != n_nodes_) throw Error(DBGPHMM_ERR_INVALID);
        std::vector<double> q(3);
        check(dbgphmm_q_score_exact(h_, x, edge_freqs.data(), init_freqs.data(), q.data()));
        return q;
    }
    dbgphmm_model* handle() const { return h_; }

private:
    static double sum(const std::vector<double>& v) { double s = 0; for (double x : v) s += x; return s; }
    PHMMOutput out(std::unique_ptr<PHMMTables> f, std::unique_ptr<PHMMTables> b) { return PHMMOutput(h_, n_nodes_, n_edges_, std::move(f), std::move(b)); }
    std::unique_ptr<PHMMTables> fwd(const std::string& x, int kind, const Mappings* m = nullptr, uint64_t i = 0) {
        dbgphmm_tables* t = nullptr;
        check(dbgphmm_forward(h_, (const uint8_t*)x.data(), x.size(), kind, m ? m->handle() : nullptr, i, &t));
        return std::make_unique<PHMMTables>(t);
    }
    std::unique_ptr<PHMMTables> bwd(const std::string& x, int kind, const Mappings* m = nullptr, uint64_t i = 0, const PHMMTables* f = nullptr) {
        dbgphmm_tables* t = nullptr;
        check(dbgphmm_backward(h_, (const uint8_t*)x.data(), x.size(), kind, m ? m->handle() : nullptr, i, f ? f->handle() : nullptr, &t));
        return std::make_unique<PHMMTables>(t);
    }
    dbgphmm_model* h_ = nullptr;
    uint32_t n_nodes_, n_edges_;
};

// Score (multi_dbg/posterior.rs:164-208), natural logs
struct Score {
    double likelihood = 0, prior = 0;
    uint64_t genome_size = 0;
    double n_euler_circuits = 0;
    double p() const { return likelihood + prior + n_euler_circuits; }   // P(R|X) P(G) #circuits
};

// Prob + Prob (prob.rs:181-197)
inline double prob_add(double a, double b) {
    const double x = a >= b ? a : b, y = a >= b ? b : a;
    if (y == -INFINITY) return x;
    if (x == y) return x + 0.693147180559945309417232121458;
    return x + std::log1p(std::exp(y - x));
}

// PosteriorSample / Posterior (multi_dbg/posterior.rs:42-161): the distinct copy-number vectors seen so far with their scores
struct PosteriorSample {
    std::vector<uint32_t> copy_nums;
    Score score;
};
class Posterior {
public:
    bool contains(const std::vector<uint32_t>& copy_nums) const { return find(copy_nums) != nullptr; }
    const PosteriorSample* find(const std::vector<uint32_t>& copy_nums) const {
        for (auto& s : samples_) if (s.copy_nums == copy_nums) return &s;
        return nullptr;
    }
    void add(const PosteriorSample& sample) {                          // posterior.rs:93-98: a vector counts once
        if (!contains(sample.copy_nums)) { p_ = prob_add(p_, sample.score.p()); samples_.push_back(sample); }
    }
    const PosteriorSample& max_sample() const {                        // posterior.rs:113-118 (max_by_key: the last of equal maxima)
        if (samples_.empty()) throw Error(DBGPHMM_ERR_INVALID);
        size_t best = 0;
        for (size_t i = 1; i < samples_.size(); i++) if (samples_[i].score.p() >= samples_[best].score.p()) best = i;
        return samples_[best];
    }
    const std::vector<uint32_t>& max_copy_nums() const { return max_sample().copy_nums; }
    const std::vector<PosteriorSample>& samples() const { return samples_; }
    double p() const { return p_; }                                    // ln of the normalisation factor
    double p_edge_x(size_t edge, uint32_t x) const {                   // ln P(X[edge] = x | R), posterior.rs:141-159 over hist.rs:48-71
        double z = -INFINITY, px = -INFINITY;
        for (auto& s : samples_) {
            const double w = s.score.p() - p_;
            z = prob_add(z, w);
            if (s.copy_nums.at(edge) == x) px = prob_add(px, w);
        }
        return px == -INFINITY ? -INFINITY : px - z;
    }
private:
    std::vector<PosteriorSample> samples_;
    double p_ = -INFINITY;
};

// The part of MultiDbg a DBG file carries (multi_dbg.rs:170-186, multi_dbg/output.rs:155-345); host only.
class MultiDbg {
public:
    static std::unique_ptr<MultiDbg> from_dbg_str(const std::string& s) {   // output.rs:340
        dbgphmm_dbg* h = nullptr;
        check(dbgphmm_dbg_from_text(s.data(), s.size(), &h));
        return std::unique_ptr<MultiDbg>(new MultiDbg(h));
    }
    static std::unique_ptr<MultiDbg> from_dbg_file(const std::string& path) {   // output.rs:346 (.dbg, .dbg.gz, .dbz)
        dbgphmm_dbg* h = nullptr;
        check(dbgphmm_dbg_from_file(path.c_str(), &h));
        return std::unique_ptr<MultiDbg>(new MultiDbg(h));
    }
    ~MultiDbg() { dbgphmm_dbg_destroy(h_); }
    MultiDbg(const MultiDbg&) = delete; MultiDbg& operator=(const MultiDbg&) = delete;
    uint32_t k() const { return sz_[0]; }
    uint32_t n_edges_full() const { return sz_[2]; }
    uint32_t n_edges_compact() const { return sz_[4]; }
    std::string to_dbg_string() const {   // output.rs:129
        uint64_t need = 0;
        check(dbgphmm_dbg_to_text(h_, nullptr, 0, &need));
        std::string s(need, '\0');
        check(dbgphmm_dbg_to_text(h_, &s[0], need, &need));
        return s;
    }
    void to_dbg_file(const std::string& path) const { check(dbgphmm_dbg_to_file(h_, path.c_str())); }
    std::vector<uint32_t> get_copy_nums() const {   // multi_dbg.rs:1056
        std::vector<uint32_t> x(sz_[4]);
        check(dbgphmm_dbg_get_copy_nums(h_, x.data()));
        return x;
    }
    void set_copy_nums(const std::vector<uint32_t>& x) {   // multi_dbg.rs:1041 (fails where the reference asserts)
        if (x.size() != sz_[4]) throw Error(DBGPHMM_ERR_INVALID);
        check(dbgphmm_dbg_set_copy_nums(h_, x.data()));
    }
    // candidates over compact edges [n_batch][n_edges_compact] -> per-k-mer copy numbers [n_batch][n_edges_full]
    std::vector<uint32_t> expand_copy_nums(uint32_t n_batch, const std::vector<uint32_t>& compact) const {
        if (compact.size() != (size_t)n_batch * sz_[4]) throw Error(DBGPHMM_ERR_INVALID);
        std::vector<uint32_t> full((size_t)n_batch * sz_[2]);
        check(dbgphmm_dbg_expand_copy_nums(h_, n_batch, compact.data(), full.data()));
        return full;
    }
    // to_phmm (mode 0) / to_non_zero_phmm (1) / to_uniform_phmm (2): multi_dbg.rs:1391-1409 ; the caller owns the handle
    dbgphmm_model* to_phmm_handle(const PHMMParams& param, int mode = 0, int device = 0, uint64_t mem_budget = 0) const {
        dbgphmm_model* m = nullptr;
        check(dbgphmm_dbg_to_model(h_, &param, mode, device, mem_budget, &m));
        return m;
    }
    // The terms of to_score beside the likelihood (multi_dbg/posterior.rs:225-277) for the current copy numbers ...
    uint64_t genome_size() const { uint64_t g = 0; check(dbgphmm_dbg_genome_size(h_, 1, nullptr, &g)); return g; }            // multi_dbg.rs:1018
    double n_euler_circuits() const { double v = 0; check(dbgphmm_dbg_n_euler_circuits(h_, 1, nullptr, &v)); return v; }        // multi_dbg.rs:831 (ln)
    double to_prior(uint32_t genome_size_expected, uint32_t genome_size_sigma) const {                                           // posterior.rs:225
        double v = 0; check(dbgphmm_prior_normal((double)genome_size(), genome_size_expected, genome_size_sigma, &v)); return v;
    }
    // ... and for a batch of candidates over compact edges [n_batch][n_edges_compact]
    std::vector<uint64_t> genome_size(uint32_t n_batch, const std::vector<uint32_t>& compact) const {
        if (compact.size() != (size_t)n_batch * sz_[4]) throw Error(DBGPHMM_ERR_INVALID);
        std::vector<uint64_t> g(n_batch); check(dbgphmm_dbg_genome_size(h_, n_batch, compact.data(), g.data())); return g;
    }
    std::vector<double> n_euler_circuits(uint32_t n_batch, const std::vector<uint32_t>& compact) const {
        if (compact.size() != (size_t)n_batch * sz_[4]) throw Error(DBGPHMM_ERR_INVALID);
        std::vector<double> v(n_batch); check(dbgphmm_dbg_n_euler_circuits(h_, n_batch, compact.data(), v.data())); return v;
    }
    // MultiDbg::to_score (posterior.rs:259-277) for every candidate at once: one expansion, one on-device (init, trans) derivation,
    // one batched to_full_prob_reads on `phmm` (a model of this graph, to_phmm), then the host-side terms
    std::vector<Score> to_scores(PHMMModel& phmm, const Reads& reads, const Mappings* mappings, uint32_t n_batch, const std::vector<uint32_t>& compact,
                                 uint32_t genome_size_expected, uint32_t genome_size_sigma, int mode = 0) const {
        const std::vector<uint32_t> full = expand_copy_nums(n_batch, compact);
        phmm.set_copy_nums_batch(n_batch, full.data(), mode);
        const std::vector<double> like = phmm.to_full_prob_reads(reads, mappings, true);
        const std::vector<uint64_t> gs = genome_size(n_batch, compact);
        const std::vector<double> ne = n_euler_circuits(n_batch, compact);
        std::vector<Score> out(n_batch);
        for (uint32_t b = 0; b < n_batch; b++) {
            out[b].likelihood = like[b]; out[b].genome_size = gs[b]; out[b].n_euler_circuits = ne[b];
            check(dbgphmm_prior_normal((double)gs[b], genome_size_expected, genome_size_sigma, &out[b].prior));
        }
        return out;
    }
    // MultiDbg::sample_posterior_once (posterior.rs:470-600), single-move form: every neighbour the posterior has not seen yet is
    // scored in ONE batched to_scores (the reference clones the graph per neighbour under rayon); returns true and the best sample if
    // it is not the current copy-number vector.  `neighbors`: copy-number vectors over compact edges (the neighbour search of
    // neighbors.rs is not part of this library).
    bool sample_posterior_once(PHMMModel& phmm, const Reads& reads, const Mappings* mappings, const std::vector<std::vector<uint32_t>>& neighbors,
                               Posterior& posterior, uint32_t genome_size_expected, uint32_t genome_size_sigma, PosteriorSample* best, int mode = 0) const {
        std::vector<std::vector<uint32_t>> todo;
        for (auto& c : neighbors) {
            if (c.size() != sz_[4]) throw Error(DBGPHMM_ERR_INVALID);
            bool dup = posterior.contains(c);
            for (auto& t : todo) dup = dup || t == c;
            if (!dup) todo.push_back(c);
        }
        if (!todo.empty()) {
            std::vector<uint32_t> flat;
            for (auto& c : todo) flat.insert(flat.end(), c.begin(), c.end());
            const std::vector<Score> sc = to_scores(phmm, reads, mappings, (uint32_t)todo.size(), flat, genome_size_expected, genome_size_sigma, mode);
            for (size_t i = 0; i < todo.size(); i++) posterior.add(PosteriorSample{todo[i], sc[i]});
        }
        const PosteriorSample& top = posterior.max_sample();
        if (top.copy_nums == get_copy_nums()) return false;
        if (best) *best = top;
        return true;
    }
    // The greedy search of MultiDbg::sample_posterior (posterior.rs:314-420): from the current copy numbers, score the neighbours, move
    // to the best one, stop at a local optimum or after max_iter moves.  neighbors_fn(const MultiDbg&) returns the candidate sets to
    // try in order (std::vector<std::vector<std::vector<uint32_t>>>) for the copy numbers the graph it is given holds.
    template <class NeighborsFn>
    Posterior sample_posterior(PHMMModel& phmm, const Reads& reads, const Mappings* mappings, uint32_t genome_size_expected,
                               uint32_t genome_size_sigma, NeighborsFn neighbors_fn, size_t max_iter, int mode = 0) const {
        Posterior post;
        std::unique_ptr<MultiDbg> dbg = from_dbg_str(to_dbg_string());
        std::vector<uint32_t> copy_nums = dbg->get_copy_nums();
        post.add(PosteriorSample{copy_nums, dbg->to_scores(phmm, reads, mappings, 1, copy_nums, genome_size_expected, genome_size_sigma, mode)[0]});
        for (size_t n_iter = 0; n_iter < max_iter;) {
            dbg->set_copy_nums(copy_nums);
            bool moved = false;
            for (auto& cand : neighbors_fn(*dbg)) {
                PosteriorSample best;
                if (dbg->sample_posterior_once(phmm, reads, mappings, cand, post, genome_size_expected, genome_size_sigma, &best, mode)) {
                    copy_nums = best.copy_nums; n_iter++; moved = true;
                    break;
                }
            }
            if (!moved) break;   // local optimum
        }
        return post;
    }
    // MultiDbg::to_phmm / to_non_zero_phmm / to_uniform_phmm (multi_dbg.rs:1391-1409; n_warmup := k)
    std::unique_ptr<PHMMModel> to_phmm(const PHMMParams& param, int device = 0) const { return std::make_unique<PHMMModel>(to_phmm_handle(param, 0, device), sz_[5]); }
    std::unique_ptr<PHMMModel> to_non_zero_phmm(const PHMMParams& param, int device = 0) const { return std::make_unique<PHMMModel>(to_phmm_handle(param, 1, device), sz_[5]); }
    std::unique_ptr<PHMMModel> to_uniform_phmm(const PHMMParams& param, int device = 0) const { return std::make_unique<PHMMModel>(to_phmm_handle(param, 2, device), sz_[5]); }
    dbgphmm_dbg* handle() const { return h_; }

private:
    explicit MultiDbg(dbgphmm_dbg* h) : h_(h) { check(dbgphmm_dbg_sizes(h_, sz_)); }
    dbgphmm_dbg* h_;
    uint32_t sz_[6];
};

inline void Mappings::to_map_file(const std::string& path, const Reads& reads, const MultiDbg* dbg) const {
    check(dbgphmm_mappings_to_map_file(h_, reads.handle(), dbg ? dbg->handle() : nullptr, path.c_str()));
}

}  // namespace dbgphmm
