// dbgphmm_b200.hpp — C++ host-side mirror of the reference's `impl PHMMModel` / `impl PHMMOutput` (src/hmmv2) over the
// C ABI of dbgphmm_b200.h.  Header only; same method names and argument meaning as the Rust methods; where the reference
// panics this throws dbgphmm::Error.  The reference is compiled Rust and no Rust toolchain exists in this image, so this
// is the compiled-language host layer (see INTEGRATION.md for the Rust shim a maintainer would add).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
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
    explicit Reads(const std::vector<std::string>& seqs) {
        std::vector<uint64_t> off(1, 0); std::string all;
        for (auto& s : seqs) { all += s; off.push_back(all.size()); }
        check(dbgphmm_reads_create(seqs.size(), off.data(), (const uint8_t*)all.data(), &h_));
    }
    ~Reads() { dbgphmm_reads_destroy(h_); }
    Reads(const Reads&) = delete; Reads& operator=(const Reads&) = delete;
    dbgphmm_reads* handle() const { return h_; }
private:
    dbgphmm_reads* h_ = nullptr;
};

class Mappings {  // hint.rs:150-152
public:
    explicit Mappings(dbgphmm_mappings* h) : h_(h) {}
    ~Mappings() { dbgphmm_mappings_destroy(h_); }
    Mappings(const Mappings&) = delete; Mappings& operator=(const Mappings&) = delete;
    dbgphmm_mappings* handle() const { return h_; }
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
    Mappings(Mappings&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
private:
    dbgphmm_mappings* h_;
};

class PHMMTables {  // table.rs:365-435
public:
    explicit PHMMTables(dbgphmm_tables* h) : h_(h) {}
    ~PHMMTables() { dbgphmm_tables_destroy(h_); }
    PHMMTables(const PHMMTables&) = delete; PHMMTables& operator=(const PHMMTables&) = delete;
    uint64_t n_emissions() const { return dbgphmm_tables_len(h_); }
    double full_prob() const { double v; check(dbgphmm_tables_full_prob(h_, &v)); return v; }   // table.rs:395
    dbgphmm_tables* handle() const { return h_; }
private:
    dbgphmm_tables* h_;
};

class PHMMModel {  // hmmv2/common.rs:61-67
public:
    PHMMModel(const std::vector<uint32_t>& edge_src, const std::vector<uint32_t>& edge_dst, const std::vector<uint8_t>& emission,
              const std::vector<double>& log_init, const std::vector<double>& log_trans, const PHMMParams& param, int device = 0,
              uint64_t mem_budget = 0) : n_nodes_((uint32_t)emission.size()) {
        check(dbgphmm_model_create(n_nodes_, (uint32_t)edge_src.size(), edge_src.data(), edge_dst.data(), emission.data(), log_init.data(),
                                   log_trans.data(), &param, device, mem_budget, &h_));
    }
    ~PHMMModel() { dbgphmm_model_destroy(h_); }
    PHMMModel(const PHMMModel&) = delete; PHMMModel& operator=(const PHMMModel&) = delete;
    uint32_t n_nodes() const { return n_nodes_; }
    // candidate copy-number assignments X -> parameter sets on the device (seq_graph.rs:160-273)
    void set_copy_nums_batch(uint32_t n_batch, const uint32_t* copy_nums, int mode = 0) { check(dbgphmm_model_set_copy_nums_batch(h_, n_batch, copy_nums, mode)); }

    std::unique_ptr<PHMMTables> forward(const std::string& x) { return fwd(x, DBGPHMM_FWD_DENSE); }                           // forward.rs:24
    std::unique_ptr<PHMMTables> forward_sparse(const std::string& x, bool use_max_ratio) { return fwd(x, use_max_ratio ? DBGPHMM_FWD_SPARSE_RATIO : DBGPHMM_FWD_SPARSE); }  // :93
    std::unique_ptr<PHMMTables> forward_with_mapping(const std::string& x, const Mappings& m, uint64_t i) { return fwd(x, DBGPHMM_FWD_MAPPING, &m, i); }  // :51
    std::unique_ptr<PHMMTables> backward(const std::string& x) { return bwd(x, DBGPHMM_BWD_DENSE); }                          // backward.rs:24
    std::unique_ptr<PHMMTables> backward_sparse(const std::string& x) { return bwd(x, DBGPHMM_BWD_SPARSE); }                  // :146
    std::unique_ptr<PHMMTables> backward_with_mapping(const std::string& x, const Mappings& m, uint64_t i) { return bwd(x, DBGPHMM_BWD_MAPPING, &m, i); }  // :59
    std::unique_ptr<PHMMTables> backward_by_forward(const std::string& x, const PHMMTables& f) { return bwd(x, DBGPHMM_BWD_BY_FORWARD, nullptr, 0, &f); }  // :101

    // freq.rs:175-192 for every candidate X: returns ln P(R|X) [n_batch]
    std::vector<double> to_full_prob_reads(const Reads& reads, const Mappings* mappings, bool use_max_ratio) {
        std::vector<double> out(dbgphmm_model_n_batch(h_));
        check(dbgphmm_to_full_prob_reads(h_, reads.handle(), mappings ? mappings->handle() : nullptr, use_max_ratio, out.data(), nullptr));
        return out;
    }
    // run / run_sparse / run_sparse_adaptive / run_with_mapping (freq.rs:42-76) + to_node_freqs (freq.rs:245), summed over reads
    std::vector<double> to_node_freqs(const Reads& reads, int run_mode, bool use_max_ratio = true, const Mappings* mappings = nullptr,
                                      std::vector<double>* logp_fwd = nullptr) {
        std::vector<double> f(n_nodes_);
        if (logp_fwd) logp_fwd->resize(0);
        check(dbgphmm_run_node_freqs(h_, reads.handle(), run_mode, use_max_ratio, mappings ? mappings->handle() : nullptr, f.data(), nullptr, nullptr, nullptr));
        return f;
    }
    std::unique_ptr<Mappings> generate_mappings(const Reads& reads, const Mappings* mappings, bool use_max_ratio) {  // hint.rs:193
        dbgphmm_mappings* out = nullptr;
        check(dbgphmm_generate_mappings(h_, reads.handle(), mappings ? mappings->handle() : nullptr, use_max_ratio, &out));
        return std::make_unique<Mappings>(out);
    }
    dbgphmm_model* handle() const { return h_; }

private:
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
    uint32_t n_nodes_;
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
    dbgphmm_dbg* handle() const { return h_; }

private:
    explicit MultiDbg(dbgphmm_dbg* h) : h_(h) { check(dbgphmm_dbg_sizes(h_, sz_)); }
    dbgphmm_dbg* h_;
    uint32_t sz_[6];
};

}  // namespace dbgphmm
