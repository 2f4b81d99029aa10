// formats.cu — DBG and MAP text formats of the reference, host side only (no kernels): the data formats either side of the
// hot path.  A .dbg written by a dbgphmm run becomes the node-centric PHMM graph + copy numbers the device
// path consumes; candidate copy-number vectors X over COMPACT edges (what the posterior sampler of multi_dbg/posterior.rs
// proposes) expand to the per-k-mer copy numbers dbgphmm_model_set_copy_nums_batch takes; Mappings round-trip through .map.
//   DBG  : MultiDbg::to_dbg_writer / from_dbg_reader          multi_dbg/output.rs:155-345   (README.md "DBG")
//   MAP  : MultiDbg::to_map_writer / from_map_reader_raw      multi_dbg/output.rs:455-623
//   graph: MultiDbg::to_node_centric_graph(add_terminal=false) multi_dbg.rs:1551-1604 via to_seq_graph :1370-1390
//   X    : MultiDbg::set_copy_nums                             multi_dbg.rs:1041-1052
// gzip (.dbg.gz / .dbz / .map.gz / .mpz, output.rs:135-139,473-477) goes through zlib.
#include <zlib.h>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cctype>
#include <cmath>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "model.h"
#include "../../include/dbgphmm_b200.h"

#define NULL_BASE 'n'   // common.rs:21

struct dbgphmm_dbg {
    uint32_t k = 0;
    std::vector<std::string> km1mer;                 // compact node -> (k-1)-mer ; node ids == full-graph ids of the same nodes
    struct CEdge { uint32_t s, t; std::string kmer; uint32_t copy_num; std::vector<uint32_t> full; };
    std::vector<CEdge> edges;                        // compact edges in EdgeIndex order
    // full graph (from_dbg_reader, output.rs:255-300): compact nodes first, then one new node per interior base of every compact edge
    uint32_t n_nodes_full = 0;
    std::vector<uint8_t> full_terminal;              // per full node
    std::vector<uint32_t> fsrc, fdst, fcompact;      // per full edge (== PHMM node): endpoints, owning compact edge
    std::vector<uint8_t> fbase;
    // node-centric PHMM edges in the reference's insertion order
    std::vector<uint32_t> psrc, pdst;
};

static bool ends_with(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
static bool is_gz_path(const std::string& p) { return ends_with(p, ".gz") || ends_with(p, ".dbz") || ends_with(p, ".mpz"); }

static int read_file(const char* path, std::string* out) {
    out->clear();
    gzFile f = gzopen(path, "rb");   // transparently reads plain files as well
    if (!f) { dbg_set_error(std::string("cannot open ") + path); return DBGPHMM_ERR_INVALID; }
    char buf[1 << 16];
    int n;
    while ((n = gzread(f, buf, sizeof(buf))) > 0) out->append(buf, (size_t)n);
    gzclose(f);
    if (n < 0) { dbg_set_error(std::string("read error in ") + path); return DBGPHMM_ERR_INVALID; }
    return DBGPHMM_OK;
}
static int write_file(const char* path, const std::string& text) {
    if (is_gz_path(path)) {
        gzFile f = gzopen(path, "wb");
        if (!f) { dbg_set_error(std::string("cannot create ") + path); return DBGPHMM_ERR_INVALID; }
        size_t off = 0;
        while (off < text.size()) {
            int n = gzwrite(f, text.data() + off, (unsigned)std::min<size_t>(text.size() - off, 1u << 30));
            if (n <= 0) { gzclose(f); dbg_set_error(std::string("write error in ") + path); return DBGPHMM_ERR_INVALID; }
            off += (size_t)n;
        }
        if (gzclose(f) != Z_OK) { dbg_set_error(std::string("write error in ") + path); return DBGPHMM_ERR_INVALID; }   // (the last block is flushed here)
        return DBGPHMM_OK;
    }
    FILE* f = fopen(path, "wb");
    if (!f) { dbg_set_error(std::string("cannot create ") + path); return DBGPHMM_ERR_INVALID; }
    size_t w = fwrite(text.data(), 1, text.size(), f);
    const int rc = fclose(f);
    if (w != text.size() || rc != 0) { dbg_set_error(std::string("write error in ") + path); return DBGPHMM_ERR_INVALID; }
    return DBGPHMM_OK;
}

// whitespace-separated fields of one line (split_whitespace, output.rs:211,216,...)
static void split_ws(const std::string& line, std::vector<std::string>* out) {
    out->clear();
    size_t i = 0, n = line.size();
    while (i < n) {
        while (i < n && isspace((unsigned char)line[i])) i++;
        size_t j = i;
        while (j < n && !isspace((unsigned char)line[j])) j++;
        if (j > i) out->push_back(line.substr(i, j - i));
        i = j;
    }
}
static bool parse_u32(const std::string& s, uint32_t* v) {
    if (s.empty()) return false;
    char* e = nullptr;
    unsigned long long x = strtoull(s.c_str(), &e, 10);
    if (*e || x > 0xffffffffull) return false;
    *v = (uint32_t)x;
    return true;
}

static int dbg_finish(dbgphmm_dbg* d) {
    const uint32_t nc = (uint32_t)d->km1mer.size();
    uint64_t n_bases = 0;
    for (auto& e : d->edges) n_bases += e.full.size();
    if (n_bases >= 0xffffffffull) { dbg_set_error("dbg: too many k-mers"); return DBGPHMM_ERR_INVALID; }
    d->full_terminal.clear();
    for (auto& s : d->km1mer) {
        bool term = true;
        for (char c : s) term = term && c == NULL_BASE;
        d->full_terminal.push_back(term ? 1 : 0);
    }
    const uint32_t NE = (uint32_t)n_bases;
    d->fsrc.assign(NE, 0xffffffffu); d->fdst.assign(NE, 0); d->fbase.assign(NE, 0); d->fcompact.assign(NE, 0);
    uint32_t next_node = nc;
    for (uint32_t ce = 0; ce < d->edges.size(); ce++) {
        auto& e = d->edges[ce];
        if (e.s >= nc || e.t >= nc) { dbg_set_error("dbg: edge endpoint is not a node"); return DBGPHMM_ERR_INVALID; }
        const size_t n = e.full.size();
        const std::string seq = e.kmer.substr(d->k - 1);
        uint32_t prev = e.s;
        for (size_t i = 0; i < n; i++) {
            const uint32_t fe = e.full[i];
            if (fe >= NE || d->fsrc[fe] != 0xffffffffu) { dbg_set_error("dbg: index of edge in full is wrong"); return DBGPHMM_ERR_INVALID; }
            const uint32_t v = prev;
            uint32_t w;
            if (i == n - 1) w = e.t; else { w = next_node++; d->full_terminal.push_back(0); }
            d->fsrc[fe] = v; d->fdst[fe] = w; d->fbase[fe] = (uint8_t)seq[i]; d->fcompact[fe] = ce;
            prev = w;
        }
    }
    d->n_nodes_full = next_node;
    // node-centric graph: PHMM node = full edge ; for every non-terminal full node in index order, parents x children nested,
    // each adjacency list newest-edge-first (petgraph 0.6 ; edges were added in index order, output.rs:297-301)
    std::vector<std::vector<uint32_t>> in_e(d->n_nodes_full), out_e(d->n_nodes_full);
    for (uint32_t fe = NE; fe-- > 0;) { out_e[d->fsrc[fe]].push_back(fe); in_e[d->fdst[fe]].push_back(fe); }
    d->psrc.clear(); d->pdst.clear();
    for (uint32_t v = 0; v < d->n_nodes_full; v++) {
        if (d->full_terminal[v]) continue;
        for (uint32_t e1 : in_e[v]) for (uint32_t e2 : out_e[v]) { d->psrc.push_back(e1); d->pdst.push_back(e2); }
    }
    return DBGPHMM_OK;
}

extern "C" int dbgphmm_dbg_from_text(const char* text, uint64_t len, dbgphmm_dbg** out) try {
    if (!text || !out) { dbg_set_error("dbg_from_text: bad argument"); return DBGPHMM_ERR_INVALID; }
    dbgphmm_dbg* d = new dbgphmm_dbg();
    std::vector<std::string> f;
    bool have_k = false;
    size_t pos = 0;
    int st = DBGPHMM_OK;
    while (pos < len && st == DBGPHMM_OK) {
        size_t eol = pos;
        while (eol < len && text[eol] != '\n') eol++;
        std::string line(text + pos, eol - pos);
        pos = eol + 1;
        if (line.empty()) continue;   // (the reference unwraps chars().nth(0) and panics on an empty line ; tolerated here)
        const char c0 = line[0];
        if (c0 == 'K') {
            split_ws(line, &f);
            if (f.size() < 2 || !parse_u32(f[1], &d->k) || d->k < 2) { dbg_set_error("dbg: bad K line"); st = DBGPHMM_ERR_INVALID; }
            have_k = true;
        } else if (c0 == 'N') {
            split_ws(line, &f);
            uint32_t id;
            if (f.size() < 3 || !parse_u32(f[1], &id)) { dbg_set_error("dbg: bad N line"); st = DBGPHMM_ERR_INVALID; break; }
            if (id != d->km1mer.size()) { dbg_set_error("dbg: node is not sorted"); st = DBGPHMM_ERR_INVALID; break; }
            d->km1mer.push_back(f[2]);
        } else if (c0 == 'E') {
            if (!have_k) { dbg_set_error("dbg: E line before K"); st = DBGPHMM_ERR_INVALID; break; }
            split_ws(line, &f);
            dbgphmm_dbg::CEdge e;
            uint32_t id;
            if (f.size() < 7 || !parse_u32(f[1], &id) || !parse_u32(f[2], &e.s) || !parse_u32(f[3], &e.t) || !parse_u32(f[5], &e.copy_num)) {
                dbg_set_error("dbg: bad E line"); st = DBGPHMM_ERR_INVALID; break;
            }
            if (id != d->edges.size()) { dbg_set_error("dbg: edge is not sorted"); st = DBGPHMM_ERR_INVALID; break; }
            e.kmer = f[4];
            size_t p = 0;
            const std::string& lst = f[6];
            while (p <= lst.size()) {
                size_t q = lst.find(',', p);
                if (q == std::string::npos) q = lst.size();
                uint32_t v;
                if (!parse_u32(lst.substr(p, q - p), &v)) { dbg_set_error("dbg: bad edge id list"); st = DBGPHMM_ERR_INVALID; break; }
                e.full.push_back(v);
                p = q + 1;
            }
            if (st != DBGPHMM_OK) break;
            if (e.kmer.size() < d->k || e.kmer.size() - (d->k - 1) != e.full.size()) {
                dbg_set_error("dbg: length of seq and edges_in_full is different"); st = DBGPHMM_ERR_INVALID; break;
            }
            d->edges.push_back(std::move(e));
        }   // '#' and anything else: ignored (output.rs:250-251)
    }
    if (st == DBGPHMM_OK && !have_k) { dbg_set_error("dbg: no K section"); st = DBGPHMM_ERR_INVALID; }
    if (st == DBGPHMM_OK) st = dbg_finish(d);
    if (st != DBGPHMM_OK) { delete d; return st; }
    *out = d;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_dbg_from_file(const char* path, dbgphmm_dbg** out) try {
    if (!path || !out) { dbg_set_error("dbg_from_file: bad argument"); return DBGPHMM_ERR_INVALID; }
    std::string text;
    ST_TRY(read_file(path, &text));
    return dbgphmm_dbg_from_text(text.data(), text.size(), out);
} ABI_CATCH
extern "C" void dbgphmm_dbg_destroy(dbgphmm_dbg* d) { delete d; }

extern "C" int dbgphmm_dbg_sizes(const dbgphmm_dbg* d, uint32_t sizes[6]) try {
    if (!d || !sizes) { dbg_set_error("dbg_sizes: bad argument"); return DBGPHMM_ERR_INVALID; }
    sizes[0] = d->k; sizes[1] = d->n_nodes_full; sizes[2] = (uint32_t)d->fsrc.size(); sizes[3] = (uint32_t)d->km1mer.size();
    sizes[4] = (uint32_t)d->edges.size(); sizes[5] = (uint32_t)d->psrc.size();
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_dbg_phmm_graph(const dbgphmm_dbg* d, uint32_t* edge_src, uint32_t* edge_dst, uint8_t* emission, uint32_t* copy_nums,
                                      uint32_t* compact_edge_of) try {
    if (!d) { dbg_set_error("dbg_phmm_graph: bad argument"); return DBGPHMM_ERR_INVALID; }
    if (edge_src) memcpy(edge_src, d->psrc.data(), 4 * d->psrc.size());
    if (edge_dst) memcpy(edge_dst, d->pdst.data(), 4 * d->pdst.size());
    if (emission) memcpy(emission, d->fbase.data(), d->fbase.size());
    if (copy_nums) for (size_t e = 0; e < d->fcompact.size(); e++) copy_nums[e] = d->edges[d->fcompact[e]].copy_num;
    if (compact_edge_of) memcpy(compact_edge_of, d->fcompact.data(), 4 * d->fcompact.size());
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_dbg_get_copy_nums(const dbgphmm_dbg* d, uint32_t* compact_copy_nums) try {
    if (!d || !compact_copy_nums) { dbg_set_error("dbg_get_copy_nums: bad argument"); return DBGPHMM_ERR_INVALID; }
    for (size_t e = 0; e < d->edges.size(); e++) compact_copy_nums[e] = d->edges[e].copy_num;
    return DBGPHMM_OK;
} ABI_CATCH
// MultiDbg::is_copy_nums_valid (multi_dbg.rs:1008-1014): copy numbers in == out at every node (interior nodes of a compact edge
// are balanced by construction, so the compact nodes decide)
static bool copy_nums_valid(const dbgphmm_dbg* d, const uint32_t* x) {
    std::vector<long long> bal(d->km1mer.size(), 0);
    for (size_t e = 0; e < d->edges.size(); e++) { bal[d->edges[e].t] += x[e]; bal[d->edges[e].s] -= x[e]; }
    for (long long b : bal) if (b != 0) return false;
    return true;
}
extern "C" int dbgphmm_dbg_set_copy_nums(dbgphmm_dbg* d, const uint32_t* compact_copy_nums) try {
    if (!d || !compact_copy_nums) { dbg_set_error("dbg_set_copy_nums: bad argument"); return DBGPHMM_ERR_INVALID; }
    if (!copy_nums_valid(d, compact_copy_nums)) { dbg_set_error("invalid new copy_nums"); return DBGPHMM_ERR_INVALID; }   // multi_dbg.rs:1051
    for (size_t e = 0; e < d->edges.size(); e++) d->edges[e].copy_num = compact_copy_nums[e];
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_dbg_expand_copy_nums(const dbgphmm_dbg* d, uint32_t n_batch, const uint32_t* compact, uint32_t* full) try {
    if (!d || !compact || !full) { dbg_set_error("dbg_expand_copy_nums: bad argument"); return DBGPHMM_ERR_INVALID; }
    const size_t Ec = d->edges.size(), N = d->fcompact.size();
    for (uint32_t b = 0; b < n_batch; b++)
        for (size_t e = 0; e < N; e++) full[(size_t)b * N + e] = compact[(size_t)b * Ec + d->fcompact[e]];
    return DBGPHMM_OK;
} ABI_CATCH

// MultiDbg::genome_size (multi_dbg.rs:1018-1028): copies of every k-mer whose last base is not the null base, for the current copy
// numbers (compact == NULL, n_batch = 1) or for a batch of candidates over compact edges
extern "C" int dbgphmm_dbg_genome_size(const dbgphmm_dbg* d, uint32_t n_batch, const uint32_t* compact, uint64_t* out) try {
    if (!d || !out || (!compact && n_batch != 1)) { dbg_set_error("dbg_genome_size: bad argument"); return DBGPHMM_ERR_INVALID; }
    const size_t Ec = d->edges.size();
    std::vector<uint64_t> emitting(Ec, 0);   // non-null bases along each compact edge
    for (size_t e = 0; e < d->fcompact.size(); e++) if (d->fbase[e] != NULL_BASE) emitting[d->fcompact[e]]++;
    for (uint32_t b = 0; b < n_batch; b++) {
        uint64_t g = 0;
        for (size_t e = 0; e < Ec; e++) g += emitting[e] * (uint64_t)(compact ? compact[(size_t)b * Ec + e] : d->edges[e].copy_num);
        out[b] = g;
    }
    return DBGPHMM_OK;
} ABI_CATCH
// MultiDbg::n_euler_circuits (multi_dbg.rs:831-837): ln of the number of Euler circuits of the compact graph with the copy numbers
// as multiplicities, separate components not allowed.  Candidates must balance at every node, as set_copy_nums asserts.
extern "C" int dbgphmm_dbg_n_euler_circuits(const dbgphmm_dbg* d, uint32_t n_batch, const uint32_t* compact, double* out) try {
    if (!d || !out || (!compact && n_batch != 1)) { dbg_set_error("dbg_n_euler_circuits: bad argument"); return DBGPHMM_ERR_INVALID; }
    const size_t Ec = d->edges.size();
    std::vector<uint32_t> s(Ec), t(Ec), w(Ec);
    for (size_t e = 0; e < Ec; e++) { s[e] = d->edges[e].s; t[e] = d->edges[e].t; }
    for (uint32_t b = 0; b < n_batch; b++) {
        for (size_t e = 0; e < Ec; e++) w[e] = compact ? compact[(size_t)b * Ec + e] : d->edges[e].copy_num;
        if (!copy_nums_valid(d, w.data())) { dbg_set_error("invalid copy_nums (candidate " + std::to_string(b) + ")"); return DBGPHMM_ERR_INVALID; }
        ST_TRY(dbgphmm_euler_circuit_count((uint32_t)d->km1mer.size(), Ec, s.data(), t.data(), w.data(), 0, &out[b]));
    }
    return DBGPHMM_OK;
} ABI_CATCH

static int copy_out(const std::string& s, char* buf, uint64_t cap, uint64_t* needed) {
    if (needed) *needed = s.size();
    if (buf && cap >= s.size()) memcpy(buf, s.data(), s.size());
    else if (buf) { dbg_set_error("output buffer too small"); return DBGPHMM_ERR_INVALID; }
    return DBGPHMM_OK;
}
static std::string dbg_text(const dbgphmm_dbg* d) {
    std::string s;
    s += "# dbgphmm_b200\n";   // (the reference writes its git hash and degree statistics here ; readers skip '#' lines)
    s += "K\t" + std::to_string(d->k) + "\n";
    for (size_t v = 0; v < d->km1mer.size(); v++) s += "N\t" + std::to_string(v) + "\t" + d->km1mer[v] + "\n";
    for (size_t e = 0; e < d->edges.size(); e++) {
        const auto& ce = d->edges[e];
        s += "E\t" + std::to_string(e) + "\t" + std::to_string(ce.s) + "\t" + std::to_string(ce.t) + "\t" + ce.kmer + "\t" + std::to_string(ce.copy_num) + "\t";
        for (size_t i = 0; i < ce.full.size(); i++) { if (i) s += ","; s += std::to_string(ce.full[i]); }
        s += "\n";
    }
    return s;
}
extern "C" int dbgphmm_dbg_to_text(const dbgphmm_dbg* d, char* buf, uint64_t cap, uint64_t* needed) try {
    if (!d) { dbg_set_error("dbg_to_text: bad argument"); return DBGPHMM_ERR_INVALID; }
    return copy_out(dbg_text(d), buf, cap, needed);
} ABI_CATCH
extern "C" int dbgphmm_dbg_to_file(const dbgphmm_dbg* d, const char* path) try {
    if (!d || !path) { dbg_set_error("dbg_to_file: bad argument"); return DBGPHMM_ERR_INVALID; }
    return write_file(path, dbg_text(d));
} ABI_CATCH

// MultiDbg::to_phmm / to_non_zero_phmm / to_uniform_phmm (multi_dbg.rs:1391-1409): n_warmup := k, probabilities from the copy numbers
extern "C" int dbgphmm_dbg_to_model(const dbgphmm_dbg* d, const dbgphmm_params* params, int mode, int device, uint64_t mem_budget_bytes, dbgphmm_model** out) try {
    if (!d || !params || !out || mode < 0 || mode > 2) { dbg_set_error("dbg_to_model: bad argument"); return DBGPHMM_ERR_INVALID; }
    dbgphmm_params p = *params;
    p.n_warmup = d->k;
    const size_t N = d->fsrc.size(), E = d->psrc.size();
    std::vector<double> li(N, 0.0), lt(E, 0.0);   // placeholders ; replaced by the copy-number derivation below
    dbgphmm_model* m = nullptr;
    ST_TRY(dbgphmm_model_create((uint32_t)N, (uint32_t)E, d->psrc.data(), d->pdst.data(), d->fbase.data(), li.data(), lt.data(), &p, device, mem_budget_bytes, &m));
    std::vector<uint32_t> cn(N);
    for (size_t e = 0; e < N; e++) cn[e] = d->edges[d->fcompact[e]].copy_num;
    int st = dbgphmm_model_set_copy_nums_batch(m, 1, cn.data(), mode);
    if (st != DBGPHMM_OK) { dbgphmm_model_destroy(m); return st; }
    *out = m;
    return DBGPHMM_OK;
} ABI_CATCH

// ------------------------------------------------------------------------------------------------ MAP
// f64 as Rust's `{}` prints it: shortest digits that round-trip, never in exponent notation ; -inf / inf / NaN by name
static void append_f64(std::string& s, double v) {
    if (v != v) { s += "NaN"; return; }
    if (v == INFINITY) { s += "inf"; return; }
    if (v == -INFINITY) { s += "-inf"; return; }
    char buf[800];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    s.append(buf, r.ptr);
}
static bool parse_f64(const std::string& t, double* v) {
    if (t == "-inf") { *v = -INFINITY; return true; }
    if (t == "inf") { *v = INFINITY; return true; }
    if (t == "NaN") { *v = NAN; return true; }
    char* e = nullptr;
    *v = strtod(t.c_str(), &e);
    return !t.empty() && *e == 0;
}

extern "C" int dbgphmm_mappings_from_map_text(const char* text, uint64_t len, dbgphmm_mappings** out) try {
    if (!text || !out) { dbg_set_error("mappings_from_map_text: bad argument"); return DBGPHMM_ERR_INVALID; }
    dbgphmm_mappings* mp = new dbgphmm_mappings();
    mp->read_off.push_back(0); mp->row_off.push_back(0);
    std::vector<std::string> f;
    uint64_t n_reads = 0, rows_in_read = 0;
    size_t pos = 0;
    auto fail = [&](const char* msg) { dbg_set_error(msg); delete mp; return DBGPHMM_ERR_INVALID; };
    while (pos < len) {
        size_t eol = pos;
        while (eol < len && text[eol] != '\n') eol++;
        std::string line(text + pos, eol - pos);
        pos = eol + 1;
        if (line.empty() || line[0] == '#') continue;
        split_ws(line, &f);
        uint32_t i, j;
        if (f.size() < 3 || !parse_u32(f[0], &i) || !parse_u32(f[1], &j)) return fail("map: bad line");
        // rows arrive read by read, base by base (the reference asserts ret.len() == i + 1 and ret[i].len() == j + 1, output.rs:551-552)
        if (i == n_reads) { if (n_reads) mp->read_off.push_back(mp->row_off.size() - 1); n_reads++; rows_in_read = 0; }
        if (i + 1 != n_reads || j != rows_in_read) return fail("map: rows are not in (read, position) order");
        rows_in_read++;
        if (f.size() >= 4) {
            const std::string& lst = f[3];
            size_t p = 0;
            while (p <= lst.size()) {
                size_t q = lst.find(',', p);
                if (q == std::string::npos) q = lst.size();
                const std::string item = lst.substr(p, q - p);
                const size_t c = item.find(':');
                uint32_t node; double lp;
                if (c == std::string::npos || !parse_u32(item.substr(0, c), &node) || !parse_f64(item.substr(c + 1), &lp)) return fail("map: bad node:prob item");
                mp->nodes.push_back(node); mp->logp.push_back(lp);
                p = q + 1;
            }
        }
        mp->row_off.push_back(mp->nodes.size());
    }
    mp->read_off.push_back(mp->row_off.size() - 1);
    if (n_reads == 0) mp->read_off.assign(1, 0);
    *out = mp;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_mappings_from_map_file(const char* path, dbgphmm_mappings** out) try {
    if (!path || !out) { dbg_set_error("mappings_from_map_file: bad argument"); return DBGPHMM_ERR_INVALID; }
    std::string text;
    ST_TRY(read_file(path, &text));
    return dbgphmm_mappings_from_map_text(text.data(), text.size(), out);
} ABI_CATCH
static int map_text(const dbgphmm_mappings* mp, const dbgphmm_reads* reads, const dbgphmm_dbg* d, std::string* out) {
    if (mp->read_off.size() != reads->n_reads + 1) { dbg_set_error("mappings / reads count mismatch"); return DBGPHMM_ERR_INVALID; }
    std::string& s = *out;
    s += "# dbgphmm_b200\n";
    if (d) s += "# k=" + std::to_string(d->k) + " n_edges_full=" + std::to_string(d->fsrc.size()) + " n_edges_compact=" + std::to_string(d->edges.size()) + "\n";
    s += "# read\tpos\tbase\tnodes_and_probs\n";
    for (uint64_t i = 0; i < reads->n_reads; i++) {
        const uint64_t n = reads->off[i + 1] - reads->off[i];
        if (mp->read_off[i + 1] - mp->read_off[i] != n) { dbg_set_error("mapping length differs from read length"); return DBGPHMM_ERR_INVALID; }
        s += "# i=" + std::to_string(i) + "\n";
        for (uint64_t j = 0; j < n; j++) {
            const uint64_t row = mp->read_off[i] + j;
            s += std::to_string(i) + "\t" + std::to_string(j) + "\t";
            s += (char)reads->bases[reads->off[i] + j];
            s += "\t";
            for (uint64_t e = mp->row_off[row]; e < mp->row_off[row + 1]; e++) {
                if (e > mp->row_off[row]) s += ",";
                s += std::to_string(mp->nodes[e]) + ":";
                append_f64(s, mp->logp[e]);
            }
            s += "\n";
        }
    }
    return DBGPHMM_OK;
}
extern "C" int dbgphmm_mappings_to_map_text(const dbgphmm_mappings* mp, const dbgphmm_reads* reads, const dbgphmm_dbg* d, char* buf, uint64_t cap, uint64_t* needed) try {
    if (!mp || !reads) { dbg_set_error("mappings_to_map_text: bad argument"); return DBGPHMM_ERR_INVALID; }
    std::string s;
    ST_TRY(map_text(mp, reads, d, &s));
    return copy_out(s, buf, cap, needed);
} ABI_CATCH
extern "C" int dbgphmm_mappings_to_map_file(const dbgphmm_mappings* mp, const dbgphmm_reads* reads, const dbgphmm_dbg* d, const char* path) try {
    if (!mp || !reads || !path) { dbg_set_error("mappings_to_map_file: bad argument"); return DBGPHMM_ERR_INVALID; }
    std::string s;
    ST_TRY(map_text(mp, reads, d, &s));
    return write_file(path, s);
} ABI_CATCH

// ------------------------------------------------------------------------------------------------ dataset JSON
// Dataset::to_json_file / from_json_file (e2e.rs:31-52,123-130): serde_json of
//   { "genome": ["L:ACGT...", ...]                      Genome(Vec<StyledSequence>), Display "{style}:{bases}"  collection.rs:371-379,460-464
//     "genome_size": usize,
//     "reads": { "reads": ["ACGT:+:0-12,0-13,I,0-15", ...] }   PositionedReads: "{bases}:{+|-}:{origins}", an origin is
//                                                               "{hap}-{pos}" or "I"  (collection.rs:711-757, genome_graph.rs:117-151)
//     "phmm_params": { "p_mismatch": "-6.907755278982137(0.0010)", ..., "n_active_nodes": 40, "active_node_max_ratio": 30.0,
//                      "n_warmup": 50, "warmup_threshold": 200, "n_max_gaps": 4 } }   Prob Display "{ln p}({p:.4})"  prob.rs:158-169
// The reads of a dataset are what the hot path consumes (dbgphmm_dataset_reads -> dbgphmm_reads); genome, origins and parameters travel
// with them so that a dataset written by a dbgphmm run feeds the B200 path and a dataset written here loads in the reference.
struct dbgphmm_dataset {
    std::vector<std::string> hap; std::vector<char> style;
    uint64_t genome_size = 0;
    std::vector<std::string> read; std::vector<uint8_t> revcomp;
    std::vector<std::vector<int64_t>> onode; std::vector<std::vector<uint64_t>> opos;   // per base: haplotype (-1 = Ins) and position
    dbgphmm_params params;
};

namespace {
struct JVal {   // just enough JSON for the schema above
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    double num = 0; bool b = false; std::string str; std::string raw;
    std::vector<JVal> arr; std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* k) const { for (auto& kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
};
struct JParser {
    const char* p; const char* e; std::string err;
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool fail(const char* m) { if (err.empty()) err = m; return false; }
    bool str(std::string* out) {
        if (p >= e || *p != '"') return fail("expected a string");
        p++; out->clear();
        while (p < e && *p != '"') {
            if (*p == '\\') {
                if (++p >= e) return fail("bad escape");
                switch (*p) {
                    case 'n': *out += '\n'; break; case 't': *out += '\t'; break; case 'r': *out += '\r'; break;
                    case 'b': *out += '\b'; break; case 'f': *out += '\f'; break;
                    case 'u': { if (e - p < 5) return fail("bad \\u escape"); unsigned c = (unsigned)strtoul(std::string(p + 1, 4).c_str(), nullptr, 16); *out += (char)(c & 0x7f); p += 4; break; }
                    default: *out += *p;
                }
                p++;
            } else *out += *p++;
        }
        if (p >= e) return fail("unterminated string");
        p++;
        return true;
    }
    bool val(JVal* v, int depth = 0) {
        if (depth > 32) return fail("nesting too deep");
        ws();
        if (p >= e) return fail("unexpected end");
        if (*p == '{') {
            v->kind = JVal::OBJ; p++; ws();
            if (p < e && *p == '}') { p++; return true; }
            for (;;) {
                ws(); std::string k; if (!str(&k)) return false;
                ws(); if (p >= e || *p != ':') return fail("expected ':'"); p++;
                JVal c; if (!val(&c, depth + 1)) return false;
                v->obj.emplace_back(std::move(k), std::move(c));
                ws(); if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == '}') { p++; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v->kind = JVal::ARR; p++; ws();
            if (p < e && *p == ']') { p++; return true; }
            for (;;) {
                JVal c; if (!val(&c, depth + 1)) return false;
                v->arr.push_back(std::move(c));
                ws(); if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == ']') { p++; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') { v->kind = JVal::STR; return str(&v->str); }
        if (!strncmp(p, "true", std::min<size_t>(4, e - p)) && e - p >= 4) { v->kind = JVal::BOOL; v->b = true; p += 4; return true; }
        if (!strncmp(p, "false", std::min<size_t>(5, e - p)) && e - p >= 5) { v->kind = JVal::BOOL; v->b = false; p += 5; return true; }
        if (!strncmp(p, "null", std::min<size_t>(4, e - p)) && e - p >= 4) { v->kind = JVal::NUL; p += 4; return true; }
        const char* q = p;
        while (q < e && (isdigit((unsigned char)*q) || *q == '-' || *q == '+' || *q == '.' || *q == 'e' || *q == 'E')) q++;
        if (q == p) return fail("unexpected character");
        v->kind = JVal::NUM; v->raw.assign(p, q);
        char* end = nullptr; v->num = strtod(v->raw.c_str(), &end);
        if (*end) return fail("bad number");
        p = q;
        return true;
    }
};
bool parse_u64(const std::string& s, uint64_t* v) {
    if (s.empty() || s.size() > 20) return false;
    for (char c : s) if (!isdigit((unsigned char)c)) return false;
    *v = strtoull(s.c_str(), nullptr, 10);
    return true;
}
// Prob: "{ln p}({p:.4})" (prob.rs:158-169) ; a bare number is accepted too
bool parse_prob(const JVal& v, double* out) {
    if (v.kind == JVal::NUM) { *out = v.num; return true; }
    if (v.kind != JVal::STR) return false;
    const size_t par = v.str.find('(');
    return parse_f64(par == std::string::npos ? v.str : v.str.substr(0, par), out);
}
bool json_u64(const JVal* v, uint64_t* out) { return v && v->kind == JVal::NUM && parse_u64(v->raw, out); }
void json_str(std::string& s, const std::string& t) {
    s += '"';
    for (char c : t) { if (c == '"' || c == '\\') s += '\\'; s += c; }
    s += '"';
}
int dataset_parse(const char* text, uint64_t len, dbgphmm_dataset** out) {
    if (!text || !out) { dbg_set_error("dataset_from_json: bad argument"); return DBGPHMM_ERR_INVALID; }
    JParser P{text, text + len, ""};
    JVal root;
    if (!P.val(&root) || root.kind != JVal::OBJ) { dbg_set_error("dataset JSON: " + (P.err.empty() ? std::string("not an object") : P.err)); return DBGPHMM_ERR_INVALID; }
    P.ws();
    if (P.p != P.e) { dbg_set_error("dataset JSON: trailing characters"); return DBGPHMM_ERR_INVALID; }
    std::unique_ptr<dbgphmm_dataset> d(new dbgphmm_dataset());
    const JVal* g = root.get("genome");
    if (!g || g->kind != JVal::ARR) { dbg_set_error("dataset JSON: genome must be an array of styled sequences"); return DBGPHMM_ERR_INVALID; }
    for (auto& h : g->arr) {
        if (h.kind != JVal::STR || h.str.size() < 2 || h.str[1] != ':' || !strchr("CLF", h.str[0])) { dbg_set_error("dataset JSON: a styled sequence is \"C|L|F:bases\" (collection.rs:371-391)"); return DBGPHMM_ERR_INVALID; }
        d->style.push_back(h.str[0]); d->hap.push_back(h.str.substr(2));
    }
    if (!json_u64(root.get("genome_size"), &d->genome_size)) { dbg_set_error("dataset JSON: genome_size"); return DBGPHMM_ERR_INVALID; }
    const JVal* rc = root.get("reads");
    const JVal* rs = rc && rc->kind == JVal::OBJ ? rc->get("reads") : nullptr;
    if (!rs || rs->kind != JVal::ARR) { dbg_set_error("dataset JSON: reads.reads must be an array"); return DBGPHMM_ERR_INVALID; }
    for (auto& r : rs->arr) {
        if (r.kind != JVal::STR) { dbg_set_error("dataset JSON: a read is a string"); return DBGPHMM_ERR_INVALID; }
        // PositionedSequence "{bases}:{+|-}:{origins}" ; a plain sequence (Reads, collection.rs:811-819) is accepted without origins
        const size_t c1 = r.str.find(':');
        std::string bases = c1 == std::string::npos ? r.str : r.str.substr(0, c1);
        for (char b : bases) if (b != 'A' && b != 'C' && b != 'G' && b != 'T') { dbg_set_error("dataset JSON: read bases must be uppercase ACGT (collection.rs:236-249)"); return DBGPHMM_ERR_INVALID; }
        std::vector<int64_t> on; std::vector<uint64_t> op; uint8_t rev = 0;
        if (c1 != std::string::npos) {
            const size_t c2 = r.str.find(':', c1 + 1);
            if (c2 == std::string::npos || c2 != c1 + 2 || (r.str[c1 + 1] != '+' && r.str[c1 + 1] != '-')) { dbg_set_error("dataset JSON: a positioned read is \"bases:+|-:origins\" (collection.rs:732-757)"); return DBGPHMM_ERR_INVALID; }
            rev = r.str[c1 + 1] == '-';
            size_t a = c2 + 1;
            while (a <= r.str.size() && !(a == r.str.size() && bases.empty())) {
                size_t b = r.str.find(',', a); if (b == std::string::npos) b = r.str.size();
                const std::string o = r.str.substr(a, b - a);
                const size_t dash = o.find('-');
                if (dash == std::string::npos) { on.push_back(-1); op.push_back(0); }   // "I" (anything without a dash parses as Ins, genome_graph.rs:147-149)
                else {
                    uint64_t hn, hp;
                    if (!parse_u64(o.substr(0, dash), &hn) || !parse_u64(o.substr(dash + 1), &hp)) { dbg_set_error("dataset JSON: an origin is \"hap-pos\" or \"I\""); return DBGPHMM_ERR_INVALID; }
                    on.push_back((int64_t)hn); op.push_back(hp);
                }
                if (b == r.str.size()) break;
                a = b + 1;
            }
            if (on.size() != bases.size()) { dbg_set_error("dataset JSON: one origin per base (collection.rs:522 asserts it)"); return DBGPHMM_ERR_INVALID; }
        } else { on.assign(bases.size(), -1); op.assign(bases.size(), 0); }
        d->read.push_back(std::move(bases)); d->revcomp.push_back(rev); d->onode.push_back(std::move(on)); d->opos.push_back(std::move(op));
    }
    const JVal* pp = root.get("phmm_params");
    if (!pp || pp->kind != JVal::OBJ) { dbg_set_error("dataset JSON: phmm_params"); return DBGPHMM_ERR_INVALID; }
    dbgphmm_params& q = d->params;
    struct { const char* name; double* dst; } probs[] = {
        {"p_mismatch", &q.p_mismatch}, {"p_match", &q.p_match}, {"p_random", &q.p_random}, {"p_gap_open", &q.p_gap_open}, {"p_gap_ext", &q.p_gap_ext},
        {"p_end", &q.p_end}, {"p_MM", &q.p_MM}, {"p_IM", &q.p_IM}, {"p_DM", &q.p_DM}, {"p_MI", &q.p_MI}, {"p_II", &q.p_II}, {"p_DI", &q.p_DI},
        {"p_MD", &q.p_MD}, {"p_ID", &q.p_ID}, {"p_DD", &q.p_DD}};
    for (auto& f : probs) {
        const JVal* v = pp->get(f.name);
        if (!v || !parse_prob(*v, f.dst)) { dbg_set_error(std::string("dataset JSON: phmm_params.") + f.name); return DBGPHMM_ERR_INVALID; }
    }
    uint64_t u = 0;
    if (!json_u64(pp->get("n_active_nodes"), &u)) { dbg_set_error("dataset JSON: phmm_params.n_active_nodes"); return DBGPHMM_ERR_INVALID; } q.n_active_nodes = (uint32_t)u;
    if (!json_u64(pp->get("n_warmup"), &u)) { dbg_set_error("dataset JSON: phmm_params.n_warmup"); return DBGPHMM_ERR_INVALID; } q.n_warmup = (uint32_t)u;
    if (!json_u64(pp->get("n_max_gaps"), &u)) { dbg_set_error("dataset JSON: phmm_params.n_max_gaps"); return DBGPHMM_ERR_INVALID; } q.n_max_gaps = (uint32_t)u;
    q.warmup_threshold = DBGPHMM_MAX_ACTIVE_NODES / 2;   // #[serde(default = "default_warmup_threshold")], params.rs:60-67
    if (pp->get("warmup_threshold")) { if (!json_u64(pp->get("warmup_threshold"), &u)) { dbg_set_error("dataset JSON: phmm_params.warmup_threshold"); return DBGPHMM_ERR_INVALID; } q.warmup_threshold = (uint32_t)u; }
    const JVal* ar = pp->get("active_node_max_ratio");
    if (!ar || ar->kind != JVal::NUM) { dbg_set_error("dataset JSON: phmm_params.active_node_max_ratio"); return DBGPHMM_ERR_INVALID; }
    q.active_node_max_ratio = ar->num;
    *out = d.release();
    return DBGPHMM_OK;
}
void append_prob(std::string& s, double lnp) {   // Prob Display: "{}({:.4})" of (ln p, p)
    s += '"'; append_f64(s, lnp);
    char buf[64]; snprintf(buf, sizeof(buf), "(%.4f)", std::exp(lnp));
    s += buf; s += '"';
}
void append_json_f64(std::string& s, double v) {   // serde_json: shortest round-trip digits, always with a fraction or exponent
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v);
    std::string t(buf, r.ptr);
    if (t.find_first_of(".eE") == std::string::npos) t += ".0";
    s += t;
}
std::string dataset_text(const dbgphmm_dataset* d) {
    std::string s = "{\"genome\":[";
    for (size_t i = 0; i < d->hap.size(); i++) { if (i) s += ','; json_str(s, std::string(1, d->style[i]) + ":" + d->hap[i]); }
    s += "],\"genome_size\":" + std::to_string(d->genome_size) + ",\"reads\":{\"reads\":[";
    for (size_t i = 0; i < d->read.size(); i++) {
        if (i) s += ',';
        std::string r = d->read[i] + ":" + (d->revcomp[i] ? "-" : "+") + ":";
        for (size_t j = 0; j < d->onode[i].size(); j++) {
            if (j) r += ',';
            if (d->onode[i][j] < 0) r += 'I'; else r += std::to_string(d->onode[i][j]) + "-" + std::to_string(d->opos[i][j]);
        }
        json_str(s, r);
    }
    s += "]},\"phmm_params\":{";
    const dbgphmm_params& q = d->params;
    const std::pair<const char*, double> probs[] = {{"p_mismatch", q.p_mismatch}, {"p_match", q.p_match}, {"p_random", q.p_random}, {"p_gap_open", q.p_gap_open},
        {"p_gap_ext", q.p_gap_ext}, {"p_end", q.p_end}, {"p_MM", q.p_MM}, {"p_IM", q.p_IM}, {"p_DM", q.p_DM}, {"p_MI", q.p_MI}, {"p_II", q.p_II}, {"p_DI", q.p_DI},
        {"p_MD", q.p_MD}, {"p_ID", q.p_ID}, {"p_DD", q.p_DD}};
    for (auto& f : probs) { s += '"'; s += f.first; s += "\":"; append_prob(s, f.second); s += ','; }
    s += "\"n_active_nodes\":" + std::to_string(q.n_active_nodes) + ",\"active_node_max_ratio\":";
    append_json_f64(s, q.active_node_max_ratio);
    s += ",\"n_warmup\":" + std::to_string(q.n_warmup) + ",\"warmup_threshold\":" + std::to_string(q.warmup_threshold) + ",\"n_max_gaps\":" + std::to_string(q.n_max_gaps) + "}}";
    return s;
}
}  // namespace

extern "C" int dbgphmm_dataset_from_json_text(const char* text, uint64_t len, dbgphmm_dataset** out) try { return dataset_parse(text, len, out); } ABI_CATCH
extern "C" int dbgphmm_dataset_from_json_file(const char* path, dbgphmm_dataset** out) try {
    if (!path || !out) { dbg_set_error("dataset_from_json_file: bad argument"); return DBGPHMM_ERR_INVALID; }
    std::string text;
    ST_TRY(read_file(path, &text));
    return dataset_parse(text.data(), text.size(), out);
} ABI_CATCH
extern "C" void dbgphmm_dataset_destroy(dbgphmm_dataset* d) { delete d; }
extern "C" int dbgphmm_dataset_create(uint32_t n_haps, const uint64_t* hap_off, const uint8_t* hap_bases, const uint8_t* hap_style, uint64_t genome_size,
                                      uint64_t n_reads, const uint64_t* read_off, const uint8_t* read_bases, const uint8_t* read_revcomp, const int64_t* origin_hap,
                                      const uint64_t* origin_pos, const dbgphmm_params* params, dbgphmm_dataset** out) try {
    if (!out || !params || (n_haps && (!hap_off || !hap_style)) || (n_reads && !read_off)) { dbg_set_error("dataset_create: bad argument"); return DBGPHMM_ERR_INVALID; }
    std::unique_ptr<dbgphmm_dataset> d(new dbgphmm_dataset());
    for (uint32_t h = 0; h < n_haps; h++) {
        if (hap_off[h + 1] < hap_off[h] || !strchr("CLF", (char)hap_style[h]) || !hap_style[h]) { dbg_set_error("dataset_create: haplotype offsets / styles ('C', 'L' or 'F')"); return DBGPHMM_ERR_INVALID; }
        d->hap.emplace_back((const char*)hap_bases + hap_off[h], (const char*)hap_bases + hap_off[h + 1]); d->style.push_back((char)hap_style[h]);
    }
    d->genome_size = genome_size;
    for (uint64_t r = 0; r < n_reads; r++) {
        if (read_off[r + 1] < read_off[r]) { dbg_set_error("dataset_create: read offsets not monotone"); return DBGPHMM_ERR_INVALID; }
        const uint64_t a = read_off[r], b = read_off[r + 1];
        std::string bases((const char*)read_bases + a, (const char*)read_bases + b);
        for (char c : bases) if (c != 'A' && c != 'C' && c != 'G' && c != 'T') { dbg_set_error("dataset_create: read bases must be uppercase ACGT"); return DBGPHMM_ERR_INVALID; }
        std::vector<int64_t> on(b - a, -1); std::vector<uint64_t> op(b - a, 0);
        if (origin_hap && origin_pos) for (uint64_t j = a; j < b; j++) { on[j - a] = origin_hap[j] < 0 ? -1 : origin_hap[j]; op[j - a] = origin_hap[j] < 0 ? 0 : origin_pos[j]; }
        d->read.push_back(std::move(bases)); d->revcomp.push_back(read_revcomp ? read_revcomp[r] != 0 : 0); d->onode.push_back(std::move(on)); d->opos.push_back(std::move(op));
    }
    d->params = *params;
    *out = d.release();
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_dataset_sizes(const dbgphmm_dataset* d, uint64_t sizes[5]) try {
    if (!d || !sizes) { dbg_set_error("dataset_sizes: bad argument"); return DBGPHMM_ERR_INVALID; }
    sizes[0] = d->hap.size(); sizes[1] = 0; for (auto& h : d->hap) sizes[1] += h.size();
    sizes[2] = d->read.size(); sizes[3] = 0; for (auto& r : d->read) sizes[3] += r.size();
    sizes[4] = d->genome_size;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_dataset_genome(const dbgphmm_dataset* d, uint64_t* hap_off, uint8_t* bases, uint8_t* style) try {
    if (!d) { dbg_set_error("dataset_genome: bad argument"); return DBGPHMM_ERR_INVALID; }
    uint64_t o = 0;
    for (size_t h = 0; h < d->hap.size(); h++) {
        if (hap_off) hap_off[h] = o;
        if (bases) memcpy(bases + o, d->hap[h].data(), d->hap[h].size());
        if (style) style[h] = (uint8_t)d->style[h];
        o += d->hap[h].size();
    }
    if (hap_off) hap_off[d->hap.size()] = o;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_dataset_reads(const dbgphmm_dataset* d, dbgphmm_reads** out) try {
    if (!d || !out) { dbg_set_error("dataset_reads: bad argument"); return DBGPHMM_ERR_INVALID; }
    std::vector<uint64_t> off(d->read.size() + 1, 0); std::vector<uint8_t> bases;
    for (size_t r = 0; r < d->read.size(); r++) { bases.insert(bases.end(), d->read[r].begin(), d->read[r].end()); off[r + 1] = bases.size(); }
    return dbgphmm_reads_create(d->read.size(), off.data(), bases.empty() ? nullptr : bases.data(), out);
} ABI_CATCH
extern "C" int dbgphmm_dataset_read_origins(const dbgphmm_dataset* d, uint64_t* read_off, uint8_t* bases, uint8_t* revcomp, int64_t* origin_hap, uint64_t* origin_pos) try {
    if (!d) { dbg_set_error("dataset_read_origins: bad argument"); return DBGPHMM_ERR_INVALID; }
    uint64_t o = 0;
    for (size_t r = 0; r < d->read.size(); r++) {
        if (read_off) read_off[r] = o;
        if (revcomp) revcomp[r] = d->revcomp[r];
        for (size_t j = 0; j < d->read[r].size(); j++, o++) {
            if (bases) bases[o] = (uint8_t)d->read[r][j];
            if (origin_hap) origin_hap[o] = d->onode[r][j];
            if (origin_pos) origin_pos[o] = d->opos[r][j];
        }
    }
    if (read_off) read_off[d->read.size()] = o;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_dataset_params(const dbgphmm_dataset* d, dbgphmm_params* out) try {
    if (!d || !out) { dbg_set_error("dataset_params: bad argument"); return DBGPHMM_ERR_INVALID; }
    *out = d->params;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_dataset_to_json_text(const dbgphmm_dataset* d, char* buf, uint64_t cap, uint64_t* needed) try {
    if (!d) { dbg_set_error("dataset_to_json_text: bad argument"); return DBGPHMM_ERR_INVALID; }
    return copy_out(dataset_text(d), buf, cap, needed);
} ABI_CATCH
extern "C" int dbgphmm_dataset_to_json_file(const dbgphmm_dataset* d, const char* path) try {
    if (!d || !path) { dbg_set_error("dataset_to_json_file: bad argument"); return DBGPHMM_ERR_INVALID; }
    return write_file(path, dataset_text(d));
} ABI_CATCH
