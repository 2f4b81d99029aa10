// api.cu — C ABI of include/dbgphmm_b200.h: handles, row export, and the bulk (read-set) entry points.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <unordered_map>
#include <condition_variable>
#include <mutex>
#include <thread>
#include "engine.h"

static const double LN2 = 0.693147180559945309417232121458;

// ------------------------------------------------------------------ reads / mappings handles
extern "C" int dbgphmm_reads_create(uint64_t n_reads, const uint64_t* offsets, const uint8_t* bases, dbgphmm_reads** out) try {
    if (!out || !offsets || (n_reads && offsets[n_reads] && !bases)) { dbg_set_error("reads_create: bad argument"); return DBGPHMM_ERR_INVALID; }
    for (uint64_t r = 0; r < n_reads; r++)
        if (offsets[r + 1] < offsets[r]) { dbg_set_error("reads_create: offsets not monotone"); return DBGPHMM_ERR_INVALID; }
    const uint64_t total = offsets[n_reads];
    for (uint64_t i = 0; i < total; i++) {
        uint8_t b = bases[i];
        if (b != 'A' && b != 'C' && b != 'G' && b != 'T') { dbg_set_error("reads_create: bases must be uppercase ACGT (collection.rs:236-249)"); return DBGPHMM_ERR_INVALID; }
    }
    dbgphmm_reads* r = new dbgphmm_reads();
    r->n_reads = n_reads;
    r->off.assign(offsets, offsets + n_reads + 1);
    r->bases.assign(bases, bases + total);
    *out = r;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" void dbgphmm_reads_destroy(dbgphmm_reads* r) {
    if (!r) return;
    if (r->d_bases) {   // (a cache block: no cudaFree in steady state ; no stream to order the reuse behind -> the next user synchronises)
        cudaSetDevice(r->device);
        cache_set_stream(nullptr);
        cache_free(r->d_bases);
    }
    delete r;
}
extern "C" int dbgphmm_reads_to_device(dbgphmm_model* m, dbgphmm_reads* r) try {
    if (!m || !r) { dbg_set_error("reads_to_device: bad argument"); return DBGPHMM_ERR_INVALID; }
    if (r->d_bases && r->device == m->device) return DBGPHMM_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    if (r->d_bases) { int cur = 0; cudaGetDevice(&cur); cudaSetDevice(r->device); cache_free(r->d_bases); r->d_bases = nullptr; cudaSetDevice(cur); }
    r->d_bases = (uint8_t*)cache_alloc(std::max<size_t>(r->bases.size(), 1));   // (from the block cache: a fresh handle per call costs no cudaMalloc / cudaFree)
    if (!r->d_bases) { dbg_set_error("out of device memory for the reads"); return DBGPHMM_ERR_OOM; }
    if (!r->bases.empty()) {
        CUDA_TRY(cudaMemcpyAsync(r->d_bases, r->bases.data(), r->bases.size(), cudaMemcpyHostToDevice, MSET(m).stream));
        CUDA_TRY(cudaStreamSynchronize(MSET(m).stream));
    }
    r->device = m->device;
    return DBGPHMM_OK;
} ABI_CATCH

extern "C" int dbgphmm_mappings_create(uint64_t n_reads, const uint64_t* read_off, const uint64_t* row_off, const uint32_t* nodes,
                                       const double* logp, dbgphmm_mappings** out) try {
    if (!out || !read_off || !row_off) { dbg_set_error("mappings_create: bad argument"); return DBGPHMM_ERR_INVALID; }
    // the two CSR levels must be well-formed before anything is copied: they are trusted as indices from here on
    if (read_off[0] != 0) { dbg_set_error("mappings_create: read_off[0] must be 0"); return DBGPHMM_ERR_INVALID; }
    for (uint64_t r = 0; r < n_reads; r++)
        if (read_off[r + 1] < read_off[r]) { dbg_set_error("mappings_create: read_off is not non-decreasing"); return DBGPHMM_ERR_INVALID; }
    const uint64_t n_rows = read_off[n_reads];
    if (row_off[0] != 0) { dbg_set_error("mappings_create: row_off[0] must be 0"); return DBGPHMM_ERR_INVALID; }
    for (uint64_t i = 0; i < n_rows; i++)
        if (row_off[i + 1] < row_off[i]) { dbg_set_error("mappings_create: row_off is not non-decreasing"); return DBGPHMM_ERR_INVALID; }
    const uint64_t n_ent = row_off[n_rows];
    if (n_ent && !nodes) { dbg_set_error("mappings_create: nodes is null"); return DBGPHMM_ERR_INVALID; }
    std::unique_ptr<dbgphmm_mappings> mp(new dbgphmm_mappings());
    mp->read_off.assign(read_off, read_off + n_reads + 1);
    mp->row_off.assign(row_off, row_off + n_rows + 1);
    if (n_ent) { mp->nodes.assign(nodes, nodes + n_ent); if (logp) mp->logp.assign(logp, logp + n_ent); else mp->logp.assign(n_ent, 0.0); }
    *out = mp.release();
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" void dbgphmm_mappings_destroy(dbgphmm_mappings* mp) { if (mp) mappings_release_device(mp); delete mp; }
extern "C" int dbgphmm_mappings_sizes(const dbgphmm_mappings* mp, uint64_t* n_reads, uint64_t* n_rows, uint64_t* n_entries) try {
    if (!mp) { dbg_set_error("mappings_sizes: null"); return DBGPHMM_ERR_INVALID; }
    if (n_reads) *n_reads = mp->read_off.empty() ? 0 : mp->read_off.size() - 1;
    if (n_rows) *n_rows = mp->row_off.empty() ? 0 : mp->row_off.size() - 1;
    if (n_entries) *n_entries = mp->nodes.size();
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_mappings_export(const dbgphmm_mappings* mp, uint64_t* read_off, uint64_t* row_off, uint32_t* nodes, double* logp) try {
    if (!mp) { dbg_set_error("mappings_export: null"); return DBGPHMM_ERR_INVALID; }
    std::memcpy(read_off, mp->read_off.data(), 8 * mp->read_off.size());
    std::memcpy(row_off, mp->row_off.data(), 8 * mp->row_off.size());
    if (!mp->nodes.empty()) { std::memcpy(nodes, mp->nodes.data(), 4 * mp->nodes.size()); std::memcpy(logp, mp->logp.data(), 8 * mp->logp.size()); }
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_mappings_to_node_freqs(const dbgphmm_mappings* mp, uint32_t n_nodes, double* freqs) try {
    if (!mp || !freqs) { dbg_set_error("mappings_to_node_freqs: bad argument"); return DBGPHMM_ERR_INVALID; }
    // hint.rs:161-171 — a plain host-side scatter of exp(logp); not on the DP path, no device work involved
    std::fill(freqs, freqs + n_nodes, 0.0);
    for (size_t i = 0; i < mp->nodes.size(); i++) {
        if (mp->nodes[i] >= n_nodes) { dbg_set_error("mapping node id out of range"); return DBGPHMM_ERR_INVALID; }
        freqs[mp->nodes[i]] += std::exp(mp->logp[i]);
    }
    return DBGPHMM_OK;
} ABI_CATCH

// Prob + Prob of the reference (prob.rs:181-197): the larger term first, then x + ln_1p(exp(y - x)), with its special cases
static double host_log_add(double a, double b) {
    const double x = a >= b ? a : b, y = a >= b ? b : a;
    if (y == -INFINITY) return x;
    if (x == y) return x + LN2;
    return x + std::log1p(std::exp(y - x));
}
extern "C" int dbgphmm_mappings_map_nodes(const dbgphmm_mappings* mp, uint32_t n_nodes_before, const uint64_t* map_off, const uint32_t* map_to,
                                          dbgphmm_mappings** out) try {
    if (!mp || !map_off || !out) { dbg_set_error("mappings_map_nodes: bad argument"); return DBGPHMM_ERR_INVALID; }
    for (uint32_t v = 0; v < n_nodes_before; v++)
        if (map_off[v + 1] < map_off[v]) { dbg_set_error("mappings_map_nodes: map_off is not non-decreasing"); return DBGPHMM_ERR_INVALID; }
    // hint.rs:66-88 — per base: every (node, p) spreads p / |node_map(node)| over its images, images that coincide are
    // added (Prob +), the row is re-sorted by probability, descending, and cut to MAX_ACTIVE_NODES.  The reference collects
    // the images in a HashMap, so the order among EQUAL probabilities is unspecified there; here: first appearance first.
    std::unique_ptr<dbgphmm_mappings> o(new dbgphmm_mappings());   // (owned until handed out: an exception below must not leak it)
    o->read_off = mp->read_off;
    o->row_off.assign(1, 0);
    std::vector<std::pair<uint32_t, double>> acc;   // (image, ln p) in order of first appearance
    std::unordered_map<uint32_t, uint32_t> where;
    std::vector<uint32_t> order;
    const size_t n_rows = mp->row_off.size() - 1;
    for (size_t r = 0; r < n_rows; r++) {
        acc.clear(); where.clear();
        for (uint64_t i = mp->row_off[r]; i < mp->row_off[r + 1]; i++) {
            const uint32_t v = mp->nodes[i];
            if (v >= n_nodes_before) { dbg_set_error("mappings_map_nodes: node id out of range of the node map"); return DBGPHMM_ERR_INVALID; }
            const uint64_t a = map_off[v], b = map_off[v + 1];
            if (a == b) continue;
            if (!map_to) { dbg_set_error("mappings_map_nodes: bad argument"); return DBGPHMM_ERR_INVALID; }
            const double share = mp->logp[i] - std::log((double)(b - a));   // Prob / usize (prob.rs:271-279)
            for (uint64_t j = a; j < b; j++) {
                auto it = where.find(map_to[j]);
                if (it == where.end()) { where.emplace(map_to[j], (uint32_t)acc.size()); acc.emplace_back(map_to[j], host_log_add(-INFINITY, share)); }
                else acc[it->second].second = host_log_add(acc[it->second].second, share);
            }
        }
        order.resize(acc.size());
        for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return acc[x].second > acc[y].second; });
        const size_t keep = std::min<size_t>(order.size(), DBGPHMM_MAX_ACTIVE_NODES);
        for (size_t i = 0; i < keep; i++) { o->nodes.push_back(acc[order[i]].first); o->logp.push_back(acc[order[i]].second); }
        o->row_off.push_back(o->nodes.size());
    }
    *out = o.release();
    return DBGPHMM_OK;
} ABI_CATCH

// ------------------------------------------------------------------ PHMMTables of one read
struct dbgphmm_tables {
    dbgphmm_model* m = nullptr;
    int device = -1;             // kept separately: the handle may be destroyed after its model
    int dir = 0, kind = 0;
    std::vector<uint8_t> bases;
    uint8_t* d_bases = nullptr;
    RowStore store;
    std::vector<RowDesc> desc;  // host copy
};

static int tables_finish(dbgphmm_tables* t) {
    t->desc.resize(t->store.n_desc);
    if (t->store.n_desc) CUDA_TRY(cudaMemcpy(t->desc.data(), t->store.d_desc, sizeof(RowDesc) * t->store.n_desc, cudaMemcpyDeviceToHost));
    return DBGPHMM_OK;
}
extern "C" void dbgphmm_tables_destroy(dbgphmm_tables* t) {
    if (!t) return;
    if (t->device >= 0) cudaSetDevice(t->device);
    cache_set_stream(nullptr);   // (the model the rows belong to may be gone: no stream to order the reuse behind)
    t->store.release();
    cudaFree(t->d_bases);
    delete t;
}
static int one_job(dbgphmm_model* m, const uint8_t* bases, uint64_t n, const dbgphmm_mappings* mp, uint64_t read_index, dbgphmm_tables* t,
                   std::vector<HJob>* jobs, DevMappings* dmap, bool need_map) {
    if (n == 0 || n > 0x7fffffffu) { dbg_set_error("read length must be in [1, 2^31)"); return DBGPHMM_ERR_INVALID; }
    for (uint64_t i = 0; i < n; i++)
        if (bases[i] != 'A' && bases[i] != 'C' && bases[i] != 'G' && bases[i] != 'T') { dbg_set_error("bases must be uppercase ACGT"); return DBGPHMM_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    t->m = m; t->device = m->device; t->bases.assign(bases, bases + n);
    CUDA_TRY(cudaMalloc((void**)&t->d_bases, n));
    CUDA_TRY(cudaMemcpy(t->d_bases, bases, n, cudaMemcpyHostToDevice));
    HJob j{}; j.read = 0; j.x = 0; j.base_off = 0; j.len = (uint32_t)n; j.map_row0 = 0;
    if (need_map) {
        if (!mp || read_index + 1 >= mp->read_off.size()) { dbg_set_error("mapping kind needs a mappings handle and a valid read index"); return DBGPHMM_ERR_INVALID; }
        if (mp->read_off[read_index + 1] - mp->read_off[read_index] != n) { dbg_set_error("mapping length differs from read length"); return DBGPHMM_ERR_INVALID; }
        ST_TRY(upload_mappings(m, mp, dmap));
        j.map_row0 = mp->read_off[read_index];
    }
    jobs->push_back(j);
    return DBGPHMM_OK;
}

extern "C" int dbgphmm_forward(dbgphmm_model* m, const uint8_t* bases, uint64_t n, int kind, const dbgphmm_mappings* mapping,
                               uint64_t read_index, dbgphmm_tables** out) try {
    if (!m || !bases || !out || kind < 0 || kind > 3) { dbg_set_error("forward: bad argument"); return DBGPHMM_ERR_INVALID; }
    dbgphmm_tables* t = new dbgphmm_tables();
    t->dir = 0; t->kind = kind;
    std::vector<HJob> jobs; DevMappings dmap;
    int st = one_job(m, bases, n, mapping, read_index, t, &jobs, &dmap, kind == DBGPHMM_FWD_MAPPING);
    if (st == DBGPHMM_OK) st = run_forward(m, jobs, t->d_bases, kind, PhaseOpts(), kind == DBGPHMM_FWD_MAPPING ? &dmap : nullptr, &t->store);
    if (st == DBGPHMM_OK) st = tables_finish(t);
    dmap.release();
    if (st != DBGPHMM_OK) { dbgphmm_tables_destroy(t); return st; }
    *out = t;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_backward(dbgphmm_model* m, const uint8_t* bases, uint64_t n, int kind, const dbgphmm_mappings* mapping,
                                uint64_t read_index, const dbgphmm_tables* fwd, dbgphmm_tables** out) try {
    if (!m || !bases || !out || kind < 0 || kind > 3) { dbg_set_error("backward: bad argument"); return DBGPHMM_ERR_INVALID; }
    if (kind == DBGPHMM_BWD_BY_FORWARD && (!fwd || fwd->dir != 0 || fwd->bases.size() != n)) { dbg_set_error("backward_by_forward needs the forward tables of the same read"); return DBGPHMM_ERR_INVALID; }
    dbgphmm_tables* t = new dbgphmm_tables();
    t->dir = 1; t->kind = kind;
    std::vector<HJob> jobs; DevMappings dmap;
    int st = one_job(m, bases, n, mapping, read_index, t, &jobs, &dmap, kind == DBGPHMM_BWD_MAPPING);
    if (st == DBGPHMM_OK) st = run_backward(m, jobs, t->d_bases, kind, PhaseOpts(), kind == DBGPHMM_BWD_MAPPING ? &dmap : nullptr, fwd ? &fwd->store : nullptr, &t->store);
    if (st == DBGPHMM_OK) st = tables_finish(t);
    dmap.release();
    if (st != DBGPHMM_OK) { dbgphmm_tables_destroy(t); return st; }
    *out = t;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" uint64_t dbgphmm_tables_len(const dbgphmm_tables* t) { return t ? t->desc.size() : 0; }
extern "C" int dbgphmm_tables_full_prob(const dbgphmm_tables* t, double* logp) try {
    if (!t || !logp || t->desc.empty()) { dbg_set_error("tables_full_prob: empty tables (table.rs:380-393 panics)"); return DBGPHMM_ERR_INVALID; }
    *logp = t->dir == 0 ? xlog(t->desc.back().e) : xlog(t->desc.front().mb);
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_tables_row_info(const dbgphmm_tables* t, int64_t row, uint64_t info[3], double sc[3]) try {
    if (!t || row < -1 || row >= (int64_t)t->desc.size()) { dbg_set_error("tables_row_info: row out of range"); return DBGPHMM_ERR_INVALID; }
    if (row < 0) {  // init_table: f_init (forward.rs:255-266) / b_init (backward.rs:197-211), both dense
        info[0] = 1; info[1] = t->m->N; info[2] = t->m->N;
        sc[0] = t->dir == 0 ? 0.0 : -INFINITY; sc[1] = -INFINITY; sc[2] = -INFINITY;
        return DBGPHMM_OK;
    }
    const RowDesc& r = t->desc[row];
    info[0] = r.kind == ROW_DENSE; info[1] = r.kind == ROW_DENSE ? t->m->N : r.n_mi; info[2] = r.kind == ROW_DENSE ? t->m->N : r.n_d;
    sc[0] = xlog(r.mb); sc[1] = xlog(r.ib); sc[2] = xlog(r.e);
    return DBGPHMM_OK;
} ABI_CATCH
static inline double lg(double v, int e) { return v == 0.0 ? -INFINITY : std::log(v) + (double)e * LN2; }

// raw copy of a stored row to the host
struct HostRow { std::vector<double> m, i, d; std::vector<int> ex; std::vector<uint32_t> id; std::vector<uint16_t> dlist; };
static int fetch_row(const dbgphmm_tables* t, const RowDesc& r, HostRow* h) {
    const dbgphmm_model* m = t->m;
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    if (r.kind == ROW_DENSE) {
        uint32_t N = m->N, Np = t->store.pool.Np;
        const char* sl = t->store.pool.base + r.off * t->store.pool.slab_bytes;
        h->m.resize(N); h->i.resize(N); h->d.resize(N); h->ex.resize(N);
        CUDA_TRY(cudaMemcpy(h->m.data(), sl, 8ull * N, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h->i.data(), sl + 8ull * Np, 8ull * N, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h->d.data(), sl + 16ull * Np, 8ull * N, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h->ex.data(), sl + 24ull * Np, 4ull * N, cudaMemcpyDeviceToHost));
    } else {
        uint32_t n = r.n_ent;
        const char* pay = t->store.arena.base + r.off;
        h->m.resize(n); h->i.resize(n); h->d.resize(n); h->ex.resize(n); h->id.resize(n); h->dlist.resize(r.n_d);
        if (n) {
            CUDA_TRY(cudaMemcpy(h->m.data(), pay, 8ull * n, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(h->i.data(), pay + 8ull * n, 8ull * n, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(h->d.data(), pay + 16ull * n, 8ull * n, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(h->id.data(), pay + 24ull * n, 4ull * n, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(h->ex.data(), pay + 28ull * n, 4ull * n, cudaMemcpyDeviceToHost));
        }
        if (r.n_d) CUDA_TRY(cudaMemcpy(h->dlist.data(), pay + 32ull * n, 2ull * r.n_d, cudaMemcpyDeviceToHost));
    }
    return DBGPHMM_OK;
}

extern "C" int dbgphmm_tables_row_export(const dbgphmm_tables* t, int64_t row, uint32_t* ids_mi, double* om, double* oi, uint32_t* ids_d, double* od) try {
    if (!t || row < -1 || row >= (int64_t)t->desc.size()) { dbg_set_error("tables_row_export: row out of range"); return DBGPHMM_ERR_INVALID; }
    const dbgphmm_model* m = t->m;
    if (row < 0) {
        double v = t->dir == 0 ? -INFINITY : m->params.p_end;
        for (uint32_t k = 0; k < m->N; k++) { om[k] = v; oi[k] = v; od[k] = v; }
        return DBGPHMM_OK;
    }
    const RowDesc& r = t->desc[row];
    HostRow h;
    ST_TRY(fetch_row(t, r, &h));
    if (r.kind == ROW_DENSE) {
        for (uint32_t p = 0; p < m->N; p++) {
            uint32_t o = m->orig_of[p];
            om[o] = lg(h.m[p], h.ex[p]); oi[o] = lg(h.i[p], h.ex[p]); od[o] = lg(h.d[p], h.ex[p]);
        }
    } else {
        for (uint32_t e = 0; e < r.n_mi; e++) { ids_mi[e] = m->orig_of[h.id[e]]; om[e] = lg(h.m[e], h.ex[e]); oi[e] = lg(h.i[e], h.ex[e]); }
        for (uint32_t a = 0; a < r.n_d; a++) { uint32_t e = h.dlist[a]; ids_d[a] = m->orig_of[h.id[e]]; od[a] = lg(h.d[e], h.ex[e]); }
    }
    return DBGPHMM_OK;
} ABI_CATCH

extern "C" int dbgphmm_tables_row_top_nodes(const dbgphmm_tables* t, int64_t row, int by_ratio, uint32_t k, double ratio, uint32_t* out, uint32_t* n_out) try {
    if (!t || !out || !n_out || row < 0 || row >= (int64_t)t->desc.size()) { dbg_set_error("tables_row_top_nodes: row out of range"); return DBGPHMM_ERR_INVALID; }
    dbgphmm_model* m = t->m;
    const RowDesc& r = t->desc[row];
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    if (r.kind == ROW_DENSE) {  // the device selection kernel used by the forward/backward drivers
        DevBuf b_req, b_ids, b_cnt;
        std::vector<SelectReq> rq(1);
        rq[0].slab = r.off; rq[0].k = by_ratio ? MAX_ACTIVE : std::min<uint32_t>(k, MAX_ACTIVE); rq[0].by_ratio = by_ratio; rq[0].ratio = ratio; rq[0].active_idx = -1; rq[0].out = 0;
        ST_TRY(dev_upload(b_req, rq, MSET(m).stream));
        ST_TRY(b_ids.alloc(sizeof(uint32_t) * MAX_ACTIVE)); ST_TRY(b_cnt.alloc(sizeof(uint32_t)));
        ST_TRY(dense_select(m, t->store.pool, b_req.as<SelectReq>(), 1, nullptr, b_ids.as<uint32_t>(), b_cnt.as<uint32_t>()));
        uint32_t cnt = 0; std::vector<uint32_t> ids(MAX_ACTIVE);
        CUDA_TRY(cudaMemcpyAsync(&cnt, b_cnt.p, 4, cudaMemcpyDeviceToHost, MSET(m).stream));
        CUDA_TRY(cudaMemcpyAsync(ids.data(), b_ids.p, 4 * MAX_ACTIVE, cudaMemcpyDeviceToHost, MSET(m).stream));
        CUDA_TRY(cudaStreamSynchronize(MSET(m).stream));
        for (uint32_t a = 0; a < cnt; a++) out[a] = m->orig_of[ids[a]];
        *n_out = cnt;
        return DBGPHMM_OK;
    }
    // sparse row: same ordering rule as the device kernels (value desc, then entry position), evaluated on the row's cells
    HostRow h;
    ST_TRY(fetch_row(t, r, &h));
    uint32_t n = r.n_ent;
    std::vector<uint32_t> ord(n);
    std::vector<XF> key(n);
    for (uint32_t e = 0; e < n; e++) { ord[e] = e; key[e] = xnorm(xf(h.m[e] + h.i[e] + h.d[e], h.ex[e])); }
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) {
        if (key[a].e != key[b].e) return key[a].e > key[b].e;
        return key[a].v > key[b].v;
    });
    uint32_t K = by_ratio ? MAX_ACTIVE : k, cnt = 0;
    double L0 = n ? xlog(key[ord[0]]) : -INFINITY;
    for (uint32_t a = 0; a < n && a < K; a++) {
        if (by_ratio && !(L0 - xlog(key[ord[a]]) < ratio)) continue;
        out[cnt++] = m->orig_of[h.id[ord[a]]];
    }
    *n_out = cnt;
    return DBGPHMM_OK;
} ABI_CATCH

// ------------------------------------------------------------------ PHMMOutput of one read
static int check_pair(const dbgphmm_model* m, const dbgphmm_tables* f, const dbgphmm_tables* b) {
    if (!m || !f || !b || f->dir != 0 || b->dir != 1 || f->m != m || b->m != m || f->desc.size() != b->desc.size()) {
        dbg_set_error("PHMMOutput::new: forward/backward tables do not match (table.rs:473-478 asserts)"); return DBGPHMM_ERR_INVALID;
    }
    return DBGPHMM_OK;
}
extern "C" int dbgphmm_output_node_freqs(dbgphmm_model* m, const dbgphmm_tables* fwd, const dbgphmm_tables* bwd, double* freqs) try {
    ST_TRY(check_pair(m, fwd, bwd));
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    DevBuf b_f;
    ST_TRY(b_f.alloc(sizeof(double) * m->N));
    CUDA_TRY(cudaMemsetAsync(b_f.p, 0, sizeof(double) * m->N, MSET(m).stream));
    std::vector<HJob> jobs(1);
    jobs[0] = HJob{0, 0, 0, (uint32_t)fwd->desc.size(), 0};
    ST_TRY(run_products_freqs(m, jobs, fwd->store, bwd->store, b_f.as<double>()));
    CUDA_TRY(cudaMemcpy(freqs, b_f.p, sizeof(double) * m->N, cudaMemcpyDeviceToHost));
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_output_edge_and_init_freqs(dbgphmm_model* m, const dbgphmm_tables* fwd, const dbgphmm_tables* bwd, double* edge_freqs,
                                                  double* init_freqs) try {
    ST_TRY(check_pair(m, fwd, bwd));
    if (!edge_freqs || !init_freqs) { dbg_set_error("to_edge_and_init_freqs: null output"); return DBGPHMM_ERR_INVALID; }
    if (fwd->bases != bwd->bases) { dbg_set_error("to_edge_and_init_freqs: the two tables come from different reads (freq.rs:281-282)"); return DBGPHMM_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    DevBuf b_e, b_i;
    ST_TRY(b_e.alloc(sizeof(double) * std::max<uint32_t>(m->E, 1))); ST_TRY(b_i.alloc(sizeof(double) * m->N));
    CUDA_TRY(cudaMemsetAsync(b_e.p, 0, sizeof(double) * std::max<uint32_t>(m->E, 1), MSET(m).stream));
    CUDA_TRY(cudaMemsetAsync(b_i.p, 0, sizeof(double) * m->N, MSET(m).stream));
    std::vector<HJob> jobs(1);
    jobs[0] = HJob{0, 0, 0, (uint32_t)fwd->desc.size(), 0};
    ST_TRY(run_products_edge_freqs(m, jobs, fwd->store, bwd->store, fwd->d_bases, b_e.as<double>(), b_i.as<double>()));
    if (m->E) CUDA_TRY(cudaMemcpy(edge_freqs, b_e.p, sizeof(double) * m->E, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(init_freqs, b_i.p, sizeof(double) * m->N, cudaMemcpyDeviceToHost));
    return DBGPHMM_OK;
} ABI_CATCH
// q_score_exact (q.rs:66-96): init = sum_v A(Begin, v) ln p_init(v), trans = sum_(v,w) A(v, w) ln p_trans(v, w) over emittable nodes
// (emission != 'n'); the prior term is always zero.  Host-side: two dot products over the model's parameters of candidate x.
extern "C" int dbgphmm_q_score_exact(const dbgphmm_model* m, uint32_t x, const double* edge_freqs, const double* init_freqs, double out[3]) try {
    if (!m || !edge_freqs || !init_freqs || !out) { dbg_set_error("q_score_exact: bad argument"); return DBGPHMM_ERR_INVALID; }
    std::vector<double> li(m->N), lt(std::max<uint32_t>(m->E, 1));
    ST_TRY(dbgphmm_model_get_probs(m, x, li.data(), lt.data()));
    double init = 0.0, trans = 0.0;
    for (uint32_t v = 0; v < m->N; v++) {   // original node ids
        if (m->emission[m->pos_of[v]] == 'n') continue;
        if (!std::isfinite(li[v])) { dbg_set_error("q_score_exact: init_prob is not finite (q.rs:79 asserts)"); return DBGPHMM_ERR_INVALID; }
        init += init_freqs[v] * li[v];
    }
    for (uint32_t e = 0; e < m->E; e++) {
        if (m->emission[m->pos_of[m->e_src[e]]] == 'n' || m->emission[m->pos_of[m->e_dst[e]]] == 'n') continue;
        if (!std::isfinite(lt[e])) { dbg_set_error("q_score_exact: trans_prob is not finite (q.rs:88 asserts)"); return DBGPHMM_ERR_INVALID; }
        trans += edge_freqs[e] * lt[e];
    }
    out[0] = init; out[1] = trans; out[2] = 0.0;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_output_mapping(dbgphmm_model* m, const dbgphmm_tables* fwd, const dbgphmm_tables* bwd, int by_ratio, uint32_t n_active,
                                      double ratio, dbgphmm_mappings** out) try {
    ST_TRY(check_pair(m, fwd, bwd));
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    std::vector<HJob> jobs(1);
    jobs[0] = HJob{0, 0, 0, (uint32_t)fwd->desc.size(), 0};
    dbgphmm_mappings* mp = new dbgphmm_mappings();
    int st = run_products_mapping(m, jobs, fwd->store, bwd->store, by_ratio, n_active, ratio, mp);
    if (st != DBGPHMM_OK) { delete mp; return st; }
    *out = mp;
    return DBGPHMM_OK;
} ABI_CATCH

// ------------------------------------------------------------------ bulk calls
extern "C" uint32_t dbgphmm_model_wave_reads(const dbgphmm_model* m) {
    if (!m) return 0;
    if (cudaSetDevice(m->device) != cudaSuccess) { cudaGetLastError(); return 0; }
    return sparse_wave_jobs(const_cast<dbgphmm_model*>(m), sparse_default_cap());
}
static void reset_times() { g_times = EngineTimes(); }
extern "C" int dbgphmm_last_timing(double ms[4], uint64_t* dense_cells) try {
    ms[0] = g_times.dense_ms; ms[1] = g_times.sparse_ms; ms[2] = g_times.product_ms; ms[3] = g_times.total_ms;
    if (dense_cells) *dense_cells = g_times.dense_cells;
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_last_dense_kernel(double* ms, uint64_t* launches, uint64_t* cells) try {
    if (ms) *ms = g_times.dense_kernel_ms;
    if (launches) *launches = g_times.dense_kernel_launches;
    if (cells) *cells = g_times.dense_kernel_cells;
    return DBGPHMM_OK;
} ABI_CATCH

static int check_reads_mappings(const dbgphmm_reads* reads, const dbgphmm_mappings* mp) {
    if (!mp) return DBGPHMM_OK;
    if (mp->read_off.size() != reads->n_reads + 1) { dbg_set_error("mappings / reads count mismatch"); return DBGPHMM_ERR_INVALID; }
    for (uint64_t r = 0; r < reads->n_reads; r++)
        if (mp->read_off[r + 1] - mp->read_off[r] != reads->off[r + 1] - reads->off[r]) { dbg_set_error("mapping length differs from read length"); return DBGPHMM_ERR_INVALID; }
    return DBGPHMM_OK;
}

// split job indices [0, n) into batches whose estimated device footprint fits the budget
// Batches of consecutive jobs within the memory budget.  quantum > 0: a batch that is not the last one is cut down to a
// multiple of `quantum` jobs (the sparse kernel runs one CTA per job and is latency-bound: a partial wave costs a full one).
static std::vector<std::pair<size_t, size_t>> plan_batches(const std::vector<uint64_t>& bytes, uint64_t budget, size_t max_jobs, size_t quantum = 0) {
    std::vector<std::pair<size_t, size_t>> out;
    size_t i = 0;
    while (i < bytes.size()) {
        uint64_t acc = 0; size_t j = i;
        while (j < bytes.size() && j - i < max_jobs && (j == i || acc + bytes[j] <= budget)) { acc += bytes[j]; j++; }
        if (quantum && j < bytes.size() && j - i > quantum) j = i + (j - i) / quantum * quantum;
        out.push_back({i, j});
        i = j;
    }
    return out;
}

extern "C" int dbgphmm_to_full_prob_reads(dbgphmm_model* m, const dbgphmm_reads* reads, const dbgphmm_mappings* mappings, int use_max_ratio,
                                          double* out_logp, double* out_per_read) try {
    if (!m || !reads || !out_logp) { dbg_set_error("to_full_prob_reads: bad argument"); return DBGPHMM_ERR_INVALID; }
    ST_TRY(check_reads_mappings(reads, mappings));
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    ST_TRY(dbgphmm_reads_to_device(m, const_cast<dbgphmm_reads*>(reads)));
    reset_times();
    EvTimer total(MSET(m).stream, &g_times.total_ms);
    const uint64_t R = reads->n_reads; const uint32_t X = m->n_batch;
    const int kind = mappings ? DBGPHMM_FWD_MAPPING : (use_max_ratio ? DBGPHMM_FWD_SPARSE_RATIO : DBGPHMM_FWD_SPARSE);
    DevMappings dmap;
    if (mappings) ST_TRY(upload_mappings(m, mappings, &dmap));
    std::vector<HJob> all; all.reserve(R * X);
    std::vector<uint64_t> bytes; bytes.reserve(R * X);
    const uint64_t slab = dense_slab_bytes(m->N);
    for (uint32_t x = 0; x < X; x++)
        for (uint64_t r = 0; r < R; r++) {
            uint64_t len = reads->off[r + 1] - reads->off[r];
            if (len == 0 || len > 0x7fffffffu) { dmap.release(); dbg_set_error("read length must be in [1, 2^31)"); return DBGPHMM_ERR_INVALID; }
            HJob j{(uint32_t)r, x, reads->off[r], (uint32_t)len, mappings ? mappings->read_off[r] : 0};
            all.push_back(j);
            bytes.push_back((mappings ? 0 : 2 * slab) + len * sizeof(RowDesc) + 65536);
        }
    std::vector<double> per(all.size());
    int st = DBGPHMM_OK;
    // Mapping-restricted scoring of several candidates: one CTA per (read, 8 candidates) shares the index work of every row
    // (mapx.cu); groups it cannot take (rows of more than 64 nodes, ...) go through the general path below.
    std::vector<uint8_t> done(all.size(), 0);
    if (mappings && X > 1 && !getenv("DBGPHMM_NO_MAPX")) {
        std::vector<MapxGroup> groups;
        for (uint64_t r = 0; r < R; r++)
            for (uint32_t x0 = 0; x0 < X; x0 += 8)
                groups.push_back(MapxGroup{reads->off[r], (uint32_t)(reads->off[r + 1] - reads->off[r]), mappings->read_off[r], x0, std::min<uint32_t>(8, X - x0),
                                           (uint32_t)((size_t)x0 * R + r)});
        std::vector<XF> fin(all.size(), xf_zero());
        std::vector<uint8_t> failed;
        st = run_mapx(m, groups, reads->d_bases, dmap, (uint32_t)R, fin.data(), failed, nullptr);
        if (st != DBGPHMM_OK) { dmap.release(); return st; }
        for (size_t g = 0; g < groups.size(); g++) {
            if (failed[g]) continue;
            for (uint32_t k = 0; k < groups[g].nx; k++) { const size_t idx = groups[g].out0 + (size_t)k * R; per[idx] = xlog(fin[idx]); done[idx] = 1; }
        }
        std::vector<HJob> rest; std::vector<uint64_t> rbytes; std::vector<size_t> ridx;
        for (size_t i = 0; i < all.size(); i++) if (!done[i]) { rest.push_back(all[i]); rbytes.push_back(bytes[i]); ridx.push_back(i); }
        for (auto& bt : plan_batches(rbytes, model_budget(m), 1u << 20, sparse_wave_jobs(m, sparse_default_cap()))) {
            std::vector<HJob> jobs(rest.begin() + bt.first, rest.begin() + bt.second);
            RowStore F;
            PhaseOpts so; so.keep_rows = false; so.store_sparse = false;
            st = run_forward(m, jobs, reads->d_bases, kind, so, &dmap, &F);
            if (st == DBGPHMM_OK) for (size_t i = 0; i < jobs.size(); i++) per[ridx[bt.first + i]] = xlog(F.h_final[i]);
            F.release();
            if (st != DBGPHMM_OK) break;
        }
        all.clear(); bytes.clear();
    }
    for (auto& bt : plan_batches(bytes, model_budget(m), mappings ? (1u << 20) : 65535, sparse_wave_jobs(m, sparse_default_cap()))) {
        std::vector<HJob> jobs(all.begin() + bt.first, all.begin() + bt.second);
        RowStore F;
        PhaseOpts so; so.keep_rows = false; so.store_sparse = false;
        st = run_forward(m, jobs, reads->d_bases, kind, so, mappings ? &dmap : nullptr, &F);
        if (st == DBGPHMM_OK) for (size_t i = 0; i < jobs.size(); i++) per[bt.first + i] = xlog(F.h_final[i]);
        F.release();
        if (st != DBGPHMM_OK) break;
    }
    dmap.release();
    if (st != DBGPHMM_OK) return st;
    for (uint32_t x = 0; x < X; x++) {
        double s = 0.0;  // Product over reads = sum of logs, fixed read order (prob.rs:245-249)
        for (uint64_t r = 0; r < R; r++) { s += per[(size_t)x * R + r]; if (out_per_read) out_per_read[(size_t)x * R + r] = per[(size_t)x * R + r]; }
        out_logp[x] = s;
    }
    return DBGPHMM_OK;
} ABI_CATCH

struct Gate {   // one-shot, idempotent
    std::mutex mu; std::condition_variable cv; bool is_open = false;
    void open() { { std::lock_guard<std::mutex> lk(mu); is_open = true; } cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return is_open; }); }
};
static void run_kinds(int mode, int use_max_ratio, int* fk, int* bk) {
    switch (mode) {
        case DBGPHMM_RUN_DENSE: *fk = DBGPHMM_FWD_DENSE; *bk = DBGPHMM_BWD_DENSE; break;
        case DBGPHMM_RUN_SPARSE: *fk = DBGPHMM_FWD_SPARSE; *bk = DBGPHMM_BWD_SPARSE; break;
        case DBGPHMM_RUN_SPARSE_ADAPTIVE: *fk = use_max_ratio ? DBGPHMM_FWD_SPARSE_RATIO : DBGPHMM_FWD_SPARSE; *bk = DBGPHMM_BWD_BY_FORWARD; break;
        default: *fk = DBGPHMM_FWD_MAPPING; *bk = DBGPHMM_BWD_MAPPING; break;
    }
}
// device bytes one read needs when both directions keep their rows
static uint64_t job_bytes_keep(const dbgphmm_model* m, int fk, int bk, uint64_t len) {
    const uint64_t slab = dense_slab_bytes(m->N), W = m->params.n_warmup;
    uint64_t fd = fk == DBGPHMM_FWD_DENSE ? len : (fk == DBGPHMM_FWD_MAPPING ? 0 : std::min(len, W));
    uint64_t bd = bk == DBGPHMM_BWD_DENSE ? len : (bk == DBGPHMM_BWD_SPARSE ? std::min(len, W) : (bk == DBGPHMM_BWD_BY_FORWARD ? std::min(len, W) + 2 : 0));
    const bool ratio = fk == DBGPHMM_FWD_SPARSE_RATIO;
    const uint64_t per_row = (ratio ? 2048 : (uint64_t)m->params.n_active_nodes * 48 + 256) + 2 * sizeof(RowDesc);   // (arena_estimate, engine.cu)
    return (fd + bd) * slab + 2 * len * per_row + ((uint64_t)1 << 20);
}

static int run_impl(dbgphmm_model* m, const dbgphmm_reads* reads, int mode, int use_max_ratio, const dbgphmm_mappings* mappings,
                    double* d_freqs, double* h_logp_fwd, double* h_logp_bwd, uint64_t cells[2], dbgphmm_mappings* map_out, int map_by_ratio) {
    ST_TRY(check_reads_mappings(reads, mappings));
    if (mode == DBGPHMM_RUN_WITH_MAPPING && !mappings) { dbg_set_error("run_with_mapping needs mappings"); return DBGPHMM_ERR_INVALID; }
    ST_TRY(dbgphmm_reads_to_device(m, const_cast<dbgphmm_reads*>(reads)));
    reset_times();
    EvTimer total(MSET(m).stream, &g_times.total_ms);
    int fk, bk;
    run_kinds(mode, use_max_ratio, &fk, &bk);
    const bool with_map = mode == DBGPHMM_RUN_WITH_MAPPING;
    DevMappings dmap;
    if (with_map) ST_TRY(upload_mappings(m, mappings, &dmap));
    const uint64_t R = reads->n_reads;
    std::vector<HJob> all; std::vector<uint64_t> bytes;
    for (uint64_t r = 0; r < R; r++) {
        uint64_t len = reads->off[r + 1] - reads->off[r];
        if (len == 0 || len > 0x7fffffffu) { dmap.release(); dbg_set_error("read length must be in [1, 2^31)"); return DBGPHMM_ERR_INVALID; }
        all.push_back(HJob{(uint32_t)r, 0, reads->off[r], (uint32_t)len, with_map ? mappings->read_off[r] : 0});
        bytes.push_back(job_bytes_keep(m, fk, bk, len));
    }
    if (cells) cells[0] = cells[1] = 0;
    int st = DBGPHMM_OK;
    // Strategy.  "store": both directions keep every row of a batch, products afterwards (any mode, any read length).
    // "stream" (run_sparse on graphs whose dense rows do not fit): dense rows live in two ping-pong slabs per read; the
    // pairs (dense row, sparse row of the other direction) are taken on the fly, which costs one recomputation of the
    // forward warm-up rows but decouples the batch size from the 28 B x N x 2W bytes a read's dense rows would need.
    const uint64_t W = m->params.n_warmup;
    bool can_stream = (mode == DBGPHMM_RUN_SPARSE) && !map_out && d_freqs;
    uint64_t store_total = 0;
    for (uint64_t r = 0; r < R; r++) { store_total += bytes[r]; if (all[r].len < 2 * W + 2) can_stream = false; }
    const uint64_t budget_now = model_budget(m);
    bool stream = can_stream && store_total > budget_now;
    if (const char* e = getenv("DBGPHMM_STRATEGY")) { if (!strcmp(e, "stream")) stream = can_stream; else if (!strcmp(e, "store")) stream = false; }
    // Stream strategy: the dense warm-up and the recompute passes can run in GROUPS of G jobs that share one pool of 2 G slabs
    // (engine.cu), so that a batch is sized by its sparse rows and fills the sparse kernel's resident wave even when two slabs per read
    // would not leave room for that many reads (C5: 187 MB slabs -> 296 reads = 2 jobs per SM without groups).  G is chosen here: no
    // groups when a full wave fits with two slabs per read (C3), else the largest G the budget leaves beside a wave's sparse rows.
    // DBGPHMM_DENSE_GROUP=G forces a group size (0: never group).
    uint32_t group = 0;
    uint64_t plan_budget = budget_now;
    if (stream) {
        const uint64_t slab = dense_slab_bytes(m->N);
        const uint64_t per_row = (uint64_t)m->params.n_active_nodes * 48 + 256 + 3 * sizeof(RowDesc);   // (arena_estimate, engine.cu)
        const uint64_t gather_part = (uint64_t)32 * sparse_gather_cap(m, m->params.n_active_nodes) + 64 * (uint64_t)m->fwd.n_chunks;
        auto rows_bytes = [&](uint64_t r) { return 2 * (uint64_t)all[r].len * per_row + ((uint64_t)1 << 20); };
        const uint64_t wave = sparse_wave_jobs(m, sparse_default_cap());
        bool forced = false;
        if (const char* e = getenv("DBGPHMM_DENSE_GROUP")) { const int g = atoi(e); forced = true; if (g > 0) group = (uint32_t)g; }
        if (!forced) {
            const uint64_t target = std::min<uint64_t>(R, wave);
            uint64_t acc = 0, fit = 0;
            while (fit < R && acc + 2 * slab + rows_bytes(fit) <= budget_now) { acc += 2 * slab + rows_bytes(fit); fit++; }
            if (fit < target) {   // two slabs per read do not leave room for a full wave
                // the largest batch (a wave, half a wave, ...) beyond what fits ungrouped whose sparse rows leave room for a pool of
                // at least 16 reads' slabs (the dense kernels want a few hundred thousand tiles per launch)
                for (uint64_t jt = target; jt > fit && !group; jt = jt / 2) {
                    uint64_t rows = 0;
                    for (uint64_t r = 0; r < jt; r++) rows += gather_part + rows_bytes(r);
                    const uint64_t reserve = budget_now >> 4;
                    if (rows + reserve < budget_now) {
                        const uint64_t g = (budget_now - rows - reserve) / (2 * slab);
                        if (g >= 16 || g >= jt) group = (uint32_t)std::min<uint64_t>(g, jt);
                    }
                }
            }
            if (group && getenv("DBGPHMM_TRACE")) fprintf(stderr, "[dbgphmm] dense warm-up in groups of %u reads (two slabs per read would fit %llu of %llu reads)\n", group, (unsigned long long)fit, (unsigned long long)R);
        }
        const uint64_t dense_part = group ? gather_part : 2 * slab;
        for (uint64_t r = 0; r < R; r++) bytes[r] = dense_part + rows_bytes(r);
        if (group) {
            const uint64_t pool = 2 * (uint64_t)group * slab;
            if (pool + (plan_budget >> 3) > plan_budget) { dbg_set_error("DBGPHMM_DENSE_GROUP: the group's slabs do not fit the memory budget"); return DBGPHMM_ERR_OOM; }
            plan_budget -= pool;
        }
    }
    // Batches in read order.  A batch that runs out of device memory before anything of it has been accumulated (the estimates
    // above are for typical rows) is split in two and tried again.
    std::vector<std::pair<size_t, size_t>> work;
    {
        auto planned = plan_batches(bytes, plan_budget, 65535, mode == DBGPHMM_RUN_DENSE ? 0 : sparse_wave_jobs(m, sparse_default_cap()));
        work.assign(planned.rbegin(), planned.rend());   // (a stack: the first batch on top)
    }
    while (!work.empty()) {
        const std::pair<size_t, size_t> bt = work.back();
        work.pop_back();
        std::vector<HJob> jobs(all.begin() + bt.first, all.begin() + bt.second);
        HostTrace tr_b("batch");
        RowStore F, B;
        bool accumulated = false;   // something of this batch has reached d_freqs / map_out
        if (!stream) {
            PhaseOpts po;
            st = run_forward(m, jobs, reads->d_bases, fk, po, with_map ? &dmap : nullptr, &F);
            if (st == DBGPHMM_OK) st = run_backward(m, jobs, reads->d_bases, bk, po, with_map ? &dmap : nullptr, &F, &B);
            if (st == DBGPHMM_OK) accumulated = true;
            if (st == DBGPHMM_OK && d_freqs) st = run_products_freqs(m, jobs, F, B, d_freqs);
            if (st == DBGPHMM_OK && map_out) st = run_products_mapping(m, jobs, F, B, map_by_ratio, m->params.n_active_nodes, m->params.active_node_max_ratio, map_out);
        } else {
            DevBuf b_err;
            st = b_err.alloc(sizeof(int));
            if (st == DBGPHMM_OK && cudaMemsetAsync(b_err.p, 0, sizeof(int), MSET(m).stream) != cudaSuccess) st = DBGPHMM_ERR_CUDA;
            PhaseOpts pf; pf.keep_rows = false; pf.store_sparse = true; pf.group = group;
            // B dense x F sparse: on the fly after every dense backward step, or (when the two-rows-per-launch kernels are available,
            // which never write the intermediate rows) by a recompute pass inside the dependency cone like the forward one below
            const bool b_recompute = dense_can_pair(m);
            StepProducts spb; spb.other = &F; spb.P = F.d_final; spb.d_freqs = d_freqs; spb.d_err = b_err.as<int>();
            PhaseOpts pb; pb.keep_rows = false; pb.store_sparse = true; pb.step = b_recompute ? nullptr : &spb; pb.group = group;
            // The two directions do not depend on each other (with the recompute passes), and their sparse phases are latency-bound per
            // read: a second host thread drives the backward direction on the handle's second stream set.  Order: forward dense rows ->
            // (first-row inputs gathered, slabs released) -> backward dense rows -> BOTH sparse phases side by side: with fewer reads
            // than ~0.4 of a resident wave the CTAs of the four launches (two primary, two rescue) are resident at once and the phases
            // share their latency (measured per pass: 250 reads 699 -> 597 ms, 500 reads 1041 -> 965 ms).  Larger batches keep one phase
            // at a time (measured: no gain from 666 reads on, and a rescue launch that finds no shared memory starts its jobs late).
            // DBGPHMM_OVERLAP=0 never overlaps, =1 always does.
            bool overlap = b_recompute && fk == DBGPHMM_FWD_SPARSE && bk == DBGPHMM_BWD_SPARSE && sparse_pair_fits(m, sparse_default_cap(), (uint32_t)jobs.size());
            if (const char* e = getenv("DBGPHMM_OVERLAP")) overlap = b_recompute && fk == DBGPHMM_FWD_SPARSE && bk == DBGPHMM_BWD_SPARSE && e[0] != '0';
            if (st == DBGPHMM_OK && overlap) {
                Gate f_dense_done, b_dense_done;
                pf.force_gather = true;
                pf.after_dense = [&] { f_dense_done.open(); b_dense_done.wait(); };
                pb.before_dense = [&] { f_dense_done.wait(); };
                pb.after_dense = [&] { b_dense_done.open(); };
                int st_b = DBGPHMM_OK; std::string err_b; EngineTimes times_b;
                std::thread tb([&] {
                    tl_stream_set = 1;
                    try {
                        if (cudaSetDevice(m->device) != cudaSuccess) { dbg_set_error("cudaSetDevice failed in the backward thread"); st_b = DBGPHMM_ERR_CUDA; }
                        else {
                            cache_set_stream(MSET(m).stream);
                            g_times = EngineTimes();
                            st_b = run_backward(m, jobs, reads->d_bases, bk, pb, nullptr, nullptr, &B);
                        }
                    } catch (const std::bad_alloc&) { dbg_set_error("out of host memory"); st_b = DBGPHMM_ERR_OOM; }
                    catch (const std::exception& e) { dbg_set_error(std::string("C++ exception: ") + e.what()); st_b = DBGPHMM_ERR_INVALID; }
                    if (st_b != DBGPHMM_OK) err_b = dbgphmm_last_error();
                    b_dense_done.open();          // (whatever happened: nobody waits forever)
                    times_b = g_times;
                    launch_timer_release();
                    cache_set_stream(nullptr);
                });
                try { st = run_forward(m, jobs, reads->d_bases, fk, pf, nullptr, &F); }
                catch (...) { f_dense_done.open(); tb.join(); throw; }
                f_dense_done.open();
                tb.join();
                g_times.dense_ms += times_b.dense_ms; g_times.sparse_ms += times_b.sparse_ms; g_times.dense_cells += times_b.dense_cells;
                g_times.dense_kernel_ms += times_b.dense_kernel_ms; g_times.dense_kernel_launches += times_b.dense_kernel_launches;
                g_times.dense_kernel_cells += times_b.dense_kernel_cells;
                if (st == DBGPHMM_OK && st_b != DBGPHMM_OK) { st = st_b; dbg_set_error(err_b); }
                spb.P = F.d_final;
            } else {
                if (st == DBGPHMM_OK) st = run_forward(m, jobs, reads->d_bases, fk, pf, nullptr, &F);                    // F: sparse rows stored
                spb.P = F.d_final;    // (allocated by run_forward)
                if (st == DBGPHMM_OK) st = run_backward(m, jobs, reads->d_bases, bk, pb, nullptr, &F, &B);
            }
            if (st == DBGPHMM_OK || !b_recompute) accumulated = true;
            if (st == DBGPHMM_OK && b_recompute) st = run_backward_recompute(m, jobs, reads->d_bases, F, B, spb, group);
            if (st == DBGPHMM_OK) st = run_products_freqs(m, jobs, F, B, d_freqs);                                    // sparse x sparse, F sparse x b_init
            StepProducts spf; spf.other = &B; spf.P = F.d_final; spf.d_freqs = d_freqs; spf.d_err = b_err.as<int>();
            if (st == DBGPHMM_OK) st = run_forward_recompute(m, jobs, reads->d_bases, F, B, spf, group);                    // F dense (cone only) x B sparse
            int err = 0;
            if (st == DBGPHMM_OK && cudaMemcpy(&err, b_err.p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) st = DBGPHMM_ERR_CUDA;
            if (st == DBGPHMM_OK && err) { dbg_set_error("P(read) == 0: emit probabilities are NaN in the reference (table.rs:500-505)"); st = DBGPHMM_ERR_ZERO_PROB; }
        }
        if (st == DBGPHMM_OK) {
            for (size_t i = 0; i < jobs.size(); i++) {
                if (h_logp_fwd) h_logp_fwd[bt.first + i] = xlog(F.h_final[i]);
                if (h_logp_bwd) h_logp_bwd[bt.first + i] = xlog(B.h_final[i]);
            }
            if (cells) { cells[0] += F.cells; cells[1] += B.cells; }
        }
        { HostTrace tr_r("release"); F.release(); B.release(); }
        if (st == DBGPHMM_ERR_OOM && !accumulated && jobs.size() > 1) {
            if (getenv("DBGPHMM_TRACE")) fprintf(stderr, "[dbgphmm] batch of %zu reads ran out of device memory: splitting it\n", jobs.size());
            CUDA_TRY(cudaStreamSynchronize(MSET(m).stream));
            cache_trim();
            const size_t mid = bt.first + jobs.size() / 2;
            work.push_back({mid, bt.second});
            work.push_back({bt.first, mid});
            st = DBGPHMM_OK;
            continue;
        }
        if (st != DBGPHMM_OK) break;
    }
    dmap.release();
    return st;
}

extern "C" int dbgphmm_run_node_freqs_dev(dbgphmm_model* m, const dbgphmm_reads* reads, int mode, int use_max_ratio,
                                          const dbgphmm_mappings* mappings, double* node_freqs_dev, double* logp_fwd_dev, double* logp_bwd_dev,
                                          uint64_t cells[2]) try {
    if (!m || !reads || mode < 0 || mode > 3) { dbg_set_error("run_node_freqs: bad argument"); return DBGPHMM_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    std::vector<double> lf(reads->n_reads), lb(reads->n_reads);
    ST_TRY(run_impl(m, reads, mode, use_max_ratio, mappings, node_freqs_dev, lf.data(), lb.data(), cells, nullptr, 0));
    if (logp_fwd_dev && !lf.empty()) CUDA_TRY(cudaMemcpy(logp_fwd_dev, lf.data(), 8 * lf.size(), cudaMemcpyHostToDevice));
    if (logp_bwd_dev && !lb.empty()) CUDA_TRY(cudaMemcpy(logp_bwd_dev, lb.data(), 8 * lb.size(), cudaMemcpyHostToDevice));
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_run_node_freqs(dbgphmm_model* m, const dbgphmm_reads* reads, int mode, int use_max_ratio, const dbgphmm_mappings* mappings,
                                      double* node_freqs, double* logp_fwd, double* logp_bwd, uint64_t cells[2]) try {
    if (!m || !reads || mode < 0 || mode > 3) { dbg_set_error("run_node_freqs: bad argument"); return DBGPHMM_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    DevBuf b_f;
    if (node_freqs) { ST_TRY(b_f.alloc(sizeof(double) * m->N)); CUDA_TRY(cudaMemsetAsync(b_f.p, 0, sizeof(double) * m->N, MSET(m).stream)); }
    ST_TRY(run_impl(m, reads, mode, use_max_ratio, mappings, node_freqs ? b_f.as<double>() : nullptr, logp_fwd, logp_bwd, cells, nullptr, 0));
    if (node_freqs) CUDA_TRY(cudaMemcpy(node_freqs, b_f.p, sizeof(double) * m->N, cudaMemcpyDeviceToHost));
    return DBGPHMM_OK;
} ABI_CATCH
extern "C" int dbgphmm_generate_mappings(dbgphmm_model* m, const dbgphmm_reads* reads, const dbgphmm_mappings* mappings, int use_max_ratio,
                                         dbgphmm_mappings** out) try {
    if (!m || !reads || !out) { dbg_set_error("generate_mappings: bad argument"); return DBGPHMM_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(m->device));
    cache_set_stream(MSET(m).stream);
    dbgphmm_mappings* mp = new dbgphmm_mappings();
    mp->read_off.push_back(0); mp->row_off.push_back(0);
    // hint.rs:205-217: run_with_mapping if hints exist, else run_sparse_adaptive(use_max_ratio)
    int st = run_impl(m, reads, mappings ? DBGPHMM_RUN_WITH_MAPPING : DBGPHMM_RUN_SPARSE_ADAPTIVE, use_max_ratio, mappings, nullptr, nullptr, nullptr,
                      nullptr, mp, use_max_ratio ? 1 : 0);
    if (st != DBGPHMM_OK) { delete mp; return st; }
    *out = mp;
    return DBGPHMM_OK;
} ABI_CATCH
