// engine.cu — drives the dense and sparse phases of forward / backward for a batch of (read, X) jobs.
//
// forward  drivers: forward.rs:24-45 (dense), :51-75 (mapping), :93-154 (sparse top-n / max-ratio)
// backward drivers: backward.rs:24-53 (dense), :59-93 (mapping), :101-142 (by forward), :146-185 (sparse)
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include "engine.h"

// Host-side state shared by every handle of the process: the instrumentation is per host thread, the block cache is guarded by a
// mutex, so different handles may be driven from different threads (one thread per handle at a time, see dbgphmm_b200.h).
thread_local EngineTimes g_times;

static thread_local std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_ev_free, g_ev_busy;
static thread_local std::vector<uint64_t> g_ev_cells;
void launch_timer_begin(cudaStream_t st) {
    std::pair<cudaEvent_t, cudaEvent_t> p;
    if (!g_ev_free.empty()) { p = g_ev_free.back(); g_ev_free.pop_back(); }
    else { cudaEventCreate(&p.first); cudaEventCreate(&p.second); }
    cudaEventRecord(p.first, st);
    g_ev_busy.push_back(p);
}
void launch_timer_end(cudaStream_t st, uint64_t cells) {
    cudaEventRecord(g_ev_busy.back().second, st);
    g_ev_cells.push_back(cells);
}
void launch_timer_flush() {
    for (size_t i = 0; i < g_ev_busy.size(); i++) {
        cudaEventSynchronize(g_ev_busy[i].second);
        float ms = 0;
        cudaEventElapsedTime(&ms, g_ev_busy[i].first, g_ev_busy[i].second);
        g_times.dense_kernel_ms += ms; g_times.dense_kernel_launches++; g_times.dense_kernel_cells += g_ev_cells[i];
        g_ev_free.push_back(g_ev_busy[i]);
    }
    g_ev_busy.clear(); g_ev_cells.clear();
}
void launch_timer_release() {
    launch_timer_flush();
    for (auto& p : g_ev_free) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    g_ev_free.clear();
}

// Device-memory block cache shared by the handles of the process, stream-ordered: a block is returned while the work that used it
// may still be in flight on the stream of the code that frees it (cache_set_stream).  Taking it again on the SAME stream needs
// nothing (stream order); another stream first waits for an event recorded on the freeing stream at the time of the free, so no
// device-wide synchronisation is needed when two host threads drive two stream sets of one handle side by side.  Code that frees
// without a current stream (none of the library's own) falls back to a device synchronisation at the next foreign reuse.
struct CacheBlock { void* p; size_t bytes; int device; bool used; cudaStream_t owner; cudaEvent_t ev; bool ev_valid; };
static std::vector<CacheBlock> g_cache;
static std::recursive_mutex g_cache_mu;
static thread_local cudaStream_t tl_cache_stream = nullptr;
void cache_set_stream(cudaStream_t st) { tl_cache_stream = st; }
void* cache_alloc(size_t bytes) {
    std::lock_guard<std::recursive_mutex> lk(g_cache_mu);
    if (bytes < 256) bytes = 256;
    if (bytes < (1u << 20)) { size_t c = 256; while (c < bytes) c <<= 1; bytes = c; }  // small blocks in power-of-two classes
    int dev = 0; cudaGetDevice(&dev);
    const cudaStream_t me = tl_cache_stream;
    // best fit within 1.5 x the request, this stream's blocks first ; then, for large requests, up to 4 x: consecutive batches of
    // different shapes (1332 reads, then the 668 left over) otherwise find nothing that "fits", allocate their own set of blocks, run
    // out of memory, trim the cache and allocate again -- 400-600 ms of cudaMalloc / cudaFree per change of shape (measured)
    for (int pass = 0; pass < 3; pass++) {
        const bool any_owner = pass >= 1;
        if (pass == 2 && bytes < ((size_t)64 << 20)) break;
        const size_t limit = pass == 2 ? 4 * bytes : bytes + bytes / 2 + (1 << 20);
        int best = -1;
        for (size_t i = 0; i < g_cache.size(); i++) {
            const CacheBlock& b = g_cache[i];
            if (!b.used && b.device == dev && (any_owner || (me && b.owner == me)) && b.bytes >= bytes && b.bytes <= limit)
                if (best < 0 || b.bytes < g_cache[best].bytes) best = (int)i;
        }
        if (best >= 0) {
            CacheBlock& b = g_cache[best];
            if (!me || b.owner != me) {
                if (b.ev_valid && me) cudaStreamWaitEvent(me, b.ev, 0);
                else if (b.ev_valid) cudaEventSynchronize(b.ev);
                else cudaDeviceSynchronize();
            }
            b.owner = me; b.used = true;
            return b.p;
        }
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        cache_trim();
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    g_cache.push_back(CacheBlock{p, bytes, dev, true, me, nullptr, false});
    return p;
}
void cache_free(void* p) {
    if (!p) return;
    std::lock_guard<std::recursive_mutex> lk(g_cache_mu);
    for (auto& b : g_cache) if (b.p == p) {
        b.used = false; b.owner = tl_cache_stream; b.ev_valid = false;
        if (tl_cache_stream) {
            if (!b.ev && cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); b.ev = nullptr; }
            if (b.ev && cudaEventRecord(b.ev, tl_cache_stream) == cudaSuccess) b.ev_valid = true; else cudaGetLastError();
        }
        return;
    }
    cudaFree(p);
}
uint64_t cache_unused_bytes() {
    std::lock_guard<std::recursive_mutex> lk(g_cache_mu);
    int dev = 0; cudaGetDevice(&dev);
    uint64_t n = 0;
    for (auto& b : g_cache) if (!b.used && b.device == dev) n += b.bytes;
    return n;
}
// Device memory a bulk call may plan with.  A fixed budget given at model creation is taken as it is; otherwise 88 % of what is free
// NOW plus what this process holds in unused cache blocks (they are reused or trimmed): other allocations of the process made after
// the model was created -- torch tensors, NCCL buffers, another handle's rows -- are then accounted for.
uint64_t model_budget(const dbgphmm_model* m) {
    if (m->mem_budget_fixed) return m->mem_budget;
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return m->mem_budget; }
    return (uint64_t)(((double)fr + (double)cache_unused_bytes()) * 0.88);
}
void cache_trim() {
    std::lock_guard<std::recursive_mutex> lk(g_cache_mu);
    int dev = 0; cudaGetDevice(&dev);
    std::vector<CacheBlock> keep;
    for (auto& b : g_cache) {
        if (!b.used && b.device == dev) { cudaFree(b.p); if (b.ev) cudaEventDestroy(b.ev); }   // (cudaFree waits for the device)
        else keep.push_back(b);
    }
    g_cache.swap(keep);
}

void RowStore::release() {
    cache_free(d_desc); d_desc = nullptr;
    cache_free(d_desc0); d_desc0 = nullptr;
    cache_free(pool.base); pool = DensePool();
    cache_free(arena.base); cache_free(arena.cursor); arena = SparseArena();
    cache_free(d_final); d_final = nullptr;
}
void DevMappings::release() { if (owned) { cudaFree(row_off); cudaFree(nodes); } row_off = nullptr; nodes = nullptr; }

int upload_mappings(dbgphmm_model* m, const dbgphmm_mappings* mp, DevMappings* out) {
    if (!(mp->d_nodes && mp->d_model_serial == m->serial)) {
        mappings_release_device(mp);
        std::vector<uint32_t> nodes(mp->nodes.size());
        for (size_t i = 0; i < nodes.size(); i++) {
            if (mp->nodes[i] >= m->N) { dbg_set_error("mapping node id out of range"); return DBGPHMM_ERR_INVALID; }
            nodes[i] = m->pos_of[mp->nodes[i]];
        }
        uint64_t* d_off = nullptr; uint32_t* d_nd = nullptr;
        CUDA_TRY(cudaMalloc((void**)&d_off, sizeof(uint64_t) * std::max<size_t>(mp->row_off.size(), 1)));
        if (cudaMalloc((void**)&d_nd, sizeof(uint32_t) * std::max<size_t>(nodes.size(), 1)) != cudaSuccess) {
            cudaGetLastError(); cudaFree(d_off); dbg_set_error("out of device memory for the mappings"); return DBGPHMM_ERR_OOM;
        }
        mp->d_row_off = d_off; mp->d_nodes = d_nd; mp->d_model_serial = m->serial; mp->d_device = m->device;
        CUDA_TRY(cudaMemcpy(d_off, mp->row_off.data(), sizeof(uint64_t) * mp->row_off.size(), cudaMemcpyHostToDevice));
        if (!nodes.empty()) CUDA_TRY(cudaMemcpy(d_nd, nodes.data(), sizeof(uint32_t) * nodes.size(), cudaMemcpyHostToDevice));
    }
    out->row_off = mp->d_row_off; out->nodes = mp->d_nodes; out->owned = false;
    out->read_off = mp->read_off;
    return DBGPHMM_OK;
}
void mappings_release_device(const dbgphmm_mappings* mp) {
    if (!mp->d_nodes && !mp->d_row_off) return;
    int cur = 0; cudaGetDevice(&cur);
    if (mp->d_device >= 0) cudaSetDevice(mp->d_device);
    cudaFree(mp->d_row_off); cudaFree(mp->d_nodes);
    mp->d_row_off = nullptr; mp->d_nodes = nullptr; mp->d_model_serial = 0; mp->d_device = -1;
    cudaSetDevice(cur);
}

// ------------------------------------------------------------------ small control kernels
// forward_sparse(use_max_ratio = true): decide after dense row s whether row s+1 stays dense (forward.rs:119-133)
__global__ void k_ratio_decide(uint32_t n_jobs, uint32_t s, uint32_t W, uint32_t thr, const uint32_t* __restrict__ len,
                               const uint32_t* __restrict__ top_cnt, int* __restrict__ active, uint32_t* __restrict__ nd) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs || !active[j]) return;
    if (s + 1 >= len[j]) { active[j] = 0; nd[j] = len[j]; return; }
    bool use_dense = (s + 1 < W) && (top_cnt[j] > thr);
    if (!use_dense) { active[j] = 0; nd[j] = s + 1; }
}
__global__ void k_final_from_desc(uint32_t n_jobs, const RowDesc* __restrict__ desc, const uint64_t* __restrict__ desc0,
                                  const uint32_t* __restrict__ len, const uint8_t* __restrict__ take, int dir, XF* __restrict__ out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs || !take[j]) return;
    out[j] = dir == 0 ? desc[desc0[j] + len[j] - 1].e : desc[desc0[j]].mb;
}
// scatter a stored sparse row into a zeroed dense slab (backward_by_forward: dense step after sparse rows)
__global__ void k_scatter_row(uint32_t n, const uint64_t* __restrict__ desc_idx, const uint64_t* __restrict__ slab, const RowDesc* __restrict__ desc,
                              const char* __restrict__ arena, char* __restrict__ pool, uint64_t slab_bytes, uint32_t Np, uint32_t N) {
    uint32_t r = blockIdx.y;
    if (r >= n) return;
    char* sl = pool + slab[r] * slab_bytes;
    double* gm = (double*)sl; double* gi = gm + Np; double* gd = gi + Np; int* ge = (int*)(gd + Np);
    const RowDesc d = desc[desc_idx[r]];
    // phase split by blockIdx.x parity is not possible without a grid sync: zero here, scatter in a second launch
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < N; g += gridDim.x * blockDim.x) { gm[g] = 0.0; gi[g] = 0.0; gd[g] = 0.0; ge[g] = 0; }
    (void)d; (void)arena;
}
__global__ void k_scatter_row2(uint32_t n, const uint64_t* __restrict__ desc_idx, const uint64_t* __restrict__ slab, const RowDesc* __restrict__ desc,
                               const char* __restrict__ arena, char* __restrict__ pool, uint64_t slab_bytes, uint32_t Np) {
    uint32_t r = blockIdx.x;
    if (r >= n) return;
    char* sl = pool + slab[r] * slab_bytes;
    double* gm = (double*)sl; double* gi = gm + Np; double* gd = gi + Np; int* ge = (int*)(gd + Np);
    const RowDesc d = desc[desc_idx[r]];
    const char* pay = arena + d.off;
    const double* fm = (const double*)pay; const double* fi = fm + d.n_ent; const double* fd = fi + d.n_ent;
    const uint32_t* fid = (const uint32_t*)(fd + d.n_ent); const int* fex = (const int*)(fid + d.n_ent);
    for (uint32_t e = threadIdx.x; e < d.n_ent; e += blockDim.x) { uint32_t g = fid[e]; gm[g] = fm[e]; gi[g] = fi[e]; gd[g] = fd[e]; ge[g] = fex[e]; }
}

static int alloc_pool(DensePool& pool, uint32_t N, uint64_t n_slabs) {
    pool.Np = (N + 1) & ~1u;
    pool.slab_bytes = dense_slab_bytes(N);
    pool.n_slabs = n_slabs;
    if (n_slabs == 0) { pool.base = nullptr; return DBGPHMM_OK; }
    pool.base = (char*)cache_alloc(pool.slab_bytes * n_slabs);
    if (!pool.base) {
        dbg_set_error("out of device memory for dense rows"); return DBGPHMM_ERR_OOM;
    }
    return DBGPHMM_OK;
}
static int alloc_arena(SparseArena& a, uint64_t bytes, cudaStream_t st) {
    if (const char* e = getenv("DBGPHMM_ARENA_MAX_BYTES")) { const uint64_t cap = strtoull(e, nullptr, 10); if (cap >= 256 && bytes > cap) bytes = cap; }   // tests: force the batch split
    a.bytes = bytes;
    a.base = (char*)cache_alloc(bytes ? bytes : 256);
    if (!a.base) { dbg_set_error("out of device memory for sparse rows"); return DBGPHMM_ERR_OOM; }
    a.cursor = (unsigned long long*)cache_alloc(sizeof(unsigned long long));
    if (!a.cursor) { dbg_set_error("out of device memory"); return DBGPHMM_ERR_OOM; }
    CUDA_TRY(cudaMemsetAsync(a.cursor, 0, sizeof(unsigned long long), st));
    return DBGPHMM_OK;
}

// ------------------------------------------------------------------ DBGPHMM_VERIFY=1: structural check of every stored row
// (debug aid: descriptors are poisoned at allocation, so a row no kernel wrote is reported with its job and row)
static bool verify_enabled() { const char* e = getenv("DBGPHMM_VERIFY"); return e && e[0] == '1'; }
struct VerifyRec { uint32_t job, row, code; int kind; uint32_t n_ent, n_mi, n_d, bad_id; unsigned long long off; };
__global__ void k_verify_rows(uint32_t n_jobs, const RowDesc* __restrict__ desc, const uint64_t* __restrict__ desc0, const uint32_t* __restrict__ len,
                              const uint32_t* __restrict__ row_lo, const uint32_t* __restrict__ row_hi, const char* __restrict__ arena, uint64_t arena_bytes,
                              uint64_t n_slabs, uint32_t N, uint32_t* __restrict__ n_bad, VerifyRec* __restrict__ recs, uint32_t max_recs) {
    const uint32_t j = blockIdx.y;
    const uint32_t lo = row_lo[j], hi = row_hi[j] < len[j] ? row_hi[j] : len[j];
    for (uint32_t r = lo + blockIdx.x; r < hi; r += gridDim.x) {
        const RowDesc d = desc[desc0[j] + r];
        uint32_t code = 0, bad_id = 0;
        if (d.kind == ROW_DENSE) { if (d.off >= n_slabs && n_slabs) code = 2; }
        else if (d.kind == ROW_SPARSE) {
            if (d.n_ent > 832 || d.n_mi > d.n_ent || d.n_d > d.n_ent) code = 3;
            else if (d.off + sparse_row_bytes(d.n_ent, d.n_d) > arena_bytes || (d.off & 7)) code = 4;
            else {
                const uint32_t* id = (const uint32_t*)(arena + d.off + 24ull * d.n_ent);
                for (uint32_t e = threadIdx.x; e < d.n_ent; e += blockDim.x) if (id[e] >= N) { code = 5; bad_id = id[e]; }
            }
        } else code = 1;
        if (code) {
            const uint32_t k = atomicAdd(n_bad, 1u);
            if (k < max_recs) { VerifyRec v; v.job = j; v.row = r; v.code = code; v.kind = d.kind; v.n_ent = d.n_ent; v.n_mi = d.n_mi; v.n_d = d.n_d; v.bad_id = bad_id; v.off = d.off; recs[k] = v; }
        }
    }
}
// rows [row_lo[j], row_hi[j]) of every job must be well-formed
static int verify_rows(dbgphmm_model* m, const char* what, const RowStore& S, const std::vector<uint32_t>& row_lo, const std::vector<uint32_t>& row_hi) {
    cudaStream_t st = MSET(m).stream;
    const uint32_t J = (uint32_t)S.len.size();
    if (J == 0) return DBGPHMM_OK;
    DevBuf b_lo, b_hi, b_len, b_n, b_recs;
    ST_TRY(dev_upload(b_lo, row_lo, st)); ST_TRY(dev_upload(b_hi, row_hi, st)); ST_TRY(dev_upload(b_len, S.len, st));
    ST_TRY(b_n.alloc(sizeof(uint32_t))); ST_TRY(b_recs.alloc(sizeof(VerifyRec) * 64));
    CUDA_TRY(cudaMemsetAsync(b_n.p, 0, sizeof(uint32_t), st));
    k_verify_rows<<<dim3(64, J), 64, 0, st>>>(J, S.d_desc, S.d_desc0, b_len.as<uint32_t>(), b_lo.as<uint32_t>(), b_hi.as<uint32_t>(), S.arena.base, S.arena.bytes,
                                               S.pool.base ? S.pool.n_slabs : 0, m->N, b_n.as<uint32_t>(), b_recs.as<VerifyRec>(), 64);
    uint32_t n_bad = 0; std::vector<VerifyRec> recs(64);
    CUDA_TRY(cudaMemcpyAsync(&n_bad, b_n.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(recs.data(), b_recs.p, sizeof(VerifyRec) * 64, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    if (!n_bad) return DBGPHMM_OK;
    fprintf(stderr, "[dbgphmm verify] %s: %u malformed rows\n", what, n_bad);
    for (uint32_t k = 0; k < std::min<uint32_t>(n_bad, 24); k++) {
        const VerifyRec& v = recs[k];
        fprintf(stderr, "   job %u (len %u, nd %u) row %u: code %u kind %d n_ent %u n_mi %u n_d %u off %llu bad_id %u\n", v.job, S.len[v.job], S.nd[v.job], v.row, v.code, v.kind,
                v.n_ent, v.n_mi, v.n_d, v.off, v.bad_id);
    }
    dbg_set_error(std::string("DBGPHMM_VERIFY: malformed rows after ") + what);
    return DBGPHMM_ERR_INVALID;
}

#define ST_ARENA_FULL (-100)
// Run a set of sparse jobs, re-running the ones that overflowed the small shared-memory capacity with the big one.
static int run_sparse_jobs_once(dbgphmm_model* m, std::vector<SJob>& sj, SparseIO io, RowStore* store, uint32_t small_cap) {
    if (sj.empty()) return DBGPHMM_OK;
    cudaStream_t st = MSET(m).stream;
    const uint32_t n = (uint32_t)sj.size();
    for (const SJob& j : sj) if (j.dir != sj[0].dir) { dbg_set_error("internal: sparse jobs of one launch must share their direction"); return DBGPHMM_ERR_INVALID; }
    DevBuf b_jobs, b_status, b_final, b_cells;
    ST_TRY(dev_upload(b_jobs, sj, st));
    ST_TRY(b_status.alloc(sizeof(int) * n)); ST_TRY(b_final.alloc(sizeof(XF) * n)); ST_TRY(b_cells.alloc(sizeof(unsigned long long) * n));
    io.status = b_status.as<int>(); io.final_scalar = b_final.as<XF>(); io.cells = b_cells.as<unsigned long long>();
    if (verify_enabled()) CUDA_TRY(cudaMemsetAsync(b_status.p, 0x7f, sizeof(int) * n, st));   // a status no CTA wrote is then not SJ_OK
    std::vector<int> status(n);
    std::vector<XF> fin(n);
    std::vector<unsigned long long> cells(n);
    std::vector<uint32_t> todo(n);
    for (uint32_t i = 0; i < n; i++) todo[i] = i;
    uint32_t caps[3] = {small_cap, 256, 832};
    // first pass: jobs whose rows outgrow small_cap are carried on by a concurrent rescue launch (sparse.cu) instead of a re-run
    const uint32_t rescue_cap = sparse_rescue_cap(small_cap);
    DevBuf b_ctl, b_items, b_hand, b_rows;
    if (rescue_cap) {
        ST_TRY(b_ctl.alloc(sizeof(uint32_t) * 4)); ST_TRY(b_items.alloc(sizeof(uint32_t) * n)); ST_TRY(b_hand.alloc(sizeof(SHandoff) * n));
        ST_TRY(b_rows.alloc((size_t)32 * small_cap * n));
        CUDA_TRY(cudaMemsetAsync(b_ctl.p, 0, sizeof(uint32_t) * 4, st));
        CUDA_TRY(cudaMemsetAsync(b_items.p, 0xff, sizeof(uint32_t) * n, st));
        io.rq_ctl = b_ctl.as<uint32_t>(); io.rq_items = b_items.as<uint32_t>(); io.rq_hand = b_hand.as<SHandoff>(); io.rq_rows = b_rows.as<char>();
    }
    for (int pass = 0; pass < 3 && !todo.empty(); pass++) {
        if (pass > 0 && (caps[pass] <= caps[pass - 1] || (pass == 1 && rescue_cap >= caps[1]))) continue;   // (the rescue launch was pass 1)
        std::vector<SJob> cur(todo.size());
        for (size_t i = 0; i < todo.size(); i++) cur[i] = sj[todo[i]];
        CUDA_TRY(cudaMemcpyAsync(b_jobs.p, cur.data(), sizeof(SJob) * cur.size(), cudaMemcpyHostToDevice, st));
        ST_TRY(sparse_run(m, b_jobs.as<SJob>(), (uint32_t)cur.size(), io, caps[pass], sj[0].dir, pass == 0 ? rescue_cap : 0));
        if (pass == 0 && rescue_cap && getenv("DBGPHMM_TRACE")) {
            uint32_t ctl[4];
            CUDA_TRY(cudaMemcpyAsync(ctl, b_ctl.p, sizeof(ctl), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            fprintf(stderr, "[dbgphmm] sparse pass 0 (cap %u): %u of %zu jobs carried on by the rescue launch (cap %u)\n", caps[0], ctl[0], cur.size(), rescue_cap);
        }
        CUDA_TRY(cudaMemcpyAsync(status.data(), io.status, sizeof(int) * cur.size(), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(fin.data(), io.final_scalar, sizeof(XF) * cur.size(), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cells.data(), io.cells, sizeof(unsigned long long) * cur.size(), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        std::vector<uint32_t> next;
        for (size_t i = 0; i < cur.size(); i++) {
            if (status[i] == SJ_OK) {
                store->h_final[cur[i].out_idx] = fin[i];
                store->cells += cells[i];
            } else if (status[i] == SJ_NEED_BIG && pass < 2) {
                // (on failure the kernel reports where it stopped instead of the cell count: step << 32 | site << 16 | entries)
                if (next.size() < 8 && getenv("DBGPHMM_TRACE"))
                    fprintf(stderr, "[dbgphmm]   job %u (dir %d, %u rows) stopped at step %llu, site %llu, %llu entries\n", todo[i], (int)cur[i].dir, cur[i].n_rows,
                            cells[i] >> 32, (cells[i] >> 16) & 0xffff, cells[i] & 0xffff);
                next.push_back(todo[i]);
            }
            else if (status[i] == SJ_OOM) return ST_ARENA_FULL;
            else if (status[i] == 0x7f7f7f7f) {
                uint32_t ctl[4] = {0, 0, 0, 0};
                if (io.rq_ctl) cudaMemcpy(ctl, io.rq_ctl, sizeof(ctl), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[dbgphmm verify] sparse pass %d: job %u (dir %d, %u rows) was never finished by any CTA ; queue: pushed %u taken %u primaries gone %u next job %u of %zu\n",
                        pass, todo[i], (int)cur[i].dir, cur[i].n_rows, ctl[0], ctl[1], ctl[2], ctl[3], cur.size());
                dbg_set_error("DBGPHMM_VERIFY: a sparse job was never finished"); return DBGPHMM_ERR_INVALID;
            }
            else { dbg_set_error("a sparse row exceeded MAX_ACTIVE_NODES entries (the reference panics: insufficient capacity)"); return DBGPHMM_ERR_CAPACITY; }
        }
        if (!next.empty() && getenv("DBGPHMM_TRACE")) fprintf(stderr, "[dbgphmm] sparse pass %d (cap %u): %zu of %zu jobs need a larger capacity\n", pass, caps[pass], next.size(), cur.size());
        todo.swap(next);
    }
    if (getenv("DBGPHMM_TRACE") && io.arena_cursor) {
        unsigned long long cur = 0; uint64_t rows = 0;
        CUDA_TRY(cudaMemcpy(&cur, io.arena_cursor, sizeof(cur), cudaMemcpyDeviceToHost));
        for (auto& j : sj) rows += j.n_rows;
        fprintf(stderr, "[dbgphmm] sparse arena: %.1f MB used of %.1f MB (%.0f B per row over %llu rows)\n", cur / 1e6, io.arena_bytes / 1e6, rows ? (double)cur / rows : 0.0,
                (unsigned long long)rows);
    }
    return DBGPHMM_OK;
}

// Bytes of sparse rows: the arena is sized for typical rows (top-n rows of n_active = 40 hold ~50 entries = ~1.7 KB on the C3
// workload ; this allows 48 B x n_active + 256) and the phase is repeated once with the bound below if a batch outgrows it.
uint64_t arena_estimate(uint64_t n_rows, uint32_t n_active, bool ratio) {
    uint64_t per_row = ratio ? 2048 : (uint64_t)n_active * 48 + 256;
    uint64_t slack = (uint64_t)64 * SPARSE_PAGE_BYTES;
    if (const char* e = getenv("DBGPHMM_ARENA_EST_PCT")) {   // tests: under-estimate on purpose, so that the phase is repeated with the upper bound
        const int pct = atoi(e);
        if (pct > 0 && pct < 100) { per_row = per_row * (uint64_t)pct / 100; slack = 2 * SPARSE_PAGE_BYTES; }
    }
    uint64_t b = n_rows * per_row + slack;
    return (b + 255) & ~(uint64_t)255;
}
static uint64_t arena_upper(uint64_t n_rows, uint32_t n_active, bool ratio) {
    uint64_t per_row = ratio ? (uint64_t)MAX_ACTIVE * 34 + 64 : (uint64_t)n_active * 7 * 34 + 256;   // nodes + children + 5 Del rounds of new children
    uint64_t b = n_rows * per_row + (uint64_t)64 * SPARSE_PAGE_BYTES;
    return (b + 255) & ~(uint64_t)255;
}
static int run_sparse_jobs(dbgphmm_model* m, std::vector<SJob>& sj, SparseIO io, RowStore* store, uint32_t small_cap, uint64_t upper_bytes) {
    const uint64_t cells0 = store->cells;
    int st = run_sparse_jobs_once(m, sj, io, store, small_cap);
    if (st != ST_ARENA_FULL) return st;
    if (io.arena_bytes >= upper_bytes || getenv("DBGPHMM_ARENA_MAX_BYTES")) { dbg_set_error("sparse row arena exhausted"); return DBGPHMM_ERR_OOM; }
    if (getenv("DBGPHMM_TRACE")) fprintf(stderr, "[dbgphmm] sparse arena of %.1f MB exhausted: repeating the phase with %.1f MB\n", io.arena_bytes / 1e6, upper_bytes / 1e6);
    CUDA_TRY(cudaStreamSynchronize(MSET(m).stream));
    cache_free(store->arena.base); cache_free(store->arena.cursor); store->arena = SparseArena();
    ST_TRY(alloc_arena(store->arena, upper_bytes, MSET(m).stream));
    io.arena = store->arena.base; io.arena_bytes = store->arena.bytes; io.arena_cursor = store->arena.cursor;
    store->cells = cells0;
    st = run_sparse_jobs_once(m, sj, io, store, small_cap);
    if (st == ST_ARENA_FULL) { dbg_set_error("sparse row arena exhausted"); return DBGPHMM_ERR_OOM; }
    return st;
}

// Stream strategy, top-n jobs: the cells of the last dense row that the first sparse row can read are gathered into a small per-job
// list right after the top-n selection, so that the ping-pong slabs are dead before the sparse phase starts -- and the dense warm-up
// can run in GROUPS of jobs that share one pool of slabs (PhaseOpts::group), which decouples the number of jobs of the sparse phase
// from the 2 x 28 B x N a job's slabs take.  Gathering alone: DBGPHMM_GATHER=1 ; groups: DBGPHMM_DENSE_GROUP.
static bool gather_enabled() { const char* e = getenv("DBGPHMM_GATHER"); return e && e[0] == '1'; }
struct GatherBufs {
    DevBuf cells, cnt, slabs, ovf;
    uint32_t cap = 0;
    bool on = false;
};
static int gather_init(dbgphmm_model* m, uint32_t J, GatherBufs* g) {
    g->cap = sparse_gather_cap(m, m->params.n_active_nodes);
    ST_TRY(g->cells.alloc((size_t)J * 32 * g->cap)); ST_TRY(g->cnt.alloc(sizeof(uint32_t) * std::max<uint32_t>(J, 1))); ST_TRY(g->ovf.alloc(sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(g->cnt.p, 0, sizeof(uint32_t) * std::max<uint32_t>(J, 1), MSET(m).stream));
    CUDA_TRY(cudaMemsetAsync(g->ovf.p, 0, sizeof(int), MSET(m).stream));
    g->on = true;
    return DBGPHMM_OK;
}
// slab_of_job[i]: the last dense row of job g0 + i (~0: none)
static int gather_group(dbgphmm_model* m, int dir, uint32_t g0, const std::vector<uint64_t>& slab_of_job, const uint32_t* d_top_ids, const uint32_t* d_top_cnt,
                        const DensePool& pool, GatherBufs* g) {
    cudaStream_t st = MSET(m).stream;
    ST_TRY(dev_upload(g->slabs, slab_of_job, st));
    ST_TRY(sparse_gather_prev0(m, dir, g0, (uint32_t)slab_of_job.size(), d_top_ids, d_top_cnt, g->slabs.as<uint64_t>(), pool.base, pool.slab_bytes, pool.Np, g->cap,
                               g->cells.as<char>(), g->cnt.as<uint32_t>(), g->ovf.as<int>()));
    int ovf = 0;
    CUDA_TRY(cudaMemcpyAsync(&ovf, g->ovf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // (also: the group's slabs may be overwritten by the next group from here on)
    if (ovf) { dbg_set_error("internal: gathered row outgrew its bound"); return DBGPHMM_ERR_INVALID; }
    return DBGPHMM_OK;
}

// ================================================================================================ forward
int run_forward(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, int kind, const PhaseOpts& opt,
                const DevMappings* dmap, RowStore* out) {
    const bool keep_rows = opt.keep_rows, store_sparse = opt.store_sparse;
    HostTrace tr_all("run_forward");
    HostTrace* tr_setup = new HostTrace("  fwd setup");
    cudaStream_t st = MSET(m).stream;
    const uint32_t J = (uint32_t)jobs.size(), N = m->N, W = m->params.n_warmup;
    out->dir = 0; out->dense_kept = keep_rows;
    out->desc0.resize(J); out->len.resize(J); out->nd.assign(J, 0); out->h_final.assign(J, xf_zero());
    uint64_t tot_rows = 0;
    for (uint32_t j = 0; j < J; j++) { out->desc0[j] = tot_rows; out->len[j] = jobs[j].len; tot_rows += jobs[j].len; }
    out->n_desc = tot_rows;
    out->d_desc = (RowDesc*)cache_alloc(sizeof(RowDesc) * std::max<uint64_t>(tot_rows, 1));
    if (!out->d_desc) { dbg_set_error("out of device memory for row descriptors"); return DBGPHMM_ERR_OOM; }
    out->d_final = (XF*)cache_alloc(sizeof(XF) * std::max<uint32_t>(J, 1));
    out->d_desc0 = (uint64_t*)cache_alloc(sizeof(uint64_t) * std::max<uint32_t>(J, 1));
    if (!out->d_final || !out->d_desc0) { dbg_set_error("out of device memory"); return DBGPHMM_ERR_OOM; }
    CUDA_TRY(cudaMemcpyAsync(out->d_desc0, out->desc0.data(), sizeof(uint64_t) * J, cudaMemcpyHostToDevice, st));
    if (verify_enabled()) CUDA_TRY(cudaMemsetAsync(out->d_desc, 0xff, sizeof(RowDesc) * std::max<uint64_t>(tot_rows, 1), st));
    // ---- dense phase layout
    // top-n jobs whose dense rows are not kept: the first sparse row reads gathered cells, and the warm-up may run in groups of G jobs
    // that share one pool of slabs (slab indices are then relative to the group)
    const bool use_gather = kind == DBGPHMM_FWD_SPARSE && !keep_rows && !opt.dense_only && !opt.step && (gather_enabled() || opt.group > 0 || opt.force_gather);
    const uint32_t G = (use_gather && opt.group > 0 && opt.group < J) ? opt.group : std::max<uint32_t>(J, 1);
    std::vector<uint32_t> nd_max(J);
    std::vector<DJob> dj(J);
    uint64_t n_slabs = 0, slab_cur = 0; uint32_t steps = 0;
    for (uint32_t j = 0; j < J; j++) {
        uint32_t n = jobs[j].len;
        if (j % G == 0) slab_cur = 0;
        nd_max[j] = kind == DBGPHMM_FWD_DENSE ? n : (kind == DBGPHMM_FWD_MAPPING ? 0 : std::min(n, W));
        DJob& d = dj[j];
        d.x = jobs[j].x; d.len = n; d.base_off = jobs[j].base_off; d.n_steps = nd_max[j]; d.first_row = 0;
        d.prev0_kind = PREV_F_INIT; d.prev0_slab = 0; d.slab0 = slab_cur; d.slab_mod = keep_rows ? 0 : 2; d.desc0 = out->desc0[j];
        d.active_idx = kind == DBGPHMM_FWD_SPARSE_RATIO ? (int32_t)j : -1;
        slab_cur += keep_rows ? nd_max[j] : std::min<uint32_t>(nd_max[j], 2);
        n_slabs = std::max(n_slabs, slab_cur);
        steps = std::max(steps, nd_max[j]);
    }
    out->slab0.resize(J);
    for (uint32_t j = 0; j < J; j++) out->slab0[j] = dj[j].slab0;
    { HostTrace t("  fwd alloc_pool"); ST_TRY(alloc_pool(out->pool, N, n_slabs)); }
    DevBuf b_dj, b_active, b_nd, b_len, b_part, b_top_ids, b_top_cnt, b_reqs;
    ST_TRY(dev_upload(b_dj, dj, st));
    std::vector<int> h_active(J, 1);
    ST_TRY(dev_upload(b_active, h_active, st));
    ST_TRY(dev_upload(b_nd, nd_max, st));
    ST_TRY(dev_upload(b_len, out->len, st));
    // two rows per launch when nothing reads the intermediate rows (ping-pong slabs, no per-row products, fixed warm-up)
    const bool can_pair = !keep_rows && !opt.step && (kind == DBGPHMM_FWD_SPARSE || kind == DBGPHMM_FWD_DENSE) && steps >= 2 && dense_can_pair(m);
    std::vector<uint8_t> job_paired(J, 0);   // the job's rows went through the two-rows-per-launch kernel (decides where its last row lies)
    ST_TRY(b_part.alloc(sizeof(XF) * (size_t)J * std::max<size_t>(m->fwd.n_chunks, can_pair ? 2 * (size_t)dense_pair_tiles(m, 0) : 0)));
    DevBuf b_wl, b_redo;
    ST_TRY(b_redo.alloc(sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(b_redo.p, 0, sizeof(int), st));
    ST_TRY(b_wl.alloc(sizeof(unsigned long long) * ((size_t)J * m->fwd.n_chunks + 1)));
    ST_TRY(b_top_ids.alloc(sizeof(uint32_t) * (size_t)J * MAX_ACTIVE));
    ST_TRY(b_top_cnt.alloc(sizeof(uint32_t) * J));
    CUDA_TRY(cudaMemsetAsync(b_top_cnt.p, 0, sizeof(uint32_t) * J, st));
    const int* d_active = kind == DBGPHMM_FWD_SPARSE_RATIO ? b_active.as<int>() : nullptr;
    std::vector<SelectReq> reqs(J);
    ST_TRY(b_reqs.alloc(sizeof(SelectReq) * J));
    auto slab_of_h = [&](uint32_t j, uint32_t s) { return job_paired[j] ? dj[j].slab0 + ((s >> 1) & 1) : dj[j].slab0 + (dj[j].slab_mod ? (s % dj[j].slab_mod) : s); };
    GatherBufs gat;
    if (use_gather) ST_TRY(gather_init(m, J, &gat));
    delete tr_setup;
    {
        HostTrace t("  fwd dense phase");
        EvTimer tm(st, &g_times.dense_ms);
        if (kind != DBGPHMM_FWD_SPARSE_RATIO) out->nd = nd_max;
        for (uint32_t g0 = 0; g0 < J; g0 += G) {
            const uint32_t g1 = std::min(J, g0 + G), Jg = g1 - g0;
            const DJob* d_gj = b_dj.as<DJob>() + g0;
            uint32_t gsteps = 0;
            for (uint32_t j = g0; j < g1; j++) gsteps = std::max(gsteps, nd_max[j]);
            bool paired = can_pair && gsteps >= 2;
            if (paired) {
                for (uint32_t s = 0; s < gsteps; s += 2) {
                    uint64_t live = 0;
                    for (uint32_t j = g0; j < g1; j++) live += (s < nd_max[j]) + (s + 1 < nd_max[j]);
                    ST_TRY(dense_forward_pair(m, out->pool, d_gj, Jg, s, d_bases, out->d_desc, b_part.as<XF>(), b_redo.as<int>(), live * N, s + 1 < gsteps));
                }
                int redo = 0;
                CUDA_TRY(cudaMemcpyAsync(&redo, b_redo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                if (redo) {   // some tile's exponent range does not fit a two-row frame: single-row steps from the start
                    if (getenv("DBGPHMM_TRACE")) fprintf(stderr, "[dbgphmm] forward dense phase repeated with single-row steps\n");
                    paired = false;
                    CUDA_TRY(cudaMemsetAsync(b_redo.p, 0, sizeof(int), st));
                }
            }
            for (uint32_t j = g0; j < g1; j++) job_paired[j] = paired ? 1 : 0;
            for (uint32_t s = 0; s < gsteps && !paired; s++) {
                uint64_t live = 0;  // jobs that (may) compute row s: the algorithmic cells of this launch
                for (uint32_t j = g0; j < g1; j++) live += s < nd_max[j];
                ST_TRY(dense_forward_step(m, out->pool, d_gj, Jg, s, d_bases, out->d_desc, d_active, b_part.as<XF>(), b_wl.as<unsigned long long>(), live * N));
                if (opt.step) ST_TRY(step_products(m, *opt.step, out->pool, d_gj, Jg, s, 0, g0));
                if (kind == DBGPHMM_FWD_SPARSE_RATIO) {   // (never grouped: g0 == 0, g1 == J)
                    // top_nodes_by_score_ratio of row s for every job still dense (forward.rs:112-116)
                    uint32_t nr = 0;
                    for (uint32_t j = 0; j < J; j++)
                        if (s < nd_max[j]) { SelectReq& r = reqs[nr++]; r.slab = slab_of_h(j, s); r.k = MAX_ACTIVE; r.by_ratio = 1; r.ratio = m->params.active_node_max_ratio; r.active_idx = (int32_t)j; r.out = j; }
                    CUDA_TRY(cudaMemcpyAsync(b_reqs.p, reqs.data(), sizeof(SelectReq) * nr, cudaMemcpyHostToDevice, st));
                    ST_TRY(dense_select(m, out->pool, b_reqs.as<SelectReq>(), nr, d_active, b_top_ids.as<uint32_t>(), b_top_cnt.as<uint32_t>()));
                    k_ratio_decide<<<(J + 127) / 128, 128, 0, st>>>(J, s, W, m->params.warmup_threshold, b_len.as<uint32_t>(), b_top_cnt.as<uint32_t>(),
                                                                     b_active.as<int>(), b_nd.as<uint32_t>());
                    COUNT_LAUNCH();
                    CUDA_TRY(cudaStreamSynchronize(st));  // reqs is reused next step
                }
            }
            if (kind == DBGPHMM_FWD_SPARSE && !opt.dense_only) {  // top_nodes(n_active) of the last dense row (forward.rs:115)
                uint32_t nr = 0;
                for (uint32_t j = g0; j < g1; j++)
                    if (jobs[j].len > out->nd[j]) { SelectReq& r = reqs[nr++]; r.slab = slab_of_h(j, out->nd[j] - 1); r.k = m->params.n_active_nodes; r.by_ratio = 0; r.ratio = 0; r.active_idx = -1; r.out = j; }
                CUDA_TRY(cudaMemcpyAsync(b_reqs.p, reqs.data(), sizeof(SelectReq) * nr, cudaMemcpyHostToDevice, st));
                ST_TRY(dense_select(m, out->pool, b_reqs.as<SelectReq>(), nr, nullptr, b_top_ids.as<uint32_t>(), b_top_cnt.as<uint32_t>()));
                CUDA_TRY(cudaStreamSynchronize(st));   // (reqs is reused by the next group)
                if (use_gather) {
                    std::vector<uint64_t> gs(Jg, ~0ull);
                    for (uint32_t j = g0; j < g1; j++) if (jobs[j].len > out->nd[j] && out->nd[j] > 0) gs[j - g0] = slab_of_h(j, out->nd[j] - 1);
                    ST_TRY(gather_group(m, 0, g0, gs, b_top_ids.as<uint32_t>(), b_top_cnt.as<uint32_t>(), out->pool, &gat));
                }
            }
        }
        launch_timer_flush();
        if (kind == DBGPHMM_FWD_SPARSE_RATIO) {
            CUDA_TRY(cudaMemcpyAsync(out->nd.data(), b_nd.p, sizeof(uint32_t) * J, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        if (opt.dense_only) { HostTrace t2("  fwd dense_only tail"); CUDA_TRY(cudaStreamSynchronize(st)); cache_free(out->pool.base); out->pool.base = nullptr; return DBGPHMM_OK; }
        if (use_gather) { cache_free(out->pool.base); out->pool.base = nullptr; }   // the ping-pong slabs are dead already
    }
    for (uint32_t j = 0; j < J; j++) { out->cells += (uint64_t)out->nd[j] * N; g_times.dense_cells += (uint64_t)out->nd[j] * N; }
    if (opt.after_dense) opt.after_dense();
    // ---- sparse phase
    std::vector<SJob> sj;
    uint64_t sparse_rows = 0;
    std::vector<uint8_t> dense_final(J, 0);
    for (uint32_t j = 0; j < J; j++) {
        uint32_t n = jobs[j].len, nd = out->nd[j];
        if (nd >= n) { dense_final[j] = n > 0; continue; }
        SJob s{};
        s.x = jobs[j].x; s.len = n; s.base_off = jobs[j].base_off; s.dir = 0;
        s.mode = kind == DBGPHMM_FWD_SPARSE ? SP_TOPN : (kind == DBGPHMM_FWD_SPARSE_RATIO ? SP_RATIO : SP_MAPPING);
        s.row_begin = (int32_t)nd; s.n_rows = n - nd;
        s.prev0_kind = nd == 0 ? SPREV_F_INIT : (gat.on ? SPREV_GATHER : SPREV_DENSE);
        s.prev0_slab = nd == 0 ? 0 : slab_of_h(j, nd - 1);
        s.top0 = j; s.desc0 = out->desc0[j]; s.fdesc0 = 0; s.map_row0 = jobs[j].map_row0;
        s.store = store_sparse ? 1 : 0; s.active_idx = -1; s.out_idx = j;
        sj.push_back(s); sparse_rows += s.n_rows;
    }
    {
        HostTrace t("  fwd sparse phase");
        EvTimer tm(st, &g_times.sparse_ms);
        ST_TRY(alloc_arena(out->arena, store_sparse ? arena_estimate(sparse_rows, m->params.n_active_nodes, kind == DBGPHMM_FWD_SPARSE_RATIO) : 256, st));
        SparseIO io{};
        io.bases = d_bases; io.desc = out->d_desc; io.fdesc = nullptr; io.farena = nullptr;
        io.top_ids = b_top_ids.as<uint32_t>(); io.top_cnt = b_top_cnt.as<uint32_t>();
        io.map_row_off = dmap ? dmap->row_off : nullptr; io.map_nodes = dmap ? dmap->nodes : nullptr;
        io.pool = out->pool.base; io.slab_bytes = out->pool.slab_bytes; io.Np = out->pool.Np;
        io.gather = gat.cells.as<char>(); io.gather_cap = gat.cap; io.gather_cnt = gat.cnt.as<uint32_t>();
        io.arena = out->arena.base; io.arena_bytes = out->arena.bytes; io.arena_cursor = out->arena.cursor; io.active = nullptr;
        ST_TRY(run_sparse_jobs(m, sj, io, out, kind == DBGPHMM_FWD_MAPPING ? 64 : sparse_default_cap(),
                               store_sparse ? arena_upper(sparse_rows, m->params.n_active_nodes, kind == DBGPHMM_FWD_SPARSE_RATIO) : 0));
    }
    if (!keep_rows) { cache_free(out->pool.base); out->pool.base = nullptr; }  // ping-pong slabs are dead now
    // final e of jobs whose last row is dense
    DevBuf b_take, b_desc0;
    ST_TRY(dev_upload(b_take, dense_final, st)); ST_TRY(dev_upload(b_desc0, out->desc0, st));
    CUDA_TRY(cudaMemcpyAsync(out->d_final, out->h_final.data(), sizeof(XF) * J, cudaMemcpyHostToDevice, st));
    k_final_from_desc<<<(J + 127) / 128, 128, 0, st>>>(J, out->d_desc, b_desc0.as<uint64_t>(), b_len.as<uint32_t>(), b_take.as<uint8_t>(), 0, out->d_final);
    COUNT_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(out->h_final.data(), out->d_final, sizeof(XF) * J, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    if (verify_enabled()) {
        std::vector<uint32_t> lo(J, 0), hi(J);
        for (uint32_t j = 0; j < J; j++) hi[j] = store_sparse ? jobs[j].len : out->nd[j];
        ST_TRY(verify_rows(m, "run_forward", *out, lo, hi));
    }
    return DBGPHMM_OK;
}

// ================================================================================================ backward
int run_backward(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, int kind, const PhaseOpts& opt,
                 const DevMappings* dmap, const RowStore* fwd, RowStore* out) {
    const bool keep_rows = opt.keep_rows;
    HostTrace tr_all("run_backward");
    cudaStream_t st = MSET(m).stream;
    const uint32_t J = (uint32_t)jobs.size(), N = m->N, W = m->params.n_warmup;
    out->dir = 1; out->dense_kept = keep_rows;
    out->desc0.resize(J); out->len.resize(J); out->nd.assign(J, 0); out->h_final.assign(J, xf_zero());
    out->bdense_lo.assign(J, -1); out->bdense_hi.assign(J, -1);
    uint64_t tot_rows = 0;
    for (uint32_t j = 0; j < J; j++) { out->desc0[j] = tot_rows; out->len[j] = jobs[j].len; tot_rows += jobs[j].len; }
    out->n_desc = tot_rows;
    out->d_desc = (RowDesc*)cache_alloc(sizeof(RowDesc) * std::max<uint64_t>(tot_rows, 1));
    if (!out->d_desc) { dbg_set_error("out of device memory for row descriptors"); return DBGPHMM_ERR_OOM; }
    out->d_final = (XF*)cache_alloc(sizeof(XF) * std::max<uint32_t>(J, 1));
    out->d_desc0 = (uint64_t*)cache_alloc(sizeof(uint64_t) * std::max<uint32_t>(J, 1));
    if (!out->d_final || !out->d_desc0) { dbg_set_error("out of device memory"); return DBGPHMM_ERR_OOM; }
    CUDA_TRY(cudaMemcpyAsync(out->d_desc0, out->desc0.data(), sizeof(uint64_t) * J, cudaMemcpyHostToDevice, st));
    if (verify_enabled()) CUDA_TRY(cudaMemsetAsync(out->d_desc, 0xff, sizeof(RowDesc) * std::max<uint64_t>(tot_rows, 1), st));
    // dense rows of job j: [lo, hi]; they are computed from hi down to lo
    std::vector<DJob> dj(J);
    uint64_t n_slabs = 0, slab_cur = 0; uint32_t steps = 0;
    const bool sparse_first = kind == DBGPHMM_BWD_BY_FORWARD;
    // (see run_forward) top-n jobs whose dense rows are not kept: gathered first-row inputs, warm-up in groups of G jobs
    const bool use_gather = kind == DBGPHMM_BWD_SPARSE && !keep_rows && !opt.step && (gather_enabled() || opt.group > 0 || opt.force_gather);
    const uint32_t G = (use_gather && opt.group > 0 && opt.group < J) ? opt.group : std::max<uint32_t>(J, 1);
    for (uint32_t j = 0; j < J; j++) {
        int n = (int)jobs[j].len;
        int lo = -1, hi = -1;
        if (j % G == 0) slab_cur = 0;
        if (n > 0) {
            if (kind == DBGPHMM_BWD_DENSE) { lo = 0; hi = n - 1; }
            else if (kind == DBGPHMM_BWD_SPARSE) { hi = n - 1; lo = std::max(0, n - (int)W); }
            else if (kind == DBGPHMM_BWD_BY_FORWARD) { lo = 0; hi = std::min<int>((int)fwd->nd[j], n - 1); }  // row r dense iff r == 0 or F row r-1 dense
        }
        out->bdense_lo[j] = lo; out->bdense_hi[j] = hi;
        uint32_t nd = hi >= 0 ? (uint32_t)(hi - lo + 1) : 0;
        out->nd[j] = nd;
        DJob& d = dj[j];
        d.x = jobs[j].x; d.len = n; d.base_off = jobs[j].base_off; d.n_steps = nd; d.first_row = hi;
        d.prev0_kind = PREV_B_INIT; d.prev0_slab = 0; d.slab0 = slab_cur; d.slab_mod = keep_rows ? 0 : 2; d.desc0 = out->desc0[j]; d.active_idx = -1;
        uint64_t need = keep_rows ? nd : std::min<uint32_t>(nd, 2);
        if (sparse_first && nd > 0 && hi < n - 1) { d.prev0_kind = PREV_SLAB; d.prev0_slab = slab_cur + need; need += 1; }  // scattered sparse row hi+1
        slab_cur += need;
        n_slabs = std::max(n_slabs, slab_cur);
        steps = std::max(steps, nd);
    }
    out->slab0.resize(J);
    for (uint32_t j = 0; j < J; j++) out->slab0[j] = dj[j].slab0;
    if (opt.before_dense) opt.before_dense();
    { HostTrace t("  bwd alloc_pool"); ST_TRY(alloc_pool(out->pool, N, n_slabs)); }
    DevBuf b_dj, b_len, b_part, b_top_ids, b_top_cnt, b_reqs;
    ST_TRY(dev_upload(b_dj, dj, st));
    ST_TRY(dev_upload(b_len, out->len, st));
    // two rows per launch when nothing reads the intermediate rows (see run_forward)
    const bool can_pair = !keep_rows && !opt.step && kind == DBGPHMM_BWD_SPARSE && steps >= 2 && dense_can_pair(m);
    std::vector<uint8_t> job_paired(J, 0);
    ST_TRY(b_part.alloc(sizeof(XF) * 2 * (size_t)J * std::max<size_t>(m->bwd.n_chunks, can_pair ? 2 * (size_t)dense_pair_tiles(m, 1) : 0)));
    DevBuf b_wl, b_redo;
    ST_TRY(b_redo.alloc(sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(b_redo.p, 0, sizeof(int), st));
    ST_TRY(b_wl.alloc(sizeof(unsigned long long) * ((size_t)J * m->bwd.n_chunks + 1)));
    ST_TRY(b_top_ids.alloc(sizeof(uint32_t) * (size_t)J * MAX_ACTIVE));
    ST_TRY(b_top_cnt.alloc(sizeof(uint32_t) * J));
    CUDA_TRY(cudaMemsetAsync(b_top_cnt.p, 0, sizeof(uint32_t) * J, st));
    auto slab_of_h = [&](uint32_t j, uint32_t s) { return job_paired[j] ? dj[j].slab0 + ((s >> 1) & 1) : dj[j].slab0 + (dj[j].slab_mod ? (s % dj[j].slab_mod) : s); };
    GatherBufs gat;
    if (use_gather) ST_TRY(gather_init(m, J, &gat));

    auto dense_phase = [&](uint32_t g0, uint32_t g1) -> int {   // the dense rows of jobs g0 .. g1 - 1
        HostTrace t("  bwd dense phase");
        EvTimer tm(st, &g_times.dense_ms);
        const uint32_t Jg = g1 - g0;
        const DJob* d_gj = b_dj.as<DJob>() + g0;
        uint32_t gsteps = 0;
        for (uint32_t j = g0; j < g1; j++) gsteps = std::max(gsteps, out->nd[j]);
        bool paired = can_pair && gsteps >= 2;
        if (paired) {
            for (uint32_t s = 0; s < gsteps; s += 2) {
                uint64_t live = 0;
                for (uint32_t j = g0; j < g1; j++) live += (s < out->nd[j]) + (s + 1 < out->nd[j]);
                ST_TRY(dense_backward_pair(m, out->pool, d_gj, Jg, s, d_bases, out->d_desc, b_part.as<XF>(), b_redo.as<int>(), live * N, s + 1 < gsteps));
            }
            int redo = 0;
            CUDA_TRY(cudaMemcpyAsync(&redo, b_redo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (redo) {
                if (getenv("DBGPHMM_TRACE")) fprintf(stderr, "[dbgphmm] backward dense phase repeated with single-row steps\n");
                paired = false;
                CUDA_TRY(cudaMemsetAsync(b_redo.p, 0, sizeof(int), st));
            }
        }
        for (uint32_t j = g0; j < g1; j++) job_paired[j] = paired ? 1 : 0;
        for (uint32_t s = 0; s < gsteps && !paired; s++) {
            uint64_t live = 0;
            for (uint32_t j = g0; j < g1; j++) live += s < out->nd[j];
            ST_TRY(dense_backward_step(m, out->pool, d_gj, Jg, s, d_bases, out->d_desc, nullptr, b_part.as<XF>(), b_wl.as<unsigned long long>(), live * N));
            if (opt.step) ST_TRY(step_products(m, *opt.step, out->pool, d_gj, Jg, s, 1, g0));
        }
        launch_timer_flush();
        for (uint32_t j = g0; j < g1; j++) { out->cells += (uint64_t)out->nd[j] * N; g_times.dense_cells += (uint64_t)out->nd[j] * N; }
        return DBGPHMM_OK;
    };
    auto sparse_phase = [&]() -> int {
        HostTrace t("  bwd sparse phase");
        EvTimer tm(st, &g_times.sparse_ms);
        std::vector<SJob> sj;
        uint64_t sparse_rows = 0;
        for (uint32_t j = 0; j < J; j++) {
            int n = (int)jobs[j].len, lo = out->bdense_lo[j], hi = out->bdense_hi[j];
            if (n == 0) continue;
            SJob s{};
            s.x = jobs[j].x; s.len = n; s.base_off = jobs[j].base_off; s.dir = 1;
            s.desc0 = out->desc0[j]; s.fdesc0 = fwd ? fwd->desc0[j] : 0; s.map_row0 = jobs[j].map_row0;
            s.store = 1; s.active_idx = -1; s.out_idx = j; s.top0 = j;
            if (kind == DBGPHMM_BWD_SPARSE) {
                if (lo <= 0) continue;  // all rows dense
                s.mode = SP_TOPN; s.row_begin = lo - 1; s.n_rows = (uint32_t)lo; s.prev0_kind = gat.on ? SPREV_GATHER : SPREV_DENSE; s.prev0_slab = slab_of_h(j, out->nd[j] - 1);
            } else if (kind == DBGPHMM_BWD_MAPPING) {
                s.mode = SP_MAPPING; s.row_begin = n - 1; s.n_rows = (uint32_t)n; s.prev0_kind = SPREV_B_INIT;
            } else if (kind == DBGPHMM_BWD_BY_FORWARD) {
                if (hi >= n - 1) continue;  // every row dense
                s.mode = SP_BYFWD; s.row_begin = n - 1; s.n_rows = (uint32_t)(n - 1 - hi); s.prev0_kind = SPREV_B_INIT;
            } else continue;
            sj.push_back(s); sparse_rows += s.n_rows;
        }
        ST_TRY(alloc_arena(out->arena, arena_estimate(sparse_rows, m->params.n_active_nodes, kind == DBGPHMM_BWD_BY_FORWARD), st));
        SparseIO io{};
        io.bases = d_bases; io.desc = out->d_desc; io.fdesc = fwd ? fwd->d_desc : nullptr; io.farena = fwd ? fwd->arena.base : nullptr;
        io.top_ids = b_top_ids.as<uint32_t>(); io.top_cnt = b_top_cnt.as<uint32_t>();
        io.map_row_off = dmap ? dmap->row_off : nullptr; io.map_nodes = dmap ? dmap->nodes : nullptr;
        io.pool = out->pool.base; io.slab_bytes = out->pool.slab_bytes; io.Np = out->pool.Np;
        io.gather = gat.cells.as<char>(); io.gather_cap = gat.cap; io.gather_cnt = gat.cnt.as<uint32_t>();
        io.arena = out->arena.base; io.arena_bytes = out->arena.bytes; io.arena_cursor = out->arena.cursor; io.active = nullptr;
        return run_sparse_jobs(m, sj, io, out, kind == DBGPHMM_BWD_MAPPING ? 64 : sparse_default_cap(),
                               arena_upper(sparse_rows, m->params.n_active_nodes, kind == DBGPHMM_BWD_BY_FORWARD));
    };

    if (!sparse_first) {
        for (uint32_t g0 = 0; g0 < J; g0 += G) {
            const uint32_t g1 = std::min(J, g0 + G);
            ST_TRY(dense_phase(g0, g1));
            if (kind == DBGPHMM_BWD_SPARSE) {  // top_nodes(n_active) of the last dense row (backward.rs:174)
                std::vector<SelectReq> reqs;
                for (uint32_t j = g0; j < g1; j++)
                    if (out->bdense_lo[j] > 0) { SelectReq r{}; r.slab = slab_of_h(j, out->nd[j] - 1); r.k = m->params.n_active_nodes; r.by_ratio = 0; r.active_idx = -1; r.out = j; reqs.push_back(r); }
                ST_TRY(dev_upload(b_reqs, reqs, st));
                ST_TRY(dense_select(m, out->pool, b_reqs.as<SelectReq>(), (uint32_t)reqs.size(), nullptr, b_top_ids.as<uint32_t>(), b_top_cnt.as<uint32_t>()));
                if (use_gather) {
                    std::vector<uint64_t> gs(g1 - g0, ~0ull);
                    for (uint32_t j = g0; j < g1; j++) if (out->bdense_lo[j] > 0) gs[j - g0] = slab_of_h(j, out->nd[j] - 1);
                    ST_TRY(gather_group(m, 1, g0, gs, b_top_ids.as<uint32_t>(), b_top_cnt.as<uint32_t>(), out->pool, &gat));
                }
            }
        }
        if (use_gather) { cache_free(out->pool.base); out->pool.base = nullptr; }   // the ping-pong slabs are dead already
        if (opt.after_dense) {
            // The dense tail (selection kernels) must be over before the other direction's persistent sparse CTAs take the SMs: they are
            // configured for maximum shared memory, and a kernel that wants another L1 / shared split waits for an idle SM -- which a
            // persistent kernel never leaves -- so this direction's sparse launch, queued behind the selection, would start only after the
            // other direction's phase had finished (measured: the two phases did not overlap at all).
            CUDA_TRY(cudaStreamSynchronize(st));
            opt.after_dense();
        }
        ST_TRY(sparse_phase());
    } else {
        ST_TRY(sparse_phase());
        // scatter sparse row hi+1 into the spare slab, then the dense tail hi..0
        std::vector<uint64_t> didx, slabs;
        for (uint32_t j = 0; j < J; j++)
            if (dj[j].prev0_kind == PREV_SLAB) { didx.push_back(out->desc0[j] + out->bdense_hi[j] + 1); slabs.push_back(dj[j].prev0_slab); }
        if (!didx.empty()) {
            DevBuf b_didx, b_slabs;
            ST_TRY(dev_upload(b_didx, didx, st)); ST_TRY(dev_upload(b_slabs, slabs, st));
            dim3 g(std::min<uint32_t>((N + 255) / 256, 256), (uint32_t)didx.size());
            k_scatter_row<<<g, 256, 0, st>>>((uint32_t)didx.size(), b_didx.as<uint64_t>(), b_slabs.as<uint64_t>(), out->d_desc, out->arena.base,
                                             out->pool.base, out->pool.slab_bytes, out->pool.Np, N);
            COUNT_LAUNCH();
            k_scatter_row2<<<(uint32_t)didx.size(), 256, 0, st>>>((uint32_t)didx.size(), b_didx.as<uint64_t>(), b_slabs.as<uint64_t>(), out->d_desc,
                                                                  out->arena.base, out->pool.base, out->pool.slab_bytes, out->pool.Np);
            COUNT_LAUNCH();
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        ST_TRY(dense_phase(0, J));
    }
    if (!keep_rows) { cache_free(out->pool.base); out->pool.base = nullptr; }
    // final mb: row 0 is dense whenever dense rows reach row 0
    std::vector<uint8_t> dense_final(J, 0);
    for (uint32_t j = 0; j < J; j++) dense_final[j] = (out->bdense_lo[j] == 0);
    DevBuf b_take, b_desc0;
    ST_TRY(dev_upload(b_take, dense_final, st)); ST_TRY(dev_upload(b_desc0, out->desc0, st));
    CUDA_TRY(cudaMemcpyAsync(out->d_final, out->h_final.data(), sizeof(XF) * J, cudaMemcpyHostToDevice, st));
    k_final_from_desc<<<(J + 127) / 128, 128, 0, st>>>(J, out->d_desc, b_desc0.as<uint64_t>(), b_len.as<uint32_t>(), b_take.as<uint8_t>(), 1, out->d_final);
    COUNT_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(out->h_final.data(), out->d_final, sizeof(XF) * J, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    if (verify_enabled()) {
        std::vector<uint32_t> lo(J, 0), hi(J);
        for (uint32_t j = 0; j < J; j++) hi[j] = jobs[j].len;
        ST_TRY(verify_rows(m, "run_backward", *out, lo, hi));
    }
    return DBGPHMM_OK;
}

// ================================================================================================ ROI recompute
// mark[j][tile] = 1 for every tile in the dependency cone of the node sets of backward rows 1..W of job j
__global__ void k_roi_mark(uint32_t W, const RowDesc* __restrict__ bdesc, const uint64_t* __restrict__ bdesc0, const uint32_t* __restrict__ len,
                           const char* __restrict__ barena, const uint32_t* __restrict__ tile_of, const uint32_t* __restrict__ roi_off,
                           const uint32_t* __restrict__ roi_tile, uint32_t n_tiles, unsigned char* __restrict__ mark) {
    const uint32_t j = blockIdx.y, t = blockIdx.x + 1;   // backward row t pairs with forward row t-1 < W
    if (t > W || t >= len[j]) return;
    const RowDesc r = bdesc[bdesc0[j] + t];
    if (r.kind != ROW_SPARSE) return;
    const uint32_t* id = (const uint32_t*)(barena + r.off + 24ull * r.n_ent);
    for (uint32_t e = threadIdx.x; e < r.n_ent; e += blockDim.x) {
        uint32_t tl = tile_of[id[e]];
        for (uint32_t a = roi_off[tl]; a < roi_off[tl + 1]; a++) mark[(size_t)j * n_tiles + roi_tile[a]] = 1;
    }
}
__global__ void k_roi_compact(uint32_t n_jobs, uint32_t n_tiles, const unsigned char* __restrict__ mark, unsigned long long* __restrict__ worklist) {
    const size_t total = (size_t)n_jobs * n_tiles;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        if (mark[i]) {
            unsigned long long w = atomicAdd(worklist, 1ull);
            worklist[1 + w] = ((unsigned long long)(i / n_tiles) << 32) | (unsigned long long)(i % n_tiles);
        }
}

// mark the backward tiles whose rows [n - W, n - 1] are needed at the nodes of the sparse forward rows they pair with
__global__ void k_roi_mark_b(uint32_t W, const RowDesc* __restrict__ fdesc, const uint64_t* __restrict__ fdesc0, const uint32_t* __restrict__ len,
                             const char* __restrict__ farena, const uint32_t* __restrict__ tile_of, const uint32_t* __restrict__ roi_off,
                             const uint32_t* __restrict__ roi_tile, uint32_t n_tiles, unsigned char* __restrict__ mark) {
    const uint32_t j = blockIdx.y, s = blockIdx.x;   // backward row n - 1 - s pairs with forward row n - 2 - s
    const uint32_t n = len[j];
    if (s >= W || s + 2 > n) return;
    const RowDesc r = fdesc[fdesc0[j] + (n - 2 - s)];
    if (r.kind != ROW_SPARSE) return;
    const uint32_t* id = (const uint32_t*)(farena + r.off + 24ull * r.n_ent);
    for (uint32_t e = threadIdx.x; e < r.n_ent; e += blockDim.x) {
        uint32_t tl = tile_of[id[e]];
        for (uint32_t a = roi_off[tl]; a < roi_off[tl + 1]; a++) mark[(size_t)j * n_tiles + roi_tile[a]] = 1;
    }
}

// Stream strategy, backward warm-up rows x sparse forward rows: the dense backward rows are recomputed inside the dependency cone of
// the forward rows' node sets and multiplied on the fly (mirror of run_forward_recompute ; lets the main backward pass run two rows
// per launch without ever writing the intermediate rows).
int run_backward_recompute(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, const RowStore& F, const RowStore& B,
                           const StepProducts& sp, uint32_t group) {
    HostTrace tr("run_backward_recompute");
    cudaStream_t st = MSET(m).stream;
    const uint32_t J = (uint32_t)jobs.size(), N = m->N, W = m->params.n_warmup, T = m->bwd.n_chunks;
    ST_TRY(model_ensure_roi(m));
    EvTimer tm(st, &g_times.dense_ms);
    const uint32_t G = (group > 0 && group < J) ? group : std::max<uint32_t>(J, 1);   // jobs per pool of slabs
    std::vector<DJob> dj(J);
    for (uint32_t j = 0; j < J; j++) {
        DJob& d = dj[j];
        d.x = jobs[j].x; d.len = jobs[j].len; d.base_off = jobs[j].base_off; d.n_steps = std::min<uint32_t>(jobs[j].len, W); d.first_row = (int32_t)jobs[j].len - 1;
        d.prev0_kind = PREV_B_INIT; d.prev0_slab = 0; d.slab0 = 2ull * (j % G); d.slab_mod = 2; d.desc0 = B.desc0[j]; d.active_idx = -1;
    }
    DensePool pool;
    pool.Np = (N + 1) & ~1u; pool.slab_bytes = dense_slab_bytes(N); pool.n_slabs = 2ull * std::min(G, J);
    DevBuf b_pool, b_dj, b_len, b_mark, b_wl, b_part;
    ST_TRY(b_pool.alloc(pool.slab_bytes * std::max<uint64_t>(pool.n_slabs, 1))); pool.base = b_pool.as<char>();
    ST_TRY(dev_upload(b_dj, dj, st)); ST_TRY(dev_upload(b_len, F.len, st));
    ST_TRY(b_mark.alloc((size_t)std::min(G, J) * T + 1));
    ST_TRY(b_wl.alloc(sizeof(unsigned long long) * ((size_t)std::min(G, J) * T + 1)));
    ST_TRY(b_part.alloc(sizeof(XF) * 2 * (size_t)std::min(G, J) * T + 16));
    for (uint32_t g0 = 0; g0 < J; g0 += G) {
        const uint32_t Jg = std::min(J, g0 + G) - g0;
        CUDA_TRY(cudaMemsetAsync(b_mark.p, 0, (size_t)Jg * T, st));
        CUDA_TRY(cudaMemsetAsync(b_wl.p, 0, sizeof(unsigned long long), st));
        // No zeroing of the slabs (88 GB per pass on C3 in round 1): a cone tile next to the cone's edge does read cells no tile of this
        // pass wrote, but what it makes of them stays outside the dependency cone of the cells the products read -- the cone is the closure
        // over HALO_HOPS hops per row and n_warmup rows, and a value travels at most HALO_HOPS hops per row.  DBGPHMM_VERIFY=1 poisons
        // the slabs with NaN patterns instead, so that any such dependence would surface as NaN frequencies in the tests that set it.
        if (verify_enabled()) CUDA_TRY(cudaMemsetAsync(pool.base, 0xff, pool.slab_bytes * 2ull * Jg, st));
        {
            dim3 g(W, Jg);
            k_roi_mark_b<<<g, 128, 0, st>>>(W, F.d_desc, F.d_desc0 + g0, b_len.as<uint32_t>() + g0, F.arena.base, m->d_tile_of_b, m->d_roi_off_b, m->d_roi_tile_b, T,
                                            b_mark.as<unsigned char>());
            COUNT_LAUNCH();
            k_roi_compact<<<4 * m->n_sm, 256, 0, st>>>(Jg, T, b_mark.as<unsigned char>(), b_wl.as<unsigned long long>());
            COUNT_LAUNCH();
        }
        for (uint32_t s = 0; s < W; s++) {
            ST_TRY(dense_backward_step_list(m, pool, b_dj.as<DJob>() + g0, s, d_bases, b_part.as<XF>(), b_wl.as<unsigned long long>()));
            ST_TRY(step_products(m, sp, pool, b_dj.as<DJob>() + g0, Jg, s, 1, g0));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}

int run_forward_recompute(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, const RowStore& F, const RowStore& B,
                          const StepProducts& sp, uint32_t group) {
    HostTrace tr("run_forward_recompute");
    cudaStream_t st = MSET(m).stream;
    const uint32_t J = (uint32_t)jobs.size(), N = m->N, W = m->params.n_warmup, T = m->fwd.n_chunks;
    ST_TRY(model_ensure_roi(m));
    EvTimer tm(st, &g_times.dense_ms);
    const uint32_t G = (group > 0 && group < J) ? group : std::max<uint32_t>(J, 1);   // jobs per pool of slabs
    std::vector<DJob> dj(J);
    for (uint32_t j = 0; j < J; j++) {
        DJob& d = dj[j];
        d.x = jobs[j].x; d.len = jobs[j].len; d.base_off = jobs[j].base_off; d.n_steps = std::min<uint32_t>(jobs[j].len, W); d.first_row = 0;
        d.prev0_kind = PREV_F_INIT; d.prev0_slab = 0; d.slab0 = 2ull * (j % G); d.slab_mod = 2; d.desc0 = F.desc0[j]; d.active_idx = -1;
    }
    DensePool pool;
    pool.Np = (N + 1) & ~1u; pool.slab_bytes = dense_slab_bytes(N); pool.n_slabs = 2ull * std::min(G, J);
    DevBuf b_pool, b_dj, b_len, b_mark, b_wl, b_part;
    ST_TRY(b_pool.alloc(pool.slab_bytes * std::max<uint64_t>(pool.n_slabs, 1))); pool.base = b_pool.as<char>();
    ST_TRY(dev_upload(b_dj, dj, st)); ST_TRY(dev_upload(b_len, F.len, st));
    ST_TRY(b_mark.alloc((size_t)std::min(G, J) * T + 1));
    ST_TRY(b_wl.alloc(sizeof(unsigned long long) * ((size_t)std::min(G, J) * T + 1)));
    ST_TRY(b_part.alloc(sizeof(XF) * (size_t)std::min(G, J) * T + 16));
    for (uint32_t g0 = 0; g0 < J; g0 += G) {
        const uint32_t Jg = std::min(J, g0 + G) - g0;
        CUDA_TRY(cudaMemsetAsync(b_mark.p, 0, (size_t)Jg * T, st));
        CUDA_TRY(cudaMemsetAsync(b_wl.p, 0, sizeof(unsigned long long), st));
        // No zeroing of the slabs (88 GB per pass on C3 in round 1): a cone tile next to the cone's edge does read cells no tile of this
        // pass wrote, but what it makes of them stays outside the dependency cone of the cells the products read -- the cone is the closure
        // over HALO_HOPS hops per row and n_warmup rows, and a value travels at most HALO_HOPS hops per row.  DBGPHMM_VERIFY=1 poisons
        // the slabs with NaN patterns instead, so that any such dependence would surface as NaN frequencies in the tests that set it.
        if (verify_enabled()) CUDA_TRY(cudaMemsetAsync(pool.base, 0xff, pool.slab_bytes * 2ull * Jg, st));
        {
            dim3 g(W, Jg);
            k_roi_mark<<<g, 128, 0, st>>>(W, B.d_desc, B.d_desc0 + g0, b_len.as<uint32_t>() + g0, B.arena.base, m->d_tile_of, m->d_roi_off, m->d_roi_tile, T,
                                          b_mark.as<unsigned char>());
            COUNT_LAUNCH();
            k_roi_compact<<<4 * m->n_sm, 256, 0, st>>>(Jg, T, b_mark.as<unsigned char>(), b_wl.as<unsigned long long>());
            COUNT_LAUNCH();
        }
        for (uint32_t s = 0; s < W; s++) {
            ST_TRY(dense_forward_step_list(m, pool, b_dj.as<DJob>() + g0, s, d_bases, F.d_desc, b_part.as<XF>(), b_wl.as<unsigned long long>()));
            ST_TRY(step_products(m, sp, pool, b_dj.as<DJob>() + g0, Jg, s, 0, g0));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return DBGPHMM_OK;
}
