// engine.h — batch orchestration of the dense / sparse phases, row stores, products (internal).
#pragma once
#include <functional>
#include <vector>
#include "dense.h"
#include "sparse.h"

struct HJob {            // one (read, candidate X) unit of work
    uint32_t read;       // index into the reads handle
    uint32_t x;          // parameter set
    uint64_t base_off;   // offset of the read in the device base array
    uint32_t len;
    uint64_t map_row0;   // first row of the read in the mapping CSR (mapping kinds)
};

struct DevMappings {     // device copy of a Mappings CSR with relabelled node ids
    uint64_t* row_off = nullptr;
    uint32_t* nodes = nullptr;
    std::vector<uint64_t> read_off;  // host: first row of each read
    bool owned = false;              // false: a view of the copy the mappings handle keeps (upload_mappings)
    void release();
};

// Rows of one direction for a batch of jobs (PHMMTables of every job).
struct RowStore {
    int dir = 0;                     // 0 forward, 1 backward
    RowDesc* d_desc = nullptr;       // [sum len]
    uint64_t n_desc = 0;
    std::vector<uint64_t> desc0;     // per job
    uint64_t* d_desc0 = nullptr;     // device copy of desc0
    std::vector<uint32_t> len;       // per job
    std::vector<uint32_t> nd;        // per job: number of dense rows (forward: rows [0,nd) ; backward: see engine.cu)
    std::vector<uint64_t> slab0;     // per job: first slab (forward row r -> slab0 + r ; backward row t -> slab0 + (hi - t)), kept rows only
    std::vector<int32_t> bdense_lo, bdense_hi;  // backward: dense rows are [lo, hi] (inclusive), -1/-1 if none
    DensePool pool;
    SparseArena arena;
    bool dense_kept = true;          // false: dense rows were computed in ping-pong slabs and are gone
    XF* d_final = nullptr;           // per job: forward e(last) / backward mb(first)
    std::vector<XF> h_final;
    uint64_t cells = 0;
    void release();
};

// Caching allocator for the large row buffers (cudaMalloc / cudaFree of tens of GB cost ~100 ms each).
void cache_set_stream(cudaStream_t st);   // the stream this host thread's allocations are used on / freed behind (stream-ordered reuse)
void* cache_alloc(size_t bytes);   // nullptr on failure
void cache_free(void* p);
void cache_trim();                 // give every unused block back to the driver
uint64_t cache_unused_bytes();     // bytes of this device held in unused cache blocks
uint64_t model_budget(const dbgphmm_model* m);   // device bytes a bulk call may plan with, from the memory free at the time of the call

struct DevBuf {  // scoped device allocation
    void* p = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cache_free(p); }
    int alloc(size_t bytes) {
        if (p) { cache_free(p); p = nullptr; }
        p = cache_alloc(bytes);
        if (!p) { dbg_set_error("out of device memory"); return DBGPHMM_ERR_OOM; }
        return DBGPHMM_OK;
    }
    template <class T> T* as() { return (T*)p; }
};
template <class T>
static inline int dev_upload(DevBuf& b, const std::vector<T>& h, cudaStream_t st) {
    ST_TRY(b.alloc(h.size() * sizeof(T)));
    if (!h.empty()) CUDA_TRY(cudaMemcpyAsync(b.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return DBGPHMM_OK;
}
struct EvTimer {  // accumulates the device time between construction and destruction on `st`
    cudaEvent_t a = nullptr, b = nullptr; cudaStream_t st; double* acc;
    EvTimer(cudaStream_t s, double* acc_) : st(s), acc(acc_) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    ~EvTimer() {
        cudaEventRecord(b, st); cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b); *acc += ms;
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
};

#include <chrono>
#include <cstdio>
#include <cstdlib>
struct HostTrace {  // DBGPHMM_TRACE=1: wall-clock trace of host-side phases on stderr
    const char* name; std::chrono::steady_clock::time_point t0; bool on;
    HostTrace(const char* n) : name(n), t0(std::chrono::steady_clock::now()) { const char* e = getenv("DBGPHMM_TRACE"); on = e && e[0] == '1'; }
    ~HostTrace() { if (on) fprintf(stderr, "[trace] %-28s %8.2f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); }
};
struct EngineTimes {
    double dense_ms = 0, sparse_ms = 0, product_ms = 0, total_ms = 0;
    uint64_t dense_cells = 0;
    double dense_kernel_ms = 0;        // sum of the k_dense_fwd / k_dense_bwd launch durations (CUDA events around each launch)
    uint64_t dense_kernel_launches = 0;
    uint64_t dense_kernel_cells = 0;   // cells those launches computed
};
// Event pairs around individual launches, resolved lazily (no sync per launch).
void launch_timer_begin(cudaStream_t st);
void launch_timer_end(cudaStream_t st, uint64_t cells);
void launch_timer_flush();
void launch_timer_release();   // destroy this thread's pooled events (helper threads call it before they end)
extern thread_local EngineTimes g_times;   // timings of the last bulk call made by this host thread

// Products taken on the fly, right after a dense row has been computed, against the OTHER direction's stored sparse
// row (forward row r pairs with backward row r+1, table.rs:500-505).  Lets the dense rows live in two ping-pong slabs.
struct StepProducts {
    const RowStore* other = nullptr;  // stored rows of the other direction
    const XF* P = nullptr;            // device, per job: forward full probability
    double* d_freqs = nullptr;        // device, ORIGINAL node order, added to
    int* d_err = nullptr;             // device flag: some P == 0
};
struct PhaseOpts {
    bool keep_rows = true;      // dense rows kept (true) or ping-pong (false)
    bool store_sparse = true;   // sparse rows written to the arena
    bool dense_only = false;    // forward: stop after the dense rows (recompute pass)
    uint32_t group = 0;         // top-n jobs without kept rows: dense warm-up in groups of this many jobs sharing one pool of slabs (0: one group)
    bool force_gather = false;  // top-n jobs without kept rows: the first sparse row reads gathered cells and the slabs are released before the sparse phase
    // hooks of the two-thread bulk path (api.cu): called before the dense rows' slabs are allocated (backward) / once the dense phase is
    // over and, with gathered first-row inputs, its slabs have been released -- right before the sparse phase
    std::function<void()> before_dense, after_dense;
    const StepProducts* step = nullptr;
};

// device view of the mappings for model m: the copy the handle keeps (made on first use with this model, reused afterwards)
int upload_mappings(dbgphmm_model* m, const dbgphmm_mappings* mp, DevMappings* out);
void mappings_release_device(const dbgphmm_mappings* mp);
int run_forward(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, int kind, const PhaseOpts& opt,
                const DevMappings* dmap, RowStore* out);
int run_backward(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, int kind, const PhaseOpts& opt,
                 const DevMappings* dmap, const RowStore* fwd, RowStore* out);
// Recompute the forward warm-up rows only inside the dependency cone of the backward rows' sparse node sets and take the
// (forward dense row, backward sparse row) products there.  F supplies the row scalars of the first pass.
// group: jobs per pool of slabs (0: all of them at once)
int run_forward_recompute(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, const RowStore& F, const RowStore& B,
                          const StepProducts& sp, uint32_t group = 0);
int run_backward_recompute(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, const RowStore& F, const RowStore& B,
                          const StepProducts& sp, uint32_t group = 0);
// d_jobs[0 .. n_jobs) are the jobs job0 .. job0 + n_jobs - 1 of the batch (sp's per-job arrays are indexed by batch position)
int step_products(dbgphmm_model* m, const StepProducts& sp, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s, int dir, uint32_t job0 = 0);
// sum over rows of F (x) B / P into d_freqs (device, ORIGINAL node order, added to); status ZERO_PROB if some P == 0
// mapx.cu: forward_with_mapping_score_only for groups of candidates X of one read
struct MapxGroup { uint64_t base_off; uint32_t len; uint64_t map_row0; uint32_t x0, nx, out0; };
int run_mapx(dbgphmm_model* m, const std::vector<MapxGroup>& groups, const uint8_t* d_bases, const DevMappings& dmap, uint32_t out_stride,
             XF* h_final, std::vector<uint8_t>& failed, uint64_t* cells_out);
int run_products_freqs(dbgphmm_model* m, const std::vector<HJob>& jobs, const RowStore& F, const RowStore& B, double* d_freqs);
int run_products_edge_freqs(dbgphmm_model* m, const std::vector<HJob>& jobs, const RowStore& F, const RowStore& B, const uint8_t* d_bases,
                            double* d_edge, double* d_init);
// per-base top nodes of the emit probabilities (hint.rs:124-142) appended to `out` (host Mappings, ORIGINAL ids)
int run_products_mapping(dbgphmm_model* m, const std::vector<HJob>& jobs, const RowStore& F, const RowStore& B, int by_ratio,
                         uint32_t n_active, double ratio, dbgphmm_mappings* out);

// device bytes reserved for n_rows stored sparse rows (typical size ; the sparse phase is repeated with the upper bound if exceeded)
uint64_t arena_estimate(uint64_t n_rows, uint32_t n_active, bool ratio);
