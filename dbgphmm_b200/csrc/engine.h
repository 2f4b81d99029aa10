// engine.h — batch orchestration of the dense / sparse phases, row stores, products (internal).
#pragma once
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
    void release();
};

// Rows of one direction for a batch of jobs (PHMMTables of every job).
struct RowStore {
    int dir = 0;                     // 0 forward, 1 backward
    RowDesc* d_desc = nullptr;       // [sum len]
    uint64_t n_desc = 0;
    std::vector<uint64_t> desc0;     // per job
    std::vector<uint32_t> len;       // per job
    std::vector<uint32_t> nd;        // per job: number of dense rows (forward: rows [0,nd) ; backward: see engine.cu)
    std::vector<uint64_t> slab0;     // per job: first slab (forward row r -> slab0 + r ; backward row t -> slab0 + (hi - t)), kept rows only
    std::vector<int32_t> bdense_lo, bdense_hi;  // backward: dense rows are [lo, hi] (inclusive), -1/-1 if none
    DensePool pool;
    SparseArena arena;
    XF* d_final = nullptr;           // per job: forward e(last) / backward mb(first)
    std::vector<XF> h_final;
    uint64_t cells = 0;
    void release();
};

struct DevBuf {  // scoped device allocation
    void* p = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        if (p) { cudaFree(p); p = nullptr; }
        if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { p = nullptr; cudaGetLastError(); dbg_set_error("out of device memory"); return DBGPHMM_ERR_OOM; }
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
extern EngineTimes g_times;

int upload_mappings(dbgphmm_model* m, const dbgphmm_mappings* mp, DevMappings* out);
int run_forward(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, int kind, bool keep_rows, bool store_sparse,
                const DevMappings* dmap, RowStore* out);
int run_backward(dbgphmm_model* m, const std::vector<HJob>& jobs, const uint8_t* d_bases, int kind, bool keep_rows,
                 const DevMappings* dmap, const RowStore* fwd, RowStore* out);
// sum over rows of F (x) B / P into d_freqs (device, ORIGINAL node order, added to); status ZERO_PROB if some P == 0
int run_products_freqs(dbgphmm_model* m, const std::vector<HJob>& jobs, const RowStore& F, const RowStore& B, double* d_freqs);
// per-base top nodes of the emit probabilities (hint.rs:124-142) appended to `out` (host Mappings, ORIGINAL ids)
int run_products_mapping(dbgphmm_model* m, const std::vector<HJob>& jobs, const RowStore& F, const RowStore& B, int by_ratio,
                         uint32_t n_active, double ratio, dbgphmm_mappings* out);
