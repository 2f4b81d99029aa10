// dense.h — dense DP rows (all N nodes): tiled forward/backward step kernels, row reductions, top-k selection.
#pragma once
#include "model.h"

#define DENSE_CORE 152      // most nodes a tile may own (the planner shrinks it until tile + halo + alignment pads fit DENSE_LMAX)
#define DENSE_LMAX 160      // tile + 6-hop halo capacity
#define DENSE_PER_LANE (DENSE_LMAX / 32)   // register-stencil kernel: consecutive positions owned by one lane
#define DENSE_EMAX 256      // local upstream edges of one tile
#define DENSE_XMAX 64       // of which beyond the first of each node ("extras")
#define DENSE_THREADS 64    // exact (per-value exponent) kernel: one CTA per tile
#define DENSE_SLOTS ((DENSE_LMAX + DENSE_THREADS - 1) / DENSE_THREADS)
#ifndef WT_WARPS
#define WT_WARPS 8          // common-frame kernel: warps (= tiles) per CTA
#endif
#define WT_SLOTS (DENSE_LMAX / 32)
#ifndef WT_MIN_CTAS
#define WT_MIN_CTAS 2
#endif
#define SELECT_THREADS 1024
#define SELECT_CAP 4096     // candidates resolved in shared memory by the top-k selection

enum { PREV_F_INIT = 0, PREV_B_INIT = 1, PREV_SLAB = 2 };

// Static description of one job's dense phase (device array, one per job).
struct DJob {
    uint32_t x;          // parameter set (candidate X)
    uint32_t len;        // read length
    uint64_t base_off;   // first base of the read in the device base array
    uint32_t n_steps;    // number of dense steps of this phase
    int32_t first_row;   // row of step 0; row(s) = first_row + s (forward) or first_row - s (backward)
    int prev0_kind;      // PREV_* for step 0
    uint64_t prev0_slab; // slab holding the row before step 0 when prev0_kind == PREV_SLAB
    uint64_t slab0;      // slab of step 0
    uint32_t slab_mod;   // 0: step s -> slab0 + s (rows kept) ; k > 0: slab0 + (s % k) (ping-pong)
    uint64_t desc0;      // RowDesc index of row 0 of this job
    int32_t active_idx;  // index into the `active` flag array, or -1
};

struct DensePool {
    char* base = nullptr;     // n_slabs slabs
    uint64_t slab_bytes = 0;  // per slab: double m[Np], i[Np], d[Np]; int ex[Np]
    uint64_t n_slabs = 0;
    uint32_t Np = 0;          // N rounded up to a multiple of 2
};
static inline uint64_t dense_slab_bytes(uint32_t N) {
    uint64_t Np = (N + 1) & ~1u;
    return ((Np * 28 + 255) / 256) * 256;
}

int dense_configure(dbgphmm_model* m);
// one forward / backward step `s` for all jobs (grid = chunks x jobs) followed by the row reduction.
int dense_forward_step(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s,
                       const uint8_t* d_bases, RowDesc* d_desc, const int* d_active, XF* d_partials, unsigned long long* d_worklist,
                       uint64_t step_cells);
int dense_backward_step(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s,
                        const uint8_t* d_bases, RowDesc* d_desc, const int* d_active, XF* d_partials, unsigned long long* d_worklist,
                       uint64_t step_cells);

// Two forward rows (s, s + 1) per launch for ping-pong jobs whose intermediate row nobody reads (dense.cu: k_dense_fwd2).
// Pair p = s / 2 reads slab0 + ((p - 1) & 1) and writes slab0 + (p & 1).  *d_redo is raised if some tile's exponent range does
// not fit the two-row frame: the caller then repeats the phase with single-row steps.  second: some job has a row s + 1.
bool dense_can_pair(const dbgphmm_model* m);
uint32_t dense_pair_tiles(const dbgphmm_model* m, int dir);
int dense_forward_pair(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s, const uint8_t* d_bases,
                       RowDesc* d_desc, XF* d_partials, int* d_redo, uint64_t pair_cells, bool second);
int dense_backward_pair(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t n_jobs, uint32_t s, const uint8_t* d_bases,
                        RowDesc* d_desc, XF* d_partials, int* d_redo, uint64_t pair_cells, bool second);
int dense_backward_step_list(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t s, const uint8_t* d_bases, XF* d_partials,
                             const unsigned long long* d_worklist);

// forward step restricted to the (job, tile) pairs of a prebuilt worklist (recompute pass of the stream strategy);
// no row reduction: the rows' scalars are taken from the descriptors written by the first pass.
int dense_forward_step_list(dbgphmm_model* m, const DensePool& pool, const DJob* d_jobs, uint32_t s, const uint8_t* d_bases,
                            const RowDesc* d_desc, XF* d_partials, const unsigned long long* d_worklist);

// Top-k of merged (m+i+d) values of dense rows (PHMMTable::top_nodes / top_nodes_by_score_ratio on a dense
// table, table.rs:127-149).  One CTA per request.
struct SelectReq {
    uint64_t slab;       // row to select from
    uint32_t k;          // number of ids wanted (<= MAX_ACTIVE)
    int by_ratio;        // keep only ids with ln(v0) - ln(v) < ratio
    double ratio;
    int32_t active_idx;  // skip unless active[active_idx] (or -1)
    uint32_t out;        // request slot: ids at out_ids[out*MAX_ACTIVE ...], count at out_cnt[out]
};
int dense_select(dbgphmm_model* m, const DensePool& pool, const SelectReq* d_reqs, uint32_t n_reqs, const int* d_active,
                 uint32_t* d_out_ids, uint32_t* d_out_cnt);
