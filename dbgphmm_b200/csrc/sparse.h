// sparse.h — sparse DP rows (<= MAX_ACTIVE active nodes): one CTA per (read, candidate X), rows in shared memory.
#pragma once
#include "model.h"

enum {  // node-list source of a sparse step
    SP_TOPN = 0,     // top_nodes(n_active) of the previous row, adaptive Del sets   (forward.rs:112-150, backward.rs:166-176)
    SP_RATIO = 1,    // top_nodes_by_score_ratio(max_ratio), adaptive (forward only)  (forward.rs:112-150)
    SP_MAPPING = 2,  // mapping.nodes(i), non-adaptive                                (forward.rs:71, backward.rs:84)
    SP_BYFWD = 3     // filled_nodes() of forward row i-1, non-adaptive (backward)    (backward.rs:128-133)
};
enum { SPREV_F_INIT = 0, SPREV_B_INIT = 1, SPREV_DENSE = 2,
       SPREV_GATHER = 3 };   // the cells of the last dense row that step 0 can read, gathered into a per-job list (sparse_gather_prev0)
enum { SJ_OK = 0, SJ_NEED_BIG = 1, SJ_CAPACITY = 2, SJ_OOM = 3 };

struct SJob {
    uint32_t x;           // parameter set
    uint32_t len;         // read length
    uint64_t base_off;    // first base of the read in the device base array
    int dir;              // 0 forward, 1 backward
    int mode;             // SP_*
    int32_t row_begin;    // row of step 0 ; row(s) = row_begin + s (forward) / row_begin - s (backward)
    uint32_t n_rows;      // rows to compute
    int prev0_kind;       // SPREV_* : what the row before step 0 is
    uint64_t prev0_slab;  // dense slab of that row (SPREV_DENSE)
    uint32_t top0;        // request slot of the dense top list for step 0 (SP_TOPN / SP_RATIO)
    uint64_t desc0;       // RowDesc index of row 0 of this job (own direction)
    uint64_t fdesc0;      // SP_BYFWD: RowDesc index of row 0 of the forward tables
    uint64_t map_row0;    // SP_MAPPING: first row of this read in the mapping CSR
    int store;            // 1: write rows to the arena, 0: score only
    int32_t active_idx;   // skip unless active flag (or -1)
    uint32_t out_idx;     // job index in the caller's batch
};

// State a job hands over when a row outgrows its launch's entry capacity: the job carries on from `step` in a CTA of the
// concurrently running rescue launch (larger capacity).  The previous row itself goes to rq_rows.
struct alignas(16) SHandoff {
    uint32_t step, n_prev;
    unsigned long long cells;
    XF mb, ib, last_scalar;
};

struct SparseArena {
    char* base = nullptr;
    uint64_t bytes = 0;
    unsigned long long* cursor = nullptr;  // device bump pointer
};
#define SPARSE_PAGE_BYTES (256 * 1024)

struct SparseIO {
    const uint8_t* bases;
    RowDesc* desc;             // own direction
    const RowDesc* fdesc;      // forward tables (SP_BYFWD)
    const char* farena;        // arena holding the forward sparse rows (SP_BYFWD)
    const uint32_t* top_ids;   // dense_select output [slot][MAX_ACTIVE]
    const uint32_t* top_cnt;
    const uint64_t* map_row_off;  // mapping CSR (device): row -> entries
    const uint32_t* map_nodes;    // relabelled node ids
    const char* pool; uint64_t slab_bytes; uint32_t Np;  // dense pool (SPREV_DENSE)
    const char* gather; uint32_t gather_cap; const uint32_t* gather_cnt;   // SPREV_GATHER: per request slot m, i, d [cap] f64, id, ex [cap] 32-bit ; entries in use
    char* arena; uint64_t arena_bytes; unsigned long long* arena_cursor;
    int* status;               // per job SJ_*
    XF* final_scalar;          // per job: forward -> e of the last row ; backward -> mb of row 0
    unsigned long long* cells; // per job: sum over rows of |nodes| (GCUPS numerator)
    const int* active;
    // rescue queue (nullptr: a job that needs more capacity reports SJ_NEED_BIG and is re-run by the caller)
    uint32_t* rq_ctl;          // [0] jobs pushed, [1] jobs taken, [2] primary CTAs that have left, [3] next job of the primary launch
    uint32_t* rq_items;        // [n_jobs] pushed job indices, 0xffffffff until published
    SHandoff* rq_hand;         // [n_jobs]
    char* rq_rows;             // [n_jobs][32 * rq_cap] previous row: m, i, d, id, ex
    uint32_t rq_cap;           // entry capacity of the primary launch
    uint32_t rq_n_primary;     // CTAs of the primary launch
    uint32_t rq_n_jobs;        // jobs of the primary launch
};

int sparse_configure(dbgphmm_model* m);
// entry capacity of the first pass over top-n / ratio / by-forward jobs (DBGPHMM_SPARSE_CAP overrides it)
uint32_t sparse_default_cap();
// jobs that are resident at once with entry capacity `cap` (one wave): batches are cut to multiples of it
uint32_t sparse_wave_jobs(dbgphmm_model* m, uint32_t cap);
// n_jobs forward jobs and n_jobs backward jobs (primary + rescue launches of both directions) are resident at once
bool sparse_pair_fits(dbgphmm_model* m, uint32_t cap, uint32_t n_jobs);
// cap: entry capacity per job in shared memory; threads per CTA chosen from cap.  rescue_cap > cap: a second launch of
// persistent CTAs with that capacity runs beside the primary one (auxiliary stream) and carries on the jobs whose rows
// outgrow `cap` (io.rq_* must be set up by the caller); 0: no rescue launch.
// dir: SJob::dir of every job of the launch (0 forward, 1 backward)
int sparse_run(dbgphmm_model* m, const SJob* d_jobs, uint32_t n_jobs, const SparseIO& io, uint32_t cap, int dir, uint32_t rescue_cap = 0);
// Gather, for the request slots j = slot0 .. slot0 + n - 1 with slabs[j - slot0] != ~0, the cells of dense row slabs[j - slot0] that
// step 0 of a top-n job started from top_ids[j] can read (forward: top, children(top) and the parents of both ; backward: top,
// parents(top) and the children of both) into out + j * 32 * cap, their number into out_cnt[j] (other slots are left alone).  cap must
// be sparse_gather_cap(m, k) for lists of at most k ids.  *d_overflow is raised if a list outgrows cap (cannot happen for that cap).
// With these lists the dense slabs are not needed during the sparse phase.
uint32_t sparse_gather_cap(const dbgphmm_model* m, uint32_t k);
int sparse_gather_prev0(dbgphmm_model* m, int dir, uint32_t slot0, uint32_t n, const uint32_t* d_top_ids, const uint32_t* d_top_cnt, const uint64_t* d_slabs,
                        const char* pool, uint64_t slab_bytes, uint32_t Np, uint32_t cap, char* d_out, uint32_t* d_out_cnt, int* d_overflow);
// capacity of the rescue launch that accompanies a primary launch of capacity `cap` (0: none)
uint32_t sparse_rescue_cap(uint32_t cap);
