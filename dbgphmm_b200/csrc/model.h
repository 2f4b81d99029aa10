// model.h — host-side structures of the device graph (internal; the public surface is include/dbgphmm_b200.h)
#pragma once
#include <atomic>
#include <cstdint>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include "common.cuh"
#include "../../include/dbgphmm_b200.h"

// thread-local error string + status helpers
void dbg_set_error(const std::string& s);
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dbg_set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + " " + #expr + ": " + cudaGetErrorString(_e));                 \
            return DBGPHMM_ERR_CUDA;                                                           \
        }                                                                                      \
    } while (0)
#define ST_TRY(expr)                 \
    do {                             \
        int _s = (expr);             \
        if (_s != DBGPHMM_OK) return _s; \
    } while (0)

// Every status-returning entry point of the C ABI is a function-try-block closed by this: no C++ exception (host allocation
// failure inside a std::vector, ...) crosses the boundary into ctypes / Rust / C callers
#define ABI_CATCH                                                                                                             \
    catch (const std::bad_alloc&) { dbg_set_error("out of host memory"); return DBGPHMM_ERR_OOM; }                            \
    catch (const std::exception& e) { dbg_set_error(std::string("C++ exception: ") + e.what()); return DBGPHMM_ERR_INVALID; } \
    catch (...) { dbg_set_error("unknown C++ exception"); return DBGPHMM_ERR_INVALID; }

extern std::atomic<unsigned long long> g_launch_count;
#define COUNT_LAUNCH() (g_launch_count.fetch_add(1, std::memory_order_relaxed))

// Tiling plan of the dense DP for one direction.  A chunk is a run of consecutive (relabelled) nodes; its
// local set adds every node within HALO_HOPS hops upstream (forward: ancestors, backward: descendants), sorted
// by hop depth, so that one CTA can compute a whole DP row of the chunk from the previous row alone.
struct DevPlan {
    uint32_t n_chunks = 0;
    uint32_t max_local = 0;
    uint32_t* chunk_start = nullptr;  // [n_chunks+1] first relabelled node of each chunk
    uint32_t* loc_base = nullptr;     // [n_chunks+1] offset of the chunk's local node list
    uint32_t* loc_node = nullptr;     // [loc_total] relabelled node id of each local slot (core nodes first)
    uint32_t* nle = nullptr;          // [n_chunks][8] number of local slots with depth <= h (h = 0..6)
    uint32_t* le_off = nullptr;       // [loc_total + n_chunks] per-slot offset into le_idx/le_eid (chunk c uses base loc_base[c]+c)
    uint16_t* le_idx = nullptr;       // local slot of the upstream neighbour
    uint32_t* le_eid = nullptr;       // original EdgeIndex of that edge (indexes trans)
    // the same adjacency split for the common-frame kernel: first upstream neighbour inline, the rest ("extras") in a CSR
    uint16_t* fp_idx = nullptr;       // [loc_total] local slot of the first upstream neighbour (own slot if none)
    uint32_t* fp_eid = nullptr;       // [loc_total] its EdgeIndex, 0xffffffff if none
    uint32_t* fx_off = nullptr;       // [loc_total + n_chunks] offsets of the extra upstream edges (chunk c uses base loc_base[c]+c)
    uint16_t* fx_idx = nullptr;
    uint32_t* fx_eid = nullptr;
    // Register-stencil layout of the same tiles (k_dense_reg): DENSE_LMAX positions per tile ordered
    // [upstream chain of the tile head, deepest first | core nodes | remaining halo nodes in first-parent chains], so that
    // "first upstream neighbour == previous position" holds for almost every node.  Fixed stride DENSE_LMAX per tile.
    uint32_t* rl_node = nullptr;      // [n_chunks * LMAX] relabelled node id, 0xffffffff = padding
    uint16_t* rl_par = nullptr;       // [n_chunks * LMAX] position of the first upstream neighbour, 0xffff = not in the tile
    uint32_t* rl_eid = nullptr;       // [n_chunks * LMAX] its EdgeIndex (0xffffffff if none)
    uint8_t* rl_flag = nullptr;       // [n_chunks * LMAX] bit0: first neighbour is NOT the previous position, bit1: has extra edges
    uint16_t* rl_core = nullptr;      // [n_chunks * 2] position of the first core node, number of core nodes
    uint32_t* rx_off = nullptr;       // [n_chunks * (LMAX + 1)] offsets into rx_idx / rx_eid (absolute)
    uint16_t* rx_idx = nullptr;       // extra upstream neighbours as positions
    uint32_t* rx_eid = nullptr;
    std::vector<uint32_t> h_chunk_start;
};

extern thread_local int tl_stream_set;   // which StreamSet of a model this host thread drives (0 unless a bulk call says otherwise)
#define MSET(m) ((m)->ss[tl_stream_set])

struct dbgphmm_model {
    uint64_t serial = 0;            // unique per created model (keys the device copies other handles keep for it)
    int device = 0;
    uint32_t N = 0, E = 0;
    uint32_t n_batch = 1;
    dbgphmm_params params;  // logs, as given
    LinParams lin;          // linear
    uint64_t mem_budget = 0;        // fixed by the caller, or 88 % of the memory free at creation (fallback only: see model_budget)
    bool mem_budget_fixed = false;
    int n_sm = 148;
    // Two sets of streams / per-launch scratch: a bulk call may drive the forward and the backward direction from two host threads
    // (api.cu: the sparse phases of both directions then run side by side); every thread works on the set its tl_stream_set names.
    struct StreamSet {
        cudaStream_t stream = nullptr;
        cudaStream_t aux = nullptr;                // rescue launch of the sparse kernel (sparse.cu)
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
        void* d_jstep = nullptr; uint32_t jstep_cap = 0;   // per-(job, step) scalars of the dense fast kernel (dense.cu: JStep)
    };
    mutable StreamSet ss[2];

    // host copies (relabelled ids unless noted)
    std::vector<uint32_t> pos_of;   // [N] original id -> relabelled
    std::vector<uint32_t> orig_of;  // [N] relabelled -> original id
    std::vector<uint8_t> emission;  // [N] relabelled
    std::vector<uint32_t> par_off, par_node, par_eid;  // parents CSR, newest-edge-first (graph/iterators.rs:133-155)
    std::vector<uint32_t> chi_off, chi_node, chi_eid;  // children CSR, newest-edge-first (graph/iterators.rs:104-131)
    std::vector<uint32_t> e_src, e_dst;                // original edge list, ORIGINAL node ids

    // device graph
    uint32_t *d_pos_of = nullptr, *d_orig_of = nullptr;
    uint8_t* d_emission = nullptr;
    uint32_t *d_par_off = nullptr, *d_par_node = nullptr, *d_par_eid = nullptr;
    uint32_t *d_chi_off = nullptr, *d_chi_node = nullptr, *d_chi_eid = nullptr;
    // per node {first CSR slot, degree, first neighbour, its edge id}: one 16-byte load answers the common degree-1 case of the
    // latency-bound sparse kernel (CSR offset -> neighbour -> edge id is a chain of dependent L2 loads otherwise)
    uint4 *d_par_rec = nullptr, *d_chi_rec = nullptr;
    uint32_t max_deg = 0;       // largest in- or out-degree of a node
    double* d_init = nullptr;   // [n_batch][N] linear, relabelled node order
    double* d_trans = nullptr;  // [n_batch][E] linear, original EdgeIndex order
    DevPlan fwd, bwd;
    DevPlan fwd2, bwd2;   // tiles with a 2 x HALO_HOPS halo (two rows per launch) ; n_chunks == 0: not available
    // Recompute support for the stream strategy: for every forward tile the tiles that intersect the upstream closure of
    // its nodes within HALO_HOPS * n_warmup hops (the dependency cone of n_warmup dense rows), built lazily.
    uint32_t roi_warmup = 0;
    uint32_t *d_roi_off = nullptr, *d_roi_tile = nullptr, *d_tile_of = nullptr;
    uint32_t *d_roi_off_b = nullptr, *d_roi_tile_b = nullptr, *d_tile_of_b = nullptr;   // the same for backward tiles (downstream closure)
};
int model_ensure_roi(dbgphmm_model* m);

struct dbgphmm_reads {
    uint64_t n_reads = 0;
    std::vector<uint64_t> off;
    std::vector<uint8_t> bases;
    uint8_t* d_bases = nullptr;  // device copy (lazy)
    int device = -1;
};

struct dbgphmm_mappings {
    std::vector<uint64_t> read_off, row_off;
    std::vector<uint32_t> nodes;  // ORIGINAL node ids
    std::vector<double> logp;
    // Device copy (row offsets + node ids relabelled for ONE model), made by the first bulk call that uses the mappings with that
    // model and kept: sample_posterior_once scores hundreds of candidate sets against the same mappings (posterior.rs:504-515), and
    // relabelling + uploading 80 M node ids was a third of a C4 call.  The handle's content never changes after creation.
    mutable uint64_t* d_row_off = nullptr;
    mutable uint32_t* d_nodes = nullptr;
    mutable uint64_t d_model_serial = 0;
    mutable int d_device = -1;
};

int model_build_graph(dbgphmm_model* m, uint32_t n_nodes, uint32_t n_edges, const uint32_t* src, const uint32_t* dst,
                      const uint8_t* emission);
int model_upload_probs(dbgphmm_model* m, const double* log_init, const double* log_trans);
void model_free(dbgphmm_model* m);
LinParams to_lin(const dbgphmm_params& p);
