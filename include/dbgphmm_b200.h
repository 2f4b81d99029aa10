/* dbgphmm_b200.h — C ABI of the B200-native hmmv2 PHMM path.
 *
 * Drop-in boundary for ryought/dbgphmm `src/hmmv2`: the reference has no FFI layer, the seam is the
 * inherent-method set `impl PHMMModel<N,E>` + `impl PHMMOutput`.  Each entry point below
 * names the reference method it replaces (paths relative to the reference's src/).  A Rust shim that
 * re-exposes these under the original method names lives in rust/ (source only).
 *
 * Conventions
 *   - every call returns an int status (0 = DBGPHMM_OK); the reference panics instead (no Result).  No C++ exception
 *     crosses the boundary: allocation failures and the like come back as a status as well.
 *     dbgphmm_last_error() returns a thread-local message for the last non-zero status.
 *   - probabilities cross the ABI as natural-log f64, exactly `Prob.0` (prob.rs:13,74-76);
 *     zero probability is -INFINITY.
 *   - node ids are the caller's petgraph NodeIndex values (u32); edges are given in EdgeIndex order,
 *     which fixes the parent/child iteration order (graph/iterators.rs:104-155).
 *   - host pointers in, host pointers out, unless the name ends in _dev.
 *   - handles are immutable after creation except where stated; a handle may be used from one host
 *     thread at a time, different handles from different threads (the device-memory cache shared by the
 *     handles of a process is locked; dbgphmm_last_timing / dbgphmm_last_dense_kernel report the calling
 *     thread's last bulk call).  All GPU work of a call is finished when it returns.  A bulk call may start one helper
 *     host thread of its own (it drives the backward direction on the handle's second stream set) and joins it before returning.
 *   - there is NO CPU fallback: every call fails with DBGPHMM_ERR_CUDA if no sm_100-class device exists.
 */
#ifndef DBGPHMM_B200_H
#define DBGPHMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBGPHMM_OK 0
#define DBGPHMM_ERR_INVALID 1   /* bad argument (the reference would panic on an assert / index) */
#define DBGPHMM_ERR_CUDA 2      /* CUDA runtime failure or no device */
#define DBGPHMM_ERR_CAPACITY 3  /* a sparse row needed more than MAX_ACTIVE_NODES entries (table.rs:22; the
                                   reference's ArrayVec-backed SparseVec panics with "insufficient capacity") */
#define DBGPHMM_ERR_ZERO_PROB 4 /* P(read) == 0 where the reference would produce NaN and panic in Ord
                                   (table.rs:500-505 with prob.rs:296-300) */
#define DBGPHMM_ERR_OOM 5       /* device memory budget too small for one read, or a host allocation failed */

#define DBGPHMM_MAX_ACTIVE_NODES 400 /* hmmv2/table.rs:22 */

/* PHMMParams (hmmv2/params.rs:16-66).  All p_* are natural logs. */
typedef struct dbgphmm_params {
    double p_mismatch, p_match, p_random, p_gap_open, p_gap_ext, p_end;
    double p_MM, p_IM, p_DM, p_MI, p_II, p_DI, p_MD, p_ID, p_DD;
    uint32_t n_active_nodes;   /* params.rs:28-38 */
    uint32_t n_warmup;         /* params.rs:43-50 ; MultiDbg::to_phmm overrides it with k (multi_dbg.rs:1395) */
    uint32_t warmup_threshold; /* params.rs:51-58 */
    uint32_t n_max_gaps;       /* params.rs:59-62 ; must be 4 (table.rs:17 MAX_DEL) */
    double active_node_max_ratio; /* params.rs:39-42 */
} dbgphmm_params;

typedef struct dbgphmm_model dbgphmm_model;       /* PHMMModel<PNode,PEdge>   hmmv2/common.rs:61-67 */
typedef struct dbgphmm_reads dbgphmm_reads;       /* ReadCollection<S>         common/collection.rs:131 */
typedef struct dbgphmm_mappings dbgphmm_mappings; /* Mappings                  hmmv2/hint.rs:150-152 */
typedef struct dbgphmm_tables dbgphmm_tables;     /* PHMMTables of one read    hmmv2/table.rs:365-435 */

const char* dbgphmm_last_error(void);
/* number of usable sm_100-class devices (0 => every compute call fails; no CPU path exists) */
int dbgphmm_device_count(void);

/* PHMMParams::new / uniform (params.rs:73-124) */
void dbgphmm_params_new(double p_mismatch, double p_gap_open, double p_gap_ext, double p_end,
                        uint32_t n_active_nodes, uint32_t n_warmup, dbgphmm_params* out);
void dbgphmm_params_uniform(double p, dbgphmm_params* out);

/* ---- model ------------------------------------------------------------------------------------------- */
/* Build the device graph from a node-centric PHMM (common.rs:61-67,202-261).
 * edge_src/edge_dst: EdgeIndex order.  emission: 'A','C','G','T' or 'n' (NULL_BASE, common.rs:21).
 * log_init[n_nodes], log_trans[n_edges]: PNode.init_prob / PEdge.trans_prob as natural logs.
 * device: CUDA ordinal.  mem_budget_bytes: device memory the handle may use for DP rows
 * (0 = 88% of what is free at the time of each bulk call, counting this library's reusable cache blocks). */
int dbgphmm_model_create(uint32_t n_nodes, uint32_t n_edges, const uint32_t* edge_src, const uint32_t* edge_dst,
                         const uint8_t* emission, const double* log_init, const double* log_trans,
                         const dbgphmm_params* params, int device, uint64_t mem_budget_bytes, dbgphmm_model** out);
void dbgphmm_model_destroy(dbgphmm_model* m);
int dbgphmm_model_set_params(dbgphmm_model* m, const dbgphmm_params* params);
/* Replace init/trans of parameter set 0 (a rebuilt PModel with the same graph). */
int dbgphmm_model_set_probs(dbgphmm_model* m, const double* log_init, const double* log_trans);
/* Derive n_batch parameter sets on the device from node copy numbers: SeqGraph::to_phmm (mode 0),
 * to_non_zero_phmm (mode 1), to_uniform_phmm (mode 2)  — graph/seq_graph.rs:160-273 with edge copy numbers
 * None (the branch MultiDbg uses, multi_dbg.rs:1386-1387).  copy_nums is [n_batch][n_nodes].  After this call
 * dbgphmm_to_full_prob_reads evaluates all n_batch candidates X in one launch (posterior.rs:504-515). */
int dbgphmm_model_set_copy_nums_batch(dbgphmm_model* m, uint32_t n_batch, const uint32_t* copy_nums, int mode);
/* Read back parameter set `x` (natural logs), for tests. */
int dbgphmm_model_get_probs(const dbgphmm_model* m, uint32_t x, double* log_init, double* log_trans);
uint32_t dbgphmm_model_n_nodes(const dbgphmm_model* m);
uint32_t dbgphmm_model_n_batch(const dbgphmm_model* m);

/* ---- reads / mappings --------------------------------------------------------------------------------- */
/* offsets[n_reads+1] into bases; bases are uppercase ACGT (collection.rs:236-249 panics otherwise). */
int dbgphmm_reads_create(uint64_t n_reads, const uint64_t* offsets, const uint8_t* bases, dbgphmm_reads** out);
void dbgphmm_reads_destroy(dbgphmm_reads* r);
/* Mappings as CSR: read_off[n_reads+1] (rows = bases), row_off[n_rows+1] (entries), nodes[], logp[]
 * (hint.rs:27-30: Mapping.nodes / Mapping.probs). */
int dbgphmm_mappings_create(uint64_t n_reads, const uint64_t* read_off, const uint64_t* row_off,
                            const uint32_t* nodes, const double* logp, dbgphmm_mappings** out);
void dbgphmm_mappings_destroy(dbgphmm_mappings* mp);
int dbgphmm_mappings_sizes(const dbgphmm_mappings* mp, uint64_t* n_reads, uint64_t* n_rows, uint64_t* n_entries);
int dbgphmm_mappings_export(const dbgphmm_mappings* mp, uint64_t* read_off, uint64_t* row_off, uint32_t* nodes, double* logp);
/* Mappings::to_node_freqs (hint.rs:161-171) == MultiDbg::mappings_to_freqs (multi_dbg/draft.rs:201-212) */
int dbgphmm_mappings_to_node_freqs(const dbgphmm_mappings* mp, uint32_t n_nodes, double* freqs);

/* Mapping::map_nodes (hint.rs:66-88) for every read of the handle: node v of the old graph corresponds to
 * map_to[map_off[v] .. map_off[v+1]) in the new one (possibly none).  Each (v, p) of a base spreads p / |images| over its images,
 * coinciding images are added (Prob +, prob.rs:181-197), the base's list is re-sorted by probability, descending, and cut to
 * DBGPHMM_MAX_ACTIVE_NODES.  This is the step that carries a hint over a graph change: MultiDbg::hint_kp1_from_hint_k
 * (multi_dbg.rs:1325-1334: v -> the k+1 edges entering node v) and PurgeEdgeMap::update_mapping (multi_dbg.rs:1783-1791: v -> the
 * surviving edge, or nothing).  Order among equal probabilities: first appearance (the reference's HashMap order is unspecified). */
int dbgphmm_mappings_map_nodes(const dbgphmm_mappings* mp, uint32_t n_nodes_before, const uint64_t* map_off, const uint32_t* map_to,
                               dbgphmm_mappings** out);

/* ---- PHMMTables of one read (forward.rs / backward.rs drivers) ---------------------------------------- */
enum {
    DBGPHMM_FWD_DENSE = 0,        /* forward               forward.rs:24  */
    DBGPHMM_FWD_SPARSE = 1,       /* forward_sparse(false) forward.rs:93  */
    DBGPHMM_FWD_SPARSE_RATIO = 2, /* forward_sparse(true)  forward.rs:93  */
    DBGPHMM_FWD_MAPPING = 3       /* forward_with_mapping  forward.rs:51  */
};
enum {
    DBGPHMM_BWD_DENSE = 0,     /* backward              backward.rs:24  */
    DBGPHMM_BWD_SPARSE = 1,    /* backward_sparse       backward.rs:146 */
    DBGPHMM_BWD_MAPPING = 2,   /* backward_with_mapping backward.rs:59  */
    DBGPHMM_BWD_BY_FORWARD = 3 /* backward_by_forward   backward.rs:101 */
};
/* mapping / read_index are used by the *_MAPPING kinds only; fwd is used by BWD_BY_FORWARD only. */
int dbgphmm_forward(dbgphmm_model* m, const uint8_t* bases, uint64_t n, int kind,
                    const dbgphmm_mappings* mapping, uint64_t read_index, dbgphmm_tables** out);
int dbgphmm_backward(dbgphmm_model* m, const uint8_t* bases, uint64_t n, int kind,
                     const dbgphmm_mappings* mapping, uint64_t read_index, const dbgphmm_tables* fwd,
                     dbgphmm_tables** out);
void dbgphmm_tables_destroy(dbgphmm_tables* t);
uint64_t dbgphmm_tables_len(const dbgphmm_tables* t);
/* PHMMTables::full_prob (table.rs:395-401): last e (forward) or first mb (backward) */
int dbgphmm_tables_full_prob(const dbgphmm_tables* t, double* logp);
/* row = -1 is init_table.  info = {is_dense, n_mi, n_d}; scalars = {mb, ib, e} (natural logs). */
int dbgphmm_tables_row_info(const dbgphmm_tables* t, int64_t row, uint64_t info[3], double scalars[3]);
/* dense row: m,i,d each n_nodes (ids unused).  sparse row: ids_mi[n_mi], m[n_mi], i[n_mi], ids_d[n_d], d[n_d]
 * in the insertion order of the reference's SparseVec. */
int dbgphmm_tables_row_export(const dbgphmm_tables* t, int64_t row, uint32_t* ids_mi, double* m, double* i,
                              uint32_t* ids_d, double* d);
/* PHMMTable::top_nodes(k) (by_ratio = 0, table.rs:127) / top_nodes_by_score_ratio(ratio) (table.rs:134);
 * out has room for DBGPHMM_MAX_ACTIVE_NODES ids. */
int dbgphmm_tables_row_top_nodes(const dbgphmm_tables* t, int64_t row, int by_ratio, uint32_t k, double ratio,
                                 uint32_t* out, uint32_t* n_out);

/* ---- PHMMOutput of one read (table.rs:450-517, freq.rs:198-255, hint.rs:120-142) ---------------------- */
/* to_node_freqs (freq.rs:245-255): freqs[n_nodes] */
int dbgphmm_output_node_freqs(dbgphmm_model* m, const dbgphmm_tables* fwd, const dbgphmm_tables* bwd, double* freqs);
/* to_edge_and_init_freqs (freq.rs:276-298, over to_trans_and_init_probs :332-389): edge_freqs[n_edges] in EdgeIndex order,
 * init_freqs[n_nodes] (the Begin -> node transitions).  to_edge_freqs (freq.rs:302-309) is the first output. */
int dbgphmm_output_edge_and_init_freqs(dbgphmm_model* m, const dbgphmm_tables* fwd, const dbgphmm_tables* bwd, double* edge_freqs,
                                       double* init_freqs);
/* q_score_exact (q.rs:66-96) for candidate x: out = {init, trans, prior (always 0)}; QScore::total = their sum */
int dbgphmm_q_score_exact(const dbgphmm_model* m, uint32_t x, const double* edge_freqs, const double* init_freqs, double out[3]);
/* to_mapping(n_active) (by_ratio = 0) / to_mapping_by_score_ratio(ratio): a 1-read Mappings handle */
int dbgphmm_output_mapping(dbgphmm_model* m, const dbgphmm_tables* fwd, const dbgphmm_tables* bwd, int by_ratio,
                           uint32_t n_active, double ratio, dbgphmm_mappings** out);

/* ---- bulk calls over a read set (the rayon-parallel entry points) ------------------------------------- */
/* PHMMModel::to_full_prob_reads (freq.rs:175-192) evaluated for every parameter set X of the model:
 * mappings != NULL -> forward_with_mapping_score_only (forward.rs:79), else forward_sparse_score_only
 * (forward.rs:158) with use_max_ratio.  out_logp[n_batch] = sum over reads in read order (fixed order, unlike
 * rayon's); out_logp_per_read (nullable) is [n_batch][n_reads].  This is MultiDbg::to_likelihood
 * (multi_dbg/posterior.rs:247-255) for each candidate of sample_posterior_once (:504-515). */
int dbgphmm_to_full_prob_reads(dbgphmm_model* m, const dbgphmm_reads* reads, const dbgphmm_mappings* mappings,
                               int use_max_ratio, double* out_logp, double* out_logp_per_read);

enum {
    DBGPHMM_RUN_DENSE = 0,           /* run                  freq.rs:42 */
    DBGPHMM_RUN_SPARSE = 1,          /* run_sparse           freq.rs:51 */
    DBGPHMM_RUN_SPARSE_ADAPTIVE = 2, /* run_sparse_adaptive  freq.rs:60 */
    DBGPHMM_RUN_WITH_MAPPING = 3     /* run_with_mapping     freq.rs:72 */
};
/* For every read: run*() then PHMMOutput::to_node_freqs, summed over reads (PHMMModel::to_node_freqs,
 * freq.rs:87-102, generalised to the four run modes).  node_freqs[n_nodes] (nullable), logp_fwd[n_reads] =
 * to_full_prob_forward, logp_bwd[n_reads] = to_full_prob_backward (nullable).  cells (nullable) receives the
 * number of (base, node) cells evaluated by f_step/b_step {forward, backward} — the GCUPS numerator. */
int dbgphmm_run_node_freqs(dbgphmm_model* m, const dbgphmm_reads* reads, int mode, int use_max_ratio,
                           const dbgphmm_mappings* mappings, double* node_freqs, double* logp_fwd, double* logp_bwd,
                           uint64_t cells[2]);
/* Same, but accumulating into / writing device buffers (f64 node_freqs_dev[n_nodes] is ADDED to, so that a
 * caller can all-reduce it with NCCL without a host round trip; logp_*_dev are [n_reads], nullable). */
int dbgphmm_run_node_freqs_dev(dbgphmm_model* m, const dbgphmm_reads* reads, int mode, int use_max_ratio,
                               const dbgphmm_mappings* mappings, double* node_freqs_dev, double* logp_fwd_dev,
                               double* logp_bwd_dev, uint64_t cells[2]);
/* PHMMModel::generate_mappings (hint.rs:193-220): run_with_mapping if mappings != NULL else
 * run_sparse_adaptive(use_max_ratio); then to_mapping_by_score_ratio(active_node_max_ratio) if use_max_ratio
 * else to_mapping(n_active_nodes). */
int dbgphmm_generate_mappings(dbgphmm_model* m, const dbgphmm_reads* reads, const dbgphmm_mappings* mappings,
                              int use_max_ratio, dbgphmm_mappings** out);

/* ---- file formats either side of the path (host only, no GPU needed) ------------------------------------ */
typedef struct dbgphmm_dbg dbgphmm_dbg;           /* MultiDbg as a DBG file describes it  multi_dbg.rs:170-186 */
/* DBG text (K / N / E lines): MultiDbg::from_dbg_reader / from_dbg_file (multi_dbg/output.rs:203-345); a path ending in
 * .gz / .dbz is gzip (output.rs:135-139).  Errors where the reference asserts (unsorted ids, seq / edge-list length). */
int dbgphmm_dbg_from_text(const char* text, uint64_t len, dbgphmm_dbg** out);
int dbgphmm_dbg_from_file(const char* path, dbgphmm_dbg** out);
void dbgphmm_dbg_destroy(dbgphmm_dbg* d);
/* sizes = {k, n_nodes_full, n_edges_full (= PHMM nodes), n_nodes_compact, n_edges_compact, n PHMM edges} */
int dbgphmm_dbg_sizes(const dbgphmm_dbg* d, uint32_t sizes[6]);
/* MultiDbg::to_seq_graph -> to_node_centric_graph(add_terminal = false) (multi_dbg.rs:1370-1390,1551-1604): PHMM node id =
 * full-graph edge id; edge_src/edge_dst[n PHMM edges] in the reference's insertion order; emission, copy_nums and
 * compact_edge_of are [n_edges_full].  Any output may be NULL. */
int dbgphmm_dbg_phmm_graph(const dbgphmm_dbg* d, uint32_t* edge_src, uint32_t* edge_dst, uint8_t* emission, uint32_t* copy_nums,
                           uint32_t* compact_edge_of);
/* MultiDbg::get_copy_nums / set_copy_nums (multi_dbg.rs:1041-1066): one copy number per COMPACT edge; set fails like the
 * reference's assert when the numbers do not balance at every node (is_copy_nums_valid, multi_dbg.rs:1008-1014). */
int dbgphmm_dbg_get_copy_nums(const dbgphmm_dbg* d, uint32_t* compact_copy_nums);
int dbgphmm_dbg_set_copy_nums(dbgphmm_dbg* d, const uint32_t* compact_copy_nums);
/* Candidate copy-number vectors over compact edges [n_batch][n_edges_compact] (what sample_posterior_once proposes,
 * multi_dbg/posterior.rs:470-515) -> per-k-mer copy numbers [n_batch][n_edges_full] for dbgphmm_model_set_copy_nums_batch. */
int dbgphmm_dbg_expand_copy_nums(const dbgphmm_dbg* d, uint32_t n_batch, const uint32_t* compact, uint32_t* full);
/* The terms of MultiDbg::to_score beside the likelihood (multi_dbg/posterior.rs:225-277), for the current copy numbers (compact ==
 * NULL, n_batch = 1) or a batch of candidates [n_batch][n_edges_compact]; with dbgphmm_to_full_prob_reads they give
 * Score::p() = ln P(R|X) + ln P(G) + ln #circuits for every candidate of sample_posterior_once without rebuilding a MultiDbg.
 *   genome_size      MultiDbg::genome_size (multi_dbg.rs:1018-1028)
 *   n_euler_circuits MultiDbg::n_euler_circuits (multi_dbg.rs:831-837; BEST theorem, graph/euler.rs:22-123), natural log, -inf = none;
 *                    candidates must balance at every node (DBGPHMM_ERR_INVALID otherwise, the assert of set_copy_nums) */
int dbgphmm_dbg_genome_size(const dbgphmm_dbg* d, uint32_t n_batch, const uint32_t* compact, uint64_t* out);
int dbgphmm_dbg_n_euler_circuits(const dbgphmm_dbg* d, uint32_t n_batch, const uint32_t* compact, double* out);
/* euler_circuit_count (graph/euler.rs:94-123) of a multigraph given as an edge list with multiplicities (parallel edges and self
 * loops allowed): ln of the number of Euler circuits; allow_multiple_component = product over the strongly connected components.
 * A multigraph that does not balance at some node has no Euler circuit: -inf. */
int dbgphmm_euler_circuit_count(uint32_t n_nodes, uint64_t n_edges, const uint32_t* edge_src, const uint32_t* edge_dst,
                                const uint32_t* multiplicity, int allow_multiple_component, double* out_ln_count);
/* MultiDbg::to_prior (posterior.rs:225-231) = distribution::normal(x, mu, sigma) (distribution.rs:22-25), natural log */
int dbgphmm_prior_normal(double x, double mu, double sigma, double* out_ln_p);
/* MultiDbg::to_dbg_writer / to_dbg_file (output.rs:140-199).  Text calls: buf may be NULL to query *needed. */
int dbgphmm_dbg_to_text(const dbgphmm_dbg* d, char* buf, uint64_t cap, uint64_t* needed);
int dbgphmm_dbg_to_file(const dbgphmm_dbg* d, const char* path);
/* MultiDbg::to_phmm (mode 0) / to_non_zero_phmm (1) / to_uniform_phmm (2) (multi_dbg.rs:1391-1409): n_warmup := k. */
int dbgphmm_dbg_to_model(const dbgphmm_dbg* d, const dbgphmm_params* params, int mode, int device, uint64_t mem_budget_bytes,
                         dbgphmm_model** out);
/* MAP text (read, pos, base, node:lnP,...): MultiDbg::from_map_reader_raw / to_map_writer (output.rs:455-623); .gz / .mpz
 * is gzip.  Log probabilities are printed like Rust's `{}` (shortest round-trip digits, no exponent, "-inf"). d (nullable)
 * only feeds the header comment. */
int dbgphmm_mappings_from_map_text(const char* text, uint64_t len, dbgphmm_mappings** out);
int dbgphmm_mappings_from_map_file(const char* path, dbgphmm_mappings** out);
int dbgphmm_mappings_to_map_text(const dbgphmm_mappings* mp, const dbgphmm_reads* reads, const dbgphmm_dbg* d, char* buf, uint64_t cap,
                                 uint64_t* needed);
int dbgphmm_mappings_to_map_file(const dbgphmm_mappings* mp, const dbgphmm_reads* reads, const dbgphmm_dbg* d, const char* path);

/* Dataset JSON (Dataset::to_json_file / from_json_file, e2e.rs:31-52,123-130: serde_json of genome, genome_size, positioned reads and the
 * PHMM parameters they were sampled with).  Styled sequences are "C|L|F:bases" (collection.rs:371-391,460-464), positioned reads
 * "bases:+|-:origins" with an origin "hap-pos" or "I" (collection.rs:711-757, genome_graph.rs:117-151), probabilities "{ln p}({p:.4})"
 * (prob.rs:158-169).  A path ending in .gz is gzip.  dataset_reads hands the read set to the hot path. */
typedef struct dbgphmm_dataset dbgphmm_dataset;
int dbgphmm_dataset_from_json_text(const char* text, uint64_t len, dbgphmm_dataset** out);
int dbgphmm_dataset_from_json_file(const char* path, dbgphmm_dataset** out);
/* hap_style[h] in {'C' circular, 'L' linear, 'F' linear fragment}; read_revcomp, origin_hap (-1 = inserted base), origin_pos may be NULL */
int dbgphmm_dataset_create(uint32_t n_haps, const uint64_t* hap_off, const uint8_t* hap_bases, const uint8_t* hap_style, uint64_t genome_size,
                           uint64_t n_reads, const uint64_t* read_off, const uint8_t* read_bases, const uint8_t* read_revcomp, const int64_t* origin_hap,
                           const uint64_t* origin_pos, const dbgphmm_params* params, dbgphmm_dataset** out);
void dbgphmm_dataset_destroy(dbgphmm_dataset* d);
/* sizes = {haplotypes, genome bases, reads, read bases, genome_size field} */
int dbgphmm_dataset_sizes(const dbgphmm_dataset* d, uint64_t sizes[5]);
int dbgphmm_dataset_genome(const dbgphmm_dataset* d, uint64_t* hap_off, uint8_t* bases, uint8_t* style);
int dbgphmm_dataset_reads(const dbgphmm_dataset* d, dbgphmm_reads** out);
int dbgphmm_dataset_read_origins(const dbgphmm_dataset* d, uint64_t* read_off, uint8_t* bases, uint8_t* revcomp, int64_t* origin_hap, uint64_t* origin_pos);
int dbgphmm_dataset_params(const dbgphmm_dataset* d, dbgphmm_params* out);
int dbgphmm_dataset_to_json_text(const dbgphmm_dataset* d, char* buf, uint64_t cap, uint64_t* needed);
int dbgphmm_dataset_to_json_file(const dbgphmm_dataset* d, const char* path);

/* ---- instrumentation ---------------------------------------------------------------------------------- */
/* Kernel launches issued by this library since the last reset (bench.py's gpu_launches). */
uint64_t dbgphmm_launch_count(int reset);
/* Milliseconds of the last bulk call spent in {dense kernels, sparse kernels, product kernels, total},
 * measured with CUDA events on the library's stream; and the algorithmic cell count of the dense kernels. */
int dbgphmm_last_timing(double ms[4], uint64_t* dense_cells);
/* The dominant kernel (dense forward/backward row step) of the last bulk call: summed launch durations from CUDA
 * events recorded around every launch on the library's stream, number of launches, and the cells they computed. */
int dbgphmm_last_dense_kernel(double* ms, uint64_t* launches, uint64_t* cells);
/* Reads whose sparse DP rows are resident on the device at once (SMs x sparse jobs per SM): the sparse rows are latency-bound per read, so
 * the bulk calls (run_node_freqs, to_full_prob_reads, generate_mappings) are most efficient on read sets that are multiples of it. */
uint32_t dbgphmm_model_wave_reads(const dbgphmm_model* m);
/* Upload-free variant for benchmarking: move a reads handle's bases to the device once (idempotent). */
int dbgphmm_reads_to_device(dbgphmm_model* m, dbgphmm_reads* r);

#ifdef __cplusplus
}
#endif
#endif /* DBGPHMM_B200_H */
