//! hmmv2_cuda — the reference-side binding of include/dbgphmm_b200.h.
//!
//! SOURCE ONLY: this image has no cargo/rustc, so this file has never been compiled (tests/test_abi.py checks what can be
//! checked without a compiler: every binding below against the header's names and argument counts, the field order of
//! `dbgphmm_params`, and build.rs against the real build).  It is the shim a dbgphmm maintainer would add so that the
//! callers of `src/hmmv2` (multi_dbg/posterior.rs:247-255,504-515,609-630, multi_dbg/draft.rs:201) keep their method
//! names: `forward`, `backward`, `run_sparse`, `to_node_freqs`, `to_full_prob_reads`, `generate_mappings`, ...
//! Graph flattening follows hmmv2/common.rs:61-67 (PModel = DiGraph<PNode, PEdge>), see INTEGRATION.md.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dbgphmm_params {
    pub p_mismatch: f64, pub p_match: f64, pub p_random: f64, pub p_gap_open: f64, pub p_gap_ext: f64, pub p_end: f64,
    pub p_mm: f64, pub p_im: f64, pub p_dm: f64, pub p_mi: f64, pub p_ii: f64, pub p_di: f64, pub p_md: f64, pub p_id: f64, pub p_dd: f64,
    pub n_active_nodes: u32, pub n_warmup: u32, pub warmup_threshold: u32, pub n_max_gaps: u32,
    pub active_node_max_ratio: f64,
}
#[repr(C)] pub struct dbgphmm_model { _p: [u8; 0] }
#[repr(C)] pub struct dbgphmm_reads { _p: [u8; 0] }
#[repr(C)] pub struct dbgphmm_mappings { _p: [u8; 0] }
#[repr(C)] pub struct dbgphmm_tables { _p: [u8; 0] }
#[repr(C)] pub struct dbgphmm_dbg { _p: [u8; 0] }
#[repr(C)] pub struct dbgphmm_dataset { _p: [u8; 0] }

pub const MAX_ACTIVE_NODES: usize = 400; // hmmv2/table.rs:22

extern "C" {
    pub fn dbgphmm_last_error() -> *const c_char;
    pub fn dbgphmm_device_count() -> c_int;
    pub fn dbgphmm_params_new(p_mismatch: f64, p_gap_open: f64, p_gap_ext: f64, p_end: f64, n_active_nodes: u32, n_warmup: u32,
                              out: *mut dbgphmm_params);
    pub fn dbgphmm_params_uniform(p: f64, out: *mut dbgphmm_params);
    // model
    pub fn dbgphmm_model_create(n_nodes: u32, n_edges: u32, edge_src: *const u32, edge_dst: *const u32, emission: *const u8,
                                log_init: *const f64, log_trans: *const f64, params: *const dbgphmm_params, device: c_int,
                                mem_budget_bytes: u64, out: *mut *mut dbgphmm_model) -> c_int;
    pub fn dbgphmm_model_destroy(m: *mut dbgphmm_model);
    pub fn dbgphmm_model_set_params(m: *mut dbgphmm_model, params: *const dbgphmm_params) -> c_int;
    pub fn dbgphmm_model_set_probs(m: *mut dbgphmm_model, log_init: *const f64, log_trans: *const f64) -> c_int;
    pub fn dbgphmm_model_set_copy_nums_batch(m: *mut dbgphmm_model, n_batch: u32, copy_nums: *const u32, mode: c_int) -> c_int;
    pub fn dbgphmm_model_get_probs(m: *const dbgphmm_model, x: u32, log_init: *mut f64, log_trans: *mut f64) -> c_int;
    pub fn dbgphmm_model_n_nodes(m: *const dbgphmm_model) -> u32;
    pub fn dbgphmm_model_n_batch(m: *const dbgphmm_model) -> u32;
    // reads / mappings
    pub fn dbgphmm_reads_create(n_reads: u64, offsets: *const u64, bases: *const u8, out: *mut *mut dbgphmm_reads) -> c_int;
    pub fn dbgphmm_reads_destroy(r: *mut dbgphmm_reads);
    pub fn dbgphmm_mappings_create(n_reads: u64, read_off: *const u64, row_off: *const u64, nodes: *const u32, logp: *const f64,
                                   out: *mut *mut dbgphmm_mappings) -> c_int;
    pub fn dbgphmm_mappings_destroy(mp: *mut dbgphmm_mappings);
    pub fn dbgphmm_mappings_sizes(mp: *const dbgphmm_mappings, n_reads: *mut u64, n_rows: *mut u64, n_entries: *mut u64) -> c_int;
    pub fn dbgphmm_mappings_export(mp: *const dbgphmm_mappings, read_off: *mut u64, row_off: *mut u64, nodes: *mut u32, logp: *mut f64) -> c_int;
    pub fn dbgphmm_mappings_to_node_freqs(mp: *const dbgphmm_mappings, n_nodes: u32, freqs: *mut f64) -> c_int;
    pub fn dbgphmm_mappings_map_nodes(mp: *const dbgphmm_mappings, n_nodes_before: u32, map_off: *const u64, map_to: *const u32,
                                      out: *mut *mut dbgphmm_mappings) -> c_int;
    // PHMMTables of one read
    pub fn dbgphmm_forward(m: *mut dbgphmm_model, bases: *const u8, n: u64, kind: c_int, mapping: *const dbgphmm_mappings,
                           read_index: u64, out: *mut *mut dbgphmm_tables) -> c_int;
    pub fn dbgphmm_backward(m: *mut dbgphmm_model, bases: *const u8, n: u64, kind: c_int, mapping: *const dbgphmm_mappings,
                            read_index: u64, fwd: *const dbgphmm_tables, out: *mut *mut dbgphmm_tables) -> c_int;
    pub fn dbgphmm_tables_destroy(t: *mut dbgphmm_tables);
    pub fn dbgphmm_tables_len(t: *const dbgphmm_tables) -> u64;
    pub fn dbgphmm_tables_full_prob(t: *const dbgphmm_tables, logp: *mut f64) -> c_int;
    pub fn dbgphmm_tables_row_info(t: *const dbgphmm_tables, row: i64, info: *mut u64, scalars: *mut f64) -> c_int;
    pub fn dbgphmm_tables_row_export(t: *const dbgphmm_tables, row: i64, ids_mi: *mut u32, m: *mut f64, i: *mut f64,
                                     ids_d: *mut u32, d: *mut f64) -> c_int;
    pub fn dbgphmm_tables_row_top_nodes(t: *const dbgphmm_tables, row: i64, by_ratio: c_int, k: u32, ratio: f64,
                                        out: *mut u32, n_out: *mut u32) -> c_int;
    // PHMMOutput of one read
    pub fn dbgphmm_output_node_freqs(m: *mut dbgphmm_model, fwd: *const dbgphmm_tables, bwd: *const dbgphmm_tables, freqs: *mut f64) -> c_int;
    pub fn dbgphmm_output_edge_and_init_freqs(m: *mut dbgphmm_model, fwd: *const dbgphmm_tables, bwd: *const dbgphmm_tables,
                                              edge_freqs: *mut f64, init_freqs: *mut f64) -> c_int;
    pub fn dbgphmm_q_score_exact(m: *const dbgphmm_model, x: u32, edge_freqs: *const f64, init_freqs: *const f64, out: *mut f64) -> c_int;
    pub fn dbgphmm_output_mapping(m: *mut dbgphmm_model, fwd: *const dbgphmm_tables, bwd: *const dbgphmm_tables, by_ratio: c_int,
                                  n_active: u32, ratio: f64, out: *mut *mut dbgphmm_mappings) -> c_int;
    // bulk calls over a read set (the rayon-parallel entry points of the reference)
    pub fn dbgphmm_to_full_prob_reads(m: *mut dbgphmm_model, reads: *const dbgphmm_reads, mappings: *const dbgphmm_mappings,
                                      use_max_ratio: c_int, out_logp: *mut f64, out_logp_per_read: *mut f64) -> c_int;
    pub fn dbgphmm_run_node_freqs(m: *mut dbgphmm_model, reads: *const dbgphmm_reads, mode: c_int, use_max_ratio: c_int,
                                  mappings: *const dbgphmm_mappings, node_freqs: *mut f64, logp_fwd: *mut f64, logp_bwd: *mut f64,
                                  cells: *mut u64) -> c_int;
    pub fn dbgphmm_run_node_freqs_dev(m: *mut dbgphmm_model, reads: *const dbgphmm_reads, mode: c_int, use_max_ratio: c_int,
                                      mappings: *const dbgphmm_mappings, node_freqs_dev: *mut f64, logp_fwd_dev: *mut f64,
                                      logp_bwd_dev: *mut f64, cells: *mut u64) -> c_int;
    pub fn dbgphmm_generate_mappings(m: *mut dbgphmm_model, reads: *const dbgphmm_reads, mappings: *const dbgphmm_mappings,
                                     use_max_ratio: c_int, out: *mut *mut dbgphmm_mappings) -> c_int;
    // file formats either side of the path (multi_dbg/output.rs:155-345, 455-623) and compact-edge copy numbers (multi_dbg.rs:1041-1066)
    pub fn dbgphmm_dbg_from_text(text: *const c_char, len: u64, out: *mut *mut dbgphmm_dbg) -> c_int;
    pub fn dbgphmm_dbg_from_file(path: *const c_char, out: *mut *mut dbgphmm_dbg) -> c_int;
    pub fn dbgphmm_dbg_destroy(d: *mut dbgphmm_dbg);
    pub fn dbgphmm_dbg_sizes(d: *const dbgphmm_dbg, sizes: *mut u32) -> c_int;
    pub fn dbgphmm_dbg_phmm_graph(d: *const dbgphmm_dbg, edge_src: *mut u32, edge_dst: *mut u32, emission: *mut u8, copy_nums: *mut u32,
                                  compact_edge_of: *mut u32) -> c_int;
    pub fn dbgphmm_dbg_get_copy_nums(d: *const dbgphmm_dbg, compact_copy_nums: *mut u32) -> c_int;
    pub fn dbgphmm_dbg_set_copy_nums(d: *mut dbgphmm_dbg, compact_copy_nums: *const u32) -> c_int;
    pub fn dbgphmm_dbg_expand_copy_nums(d: *const dbgphmm_dbg, n_batch: u32, compact: *const u32, full: *mut u32) -> c_int;
    pub fn dbgphmm_dbg_genome_size(d: *const dbgphmm_dbg, n_batch: u32, compact: *const u32, out: *mut u64) -> c_int;
    pub fn dbgphmm_dbg_n_euler_circuits(d: *const dbgphmm_dbg, n_batch: u32, compact: *const u32, out: *mut f64) -> c_int;
    pub fn dbgphmm_euler_circuit_count(n_nodes: u32, n_edges: u64, edge_src: *const u32, edge_dst: *const u32, multiplicity: *const u32,
                                       allow_multiple_component: c_int, out_ln_count: *mut f64) -> c_int;
    pub fn dbgphmm_prior_normal(x: f64, mu: f64, sigma: f64, out_ln_p: *mut f64) -> c_int;
    pub fn dbgphmm_dbg_to_text(d: *const dbgphmm_dbg, buf: *mut c_char, cap: u64, needed: *mut u64) -> c_int;
    pub fn dbgphmm_dbg_to_file(d: *const dbgphmm_dbg, path: *const c_char) -> c_int;
    pub fn dbgphmm_dbg_to_model(d: *const dbgphmm_dbg, params: *const dbgphmm_params, mode: c_int, device: c_int, mem_budget_bytes: u64,
                                out: *mut *mut dbgphmm_model) -> c_int;
    pub fn dbgphmm_mappings_from_map_text(text: *const c_char, len: u64, out: *mut *mut dbgphmm_mappings) -> c_int;
    pub fn dbgphmm_mappings_from_map_file(path: *const c_char, out: *mut *mut dbgphmm_mappings) -> c_int;
    pub fn dbgphmm_mappings_to_map_text(mp: *const dbgphmm_mappings, reads: *const dbgphmm_reads, d: *const dbgphmm_dbg, buf: *mut c_char,
                                        cap: u64, needed: *mut u64) -> c_int;
    pub fn dbgphmm_mappings_to_map_file(mp: *const dbgphmm_mappings, reads: *const dbgphmm_reads, d: *const dbgphmm_dbg, path: *const c_char) -> c_int;
    // dataset JSON (Dataset::to_json_file / from_json_file, e2e.rs:123-130)
    pub fn dbgphmm_dataset_from_json_text(text: *const c_char, len: u64, out: *mut *mut dbgphmm_dataset) -> c_int;
    pub fn dbgphmm_dataset_from_json_file(path: *const c_char, out: *mut *mut dbgphmm_dataset) -> c_int;
    pub fn dbgphmm_dataset_create(n_haps: u32, hap_off: *const u64, hap_bases: *const u8, hap_style: *const u8, genome_size: u64,
                                  n_reads: u64, read_off: *const u64, read_bases: *const u8, read_revcomp: *const u8, origin_hap: *const i64,
                                  origin_pos: *const u64, params: *const dbgphmm_params, out: *mut *mut dbgphmm_dataset) -> c_int;
    pub fn dbgphmm_dataset_destroy(d: *mut dbgphmm_dataset);
    pub fn dbgphmm_dataset_sizes(d: *const dbgphmm_dataset, sizes: *mut u64) -> c_int;
    pub fn dbgphmm_dataset_genome(d: *const dbgphmm_dataset, hap_off: *mut u64, bases: *mut u8, style: *mut u8) -> c_int;
    pub fn dbgphmm_dataset_reads(d: *const dbgphmm_dataset, out: *mut *mut dbgphmm_reads) -> c_int;
    pub fn dbgphmm_dataset_read_origins(d: *const dbgphmm_dataset, read_off: *mut u64, bases: *mut u8, revcomp: *mut u8, origin_hap: *mut i64,
                                        origin_pos: *mut u64) -> c_int;
    pub fn dbgphmm_dataset_params(d: *const dbgphmm_dataset, out: *mut dbgphmm_params) -> c_int;
    pub fn dbgphmm_dataset_to_json_text(d: *const dbgphmm_dataset, buf: *mut c_char, cap: u64, needed: *mut u64) -> c_int;
    pub fn dbgphmm_dataset_to_json_file(d: *const dbgphmm_dataset, path: *const c_char) -> c_int;
    // instrumentation
    pub fn dbgphmm_launch_count(reset: c_int) -> u64;
    pub fn dbgphmm_last_timing(ms: *mut f64, dense_cells: *mut u64) -> c_int;
    pub fn dbgphmm_last_dense_kernel(ms: *mut f64, launches: *mut u64, cells: *mut u64) -> c_int;
    pub fn dbgphmm_model_wave_reads(m: *const dbgphmm_model) -> u32;
    pub fn dbgphmm_reads_to_device(m: *mut dbgphmm_model, r: *mut dbgphmm_reads) -> c_int;
}

fn check(st: c_int) {
    if st != 0 {
        // the reference's hot path never returns Result: it panics (table.rs:388, prob.rs:275, float.rs:11)
        let msg = unsafe { std::ffi::CStr::from_ptr(dbgphmm_last_error()) }.to_string_lossy().into_owned();
        panic!("dbgphmm_b200 status {st}: {msg}");
    }
}

/// `PHMMParams::uniform(p)` (params.rs:116-124) as the C struct; `MultiDbg::to_phmm` then sets `n_warmup = k` (multi_dbg.rs:1395).
pub fn params_uniform(p: f64) -> dbgphmm_params {
    let mut q = std::mem::MaybeUninit::<dbgphmm_params>::uninit();
    unsafe { dbgphmm_params_uniform(p, q.as_mut_ptr()); q.assume_init() }
}

/// `ReadCollection<S>` (common/collection.rs:131) as the CSR the library takes: `offsets[n + 1]` into the concatenated bases.
pub struct CudaReads { h: *mut dbgphmm_reads, n: usize }
impl CudaReads {
    pub fn new<S: AsRef<[u8]>>(reads: &[S]) -> Self {
        let mut off = Vec::with_capacity(reads.len() + 1);
        let mut bases = Vec::new();
        off.push(0u64);
        for r in reads { bases.extend_from_slice(r.as_ref()); off.push(bases.len() as u64); }
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_reads_create(reads.len() as u64, off.as_ptr(), bases.as_ptr(), &mut h) });
        CudaReads { h, n: reads.len() }
    }
    pub fn len(&self) -> usize { self.n }
}
impl Drop for CudaReads { fn drop(&mut self) { unsafe { dbgphmm_reads_destroy(self.h) } } }

/// `e2e::Dataset` (e2e.rs:31-130) in the reference's JSON form (`Dataset::to_json_file` / `from_json_file`): genome, genome size,
/// positioned reads and the PHMM parameters they were sampled with.  `reads()` is the read set the hot path takes.
pub struct CudaDataset { h: *mut dbgphmm_dataset }
impl CudaDataset {
    pub fn from_json_file(path: &str) -> Self {
        let c = std::ffi::CString::new(path).unwrap();
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_dataset_from_json_file(c.as_ptr(), &mut h) });
        CudaDataset { h }
    }
    pub fn to_json_file(&self, path: &str) {
        let c = std::ffi::CString::new(path).unwrap();
        check(unsafe { dbgphmm_dataset_to_json_file(self.h, c.as_ptr()) });
    }
    fn sizes(&self) -> [u64; 5] { let mut s = [0u64; 5]; check(unsafe { dbgphmm_dataset_sizes(self.h, s.as_mut_ptr()) }); s }
    pub fn genome_size(&self) -> usize { self.sizes()[4] as usize }
    /// `Dataset::coverage` (e2e.rs:72-74)
    pub fn coverage(&self) -> f64 { let s = self.sizes(); s[3] as f64 / s[4].max(1) as f64 }
    pub fn params(&self) -> dbgphmm_params {
        let mut p = std::mem::MaybeUninit::<dbgphmm_params>::uninit();
        check(unsafe { dbgphmm_dataset_params(self.h, p.as_mut_ptr()) });
        unsafe { p.assume_init() }
    }
    pub fn reads(&self) -> CudaReads {
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_dataset_reads(self.h, &mut h) });
        CudaReads { h, n: self.sizes()[2] as usize }
    }
}
impl Drop for CudaDataset { fn drop(&mut self) { unsafe { dbgphmm_dataset_destroy(self.h) } } }

/// `Mappings` (hint.rs:150-152): per read, per base, the candidate nodes and their probabilities (hint.rs:27-30).
pub struct CudaMappings { h: *mut dbgphmm_mappings }
impl CudaMappings {
    /// rows[r][i] = (nodes, ln probs) of base i of read r — what `Mapping { nodes, probs }` holds.
    pub fn new(rows: &[Vec<(Vec<u32>, Vec<f64>)>]) -> Self {
        let (mut read_off, mut row_off, mut nodes, mut logp) = (vec![0u64], vec![0u64], Vec::new(), Vec::new());
        for read in rows {
            for (ns, ps) in read {
                assert_eq!(ns.len(), ps.len());
                nodes.extend_from_slice(ns); logp.extend_from_slice(ps);
                row_off.push(nodes.len() as u64);
            }
            read_off.push(row_off.len() as u64 - 1);
        }
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_mappings_create(rows.len() as u64, read_off.as_ptr(), row_off.as_ptr(), nodes.as_ptr(), logp.as_ptr(), &mut h) });
        CudaMappings { h }
    }
    /// `Mappings::to_node_freqs` (hint.rs:161-171) = `MultiDbg::mappings_to_freqs` (multi_dbg/draft.rs:201-212)
    pub fn to_node_freqs(&self, n_nodes: usize) -> Vec<f64> {
        let mut f = vec![0f64; n_nodes];
        check(unsafe { dbgphmm_mappings_to_node_freqs(self.h, n_nodes as u32, f.as_mut_ptr()) });
        f
    }
    /// `Mapping::map_nodes` for every read (hint.rs:66-88): `images[v]` = the nodes of the new graph that old node `v` becomes
    /// (`hint_kp1_from_hint_k`, multi_dbg.rs:1325-1334; `PurgeEdgeMap::update_mapping`, multi_dbg.rs:1783-1791).
    pub fn map_nodes(&self, images: &[Vec<u32>]) -> CudaMappings {
        let (mut off, mut to) = (vec![0u64], Vec::new());
        for im in images { to.extend_from_slice(im); off.push(to.len() as u64); }
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_mappings_map_nodes(self.h, images.len() as u32, off.as_ptr(), to.as_ptr(), &mut h) });
        CudaMappings { h }
    }
    fn ptr(m: Option<&CudaMappings>) -> *const dbgphmm_mappings { m.map_or(std::ptr::null(), |m| m.h as *const _) }
}
impl Drop for CudaMappings { fn drop(&mut self) { unsafe { dbgphmm_mappings_destroy(self.h) } } }

/// `PHMMTables` of one read (table.rs:365-435), resident on the device.
pub struct CudaTables { h: *mut dbgphmm_tables }
impl CudaTables {
    pub fn n_emissions(&self) -> usize { unsafe { dbgphmm_tables_len(self.h) as usize } }
    /// `PHMMTables::full_prob` (table.rs:395-401) as a natural log (`Prob.0`)
    pub fn full_prob(&self) -> f64 {
        let mut p = 0f64;
        check(unsafe { dbgphmm_tables_full_prob(self.h, &mut p) });
        p
    }
    /// `PHMMTable::top_nodes(k)` (table.rs:127) of row `i`
    pub fn top_nodes(&self, i: usize, k: usize) -> Vec<u32> { self.top(i, 0, k as u32, 0.0) }
    /// `PHMMTable::top_nodes_by_score_ratio(ratio)` (table.rs:134) of row `i`
    pub fn top_nodes_by_score_ratio(&self, i: usize, ratio: f64) -> Vec<u32> { self.top(i, 1, 0, ratio) }
    fn top(&self, i: usize, by_ratio: c_int, k: u32, ratio: f64) -> Vec<u32> {
        let mut out = vec![0u32; MAX_ACTIVE_NODES];
        let mut n = 0u32;
        check(unsafe { dbgphmm_tables_row_top_nodes(self.h, i as i64, by_ratio, k, ratio, out.as_mut_ptr(), &mut n) });
        out.truncate(n as usize);
        out
    }
}
impl Drop for CudaTables { fn drop(&mut self) { unsafe { dbgphmm_tables_destroy(self.h) } } }

/// `PHMMOutput { forward, backward }` of one read (table.rs:450-517).
pub struct CudaOutput<'a> { model: &'a CudaPHMM, pub forward: CudaTables, pub backward: CudaTables }
impl<'a> CudaOutput<'a> {
    /// `PHMMOutput::to_node_freqs` (freq.rs:245-255)
    pub fn to_node_freqs(&self) -> Vec<f64> {
        let mut f = vec![0f64; self.model.n_nodes];
        check(unsafe { dbgphmm_output_node_freqs(self.model.h, self.forward.h, self.backward.h, f.as_mut_ptr()) });
        f
    }
    /// `PHMMOutput::to_edge_and_init_freqs` (freq.rs:276-298): (edge freqs in EdgeIndex order, Begin -> node freqs)
    pub fn to_edge_and_init_freqs(&self) -> (Vec<f64>, Vec<f64>) {
        let (mut e, mut i) = (vec![0f64; self.model.n_edges], vec![0f64; self.model.n_nodes]);
        check(unsafe { dbgphmm_output_edge_and_init_freqs(self.model.h, self.forward.h, self.backward.h, e.as_mut_ptr(), i.as_mut_ptr()) });
        (e, i)
    }
    /// `PHMMOutput::to_mapping(n_active_nodes)` (hint.rs:124-133)
    pub fn to_mapping(&self, n_active_nodes: usize) -> CudaMappings {
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_output_mapping(self.model.h, self.forward.h, self.backward.h, 0, n_active_nodes as u32, 0.0, &mut h) });
        CudaMappings { h }
    }
    /// `PHMMOutput::to_mapping_by_score_ratio(ratio)` (hint.rs:134-142)
    pub fn to_mapping_by_score_ratio(&self, ratio: f64) -> CudaMappings {
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_output_mapping(self.model.h, self.forward.h, self.backward.h, 1, 0, ratio, &mut h) });
        CudaMappings { h }
    }
}

/// Device-resident PHMMModel.  In dbgphmm this would be built by `impl From<&PModel> for CudaPHMM`:
/// edges in `graph.edge_references()` order, `emission()`/`init_prob().to_log_value()` per node,
/// `trans_prob().to_log_value()` per edge (hmmv2/common.rs:202-261).
pub struct CudaPHMM { h: *mut dbgphmm_model, n_nodes: usize, n_edges: usize }

impl CudaPHMM {
    pub fn new(edge_src: &[u32], edge_dst: &[u32], emission: &[u8], log_init: &[f64], log_trans: &[f64], param: &dbgphmm_params, device: i32) -> Self {
        assert_eq!(edge_src.len(), edge_dst.len());
        assert_eq!(edge_src.len(), log_trans.len());
        assert_eq!(emission.len(), log_init.len());
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_model_create(emission.len() as u32, edge_src.len() as u32, edge_src.as_ptr(), edge_dst.as_ptr(), emission.as_ptr(),
                                            log_init.as_ptr(), log_trans.as_ptr(), param, device, 0, &mut h) });
        CudaPHMM { h, n_nodes: emission.len(), n_edges: edge_src.len() }
    }
    fn fwd(&self, emissions: &[u8], kind: c_int, mapping: Option<&CudaMappings>, read_index: usize) -> CudaTables {
        let mut t = std::ptr::null_mut();
        check(unsafe { dbgphmm_forward(self.h, emissions.as_ptr(), emissions.len() as u64, kind, CudaMappings::ptr(mapping), read_index as u64, &mut t) });
        CudaTables { h: t }
    }
    fn bwd(&self, emissions: &[u8], kind: c_int, mapping: Option<&CudaMappings>, read_index: usize, fwd: Option<&CudaTables>) -> CudaTables {
        let mut t = std::ptr::null_mut();
        let f = fwd.map_or(std::ptr::null(), |f| f.h as *const _);
        check(unsafe { dbgphmm_backward(self.h, emissions.as_ptr(), emissions.len() as u64, kind, CudaMappings::ptr(mapping), read_index as u64, f, &mut t) });
        CudaTables { h: t }
    }
    /// `PHMMModel::forward` (forward.rs:24)
    pub fn forward(&self, emissions: &[u8]) -> CudaTables { self.fwd(emissions, 0, None, 0) }
    /// `PHMMModel::forward_sparse(emissions, use_max_ratio)` (forward.rs:93)
    pub fn forward_sparse(&self, emissions: &[u8], use_max_ratio: bool) -> CudaTables { self.fwd(emissions, if use_max_ratio { 2 } else { 1 }, None, 0) }
    /// `PHMMModel::forward_with_mapping` (forward.rs:51); the mapping of read `read_index` of `mappings`
    pub fn forward_with_mapping(&self, emissions: &[u8], mappings: &CudaMappings, read_index: usize) -> CudaTables { self.fwd(emissions, 3, Some(mappings), read_index) }
    /// `PHMMModel::backward` (backward.rs:24)
    pub fn backward(&self, emissions: &[u8]) -> CudaTables { self.bwd(emissions, 0, None, 0, None) }
    /// `PHMMModel::backward_sparse` (backward.rs:146)
    pub fn backward_sparse(&self, emissions: &[u8]) -> CudaTables { self.bwd(emissions, 1, None, 0, None) }
    /// `PHMMModel::backward_with_mapping` (backward.rs:59)
    pub fn backward_with_mapping(&self, emissions: &[u8], mappings: &CudaMappings, read_index: usize) -> CudaTables { self.bwd(emissions, 2, Some(mappings), read_index, None) }
    /// `PHMMModel::backward_by_forward` (backward.rs:101)
    pub fn backward_by_forward(&self, emissions: &[u8], forward: &CudaTables) -> CudaTables { self.bwd(emissions, 3, None, 0, Some(forward)) }
    /// `PHMMModel::run` (freq.rs:42)
    pub fn run(&self, emissions: &[u8]) -> CudaOutput<'_> {
        CudaOutput { model: self, forward: self.forward(emissions), backward: self.backward(emissions) }
    }
    /// `PHMMModel::run_sparse` (freq.rs:51)
    pub fn run_sparse(&self, emissions: &[u8]) -> CudaOutput<'_> {
        CudaOutput { model: self, forward: self.forward_sparse(emissions, false), backward: self.backward_sparse(emissions) }
    }
    /// `PHMMModel::run_sparse_adaptive` (freq.rs:60): backward over the nodes the forward pass kept
    pub fn run_sparse_adaptive(&self, emissions: &[u8], use_max_ratio: bool) -> CudaOutput<'_> {
        let forward = self.forward_sparse(emissions, use_max_ratio);
        let backward = self.backward_by_forward(emissions, &forward);
        CudaOutput { model: self, forward, backward }
    }
    /// `PHMMModel::run_with_mapping` (freq.rs:72)
    pub fn run_with_mapping(&self, emissions: &[u8], mappings: &CudaMappings, read_index: usize) -> CudaOutput<'_> {
        CudaOutput { model: self, forward: self.forward_with_mapping(emissions, mappings, read_index),
                     backward: self.backward_with_mapping(emissions, mappings, read_index) }
    }
    /// `PHMMModel::to_full_prob_reads` (freq.rs:175-192) for every candidate X installed by `set_copy_nums_batch`
    /// (one value when none was): ln P(R|X) in candidate order.
    pub fn to_full_prob_reads(&self, reads: &CudaReads, mappings: Option<&CudaMappings>, use_max_ratio: bool) -> Vec<f64> {
        let n_batch = unsafe { dbgphmm_model_n_batch(self.h) } as usize;
        let mut out = vec![0f64; n_batch];
        check(unsafe { dbgphmm_to_full_prob_reads(self.h, reads.h, CudaMappings::ptr(mappings), use_max_ratio as c_int, out.as_mut_ptr(), std::ptr::null_mut()) });
        out
    }
    /// run*() + `to_node_freqs` summed over reads (freq.rs:42-102,245-255); mode: 0 run, 1 run_sparse, 2 run_sparse_adaptive, 3 run_with_mapping
    pub fn to_node_freqs(&self, reads: &CudaReads, mode: i32, use_max_ratio: bool, mappings: Option<&CudaMappings>) -> Vec<f64> {
        let mut f = vec![0f64; self.n_nodes];
        check(unsafe { dbgphmm_run_node_freqs(self.h, reads.h, mode, use_max_ratio as c_int, CudaMappings::ptr(mappings), f.as_mut_ptr(),
                                              std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) });
        f
    }
    /// `PHMMModel::generate_mappings` (hint.rs:193-220)
    pub fn generate_mappings(&self, reads: &CudaReads, mappings: Option<&CudaMappings>, use_max_ratio: bool) -> CudaMappings {
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_generate_mappings(self.h, reads.h, CudaMappings::ptr(mappings), use_max_ratio as c_int, &mut h) });
        CudaMappings { h }
    }
    /// `MultiDbg::set_copy_nums` + `to_phmm` for a batch of candidates (multi_dbg.rs:1041-1052,1394-1397; seq_graph.rs:160-223)
    pub fn set_copy_nums_batch(&mut self, n_batch: usize, node_copy_nums: &[u32], non_zero: bool) {
        assert_eq!(node_copy_nums.len(), n_batch * self.n_nodes);
        check(unsafe { dbgphmm_model_set_copy_nums_batch(self.h, n_batch as u32, node_copy_nums.as_ptr(), if non_zero { 1 } else { 0 }) });
    }
    /// `q_score_exact` (q.rs:66-96) of candidate `x`: (init, trans, prior)
    pub fn q_score_exact(&self, x: usize, edge_freqs: &[f64], init_freqs: &[f64]) -> (f64, f64, f64) {
        assert_eq!(edge_freqs.len(), self.n_edges);
        assert_eq!(init_freqs.len(), self.n_nodes);
        let mut q = [0f64; 3];
        check(unsafe { dbgphmm_q_score_exact(self.h, x as u32, edge_freqs.as_ptr(), init_freqs.as_ptr(), q.as_mut_ptr()) });
        (q[0], q[1], q[2])
    }
}
/// `Score` (multi_dbg/posterior.rs:164-208), natural logs.
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct Score { pub likelihood: f64, pub prior: f64, pub genome_size: u64, pub n_euler_circuits: f64 }
impl Score {
    /// `P(R|X) P(G) #circuits` (posterior.rs:199-201)
    pub fn p(&self) -> f64 { self.likelihood + self.prior + self.n_euler_circuits }
}

/// The part of `MultiDbg` a DBG file carries (multi_dbg.rs:170-186): compact + full graph, copy numbers.  Host only.
pub struct CudaDbg { h: *mut dbgphmm_dbg, k: usize, n_edges_full: usize, n_edges_compact: usize, n_phmm_edges: usize }
impl CudaDbg {
    fn adopt(h: *mut dbgphmm_dbg) -> Self {
        let mut sz = [0u32; 6];
        check(unsafe { dbgphmm_dbg_sizes(h, sz.as_mut_ptr()) });
        CudaDbg { h, k: sz[0] as usize, n_edges_full: sz[2] as usize, n_edges_compact: sz[4] as usize, n_phmm_edges: sz[5] as usize }
    }
    /// `MultiDbg::from_dbg_file` (multi_dbg/output.rs:346-357; .dbg, .dbg.gz, .dbz)
    pub fn from_dbg_file(path: &str) -> Self {
        let c = std::ffi::CString::new(path).unwrap();
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_dbg_from_file(c.as_ptr(), &mut h) });
        Self::adopt(h)
    }
    pub fn k(&self) -> usize { self.k }
    pub fn n_edges_full(&self) -> usize { self.n_edges_full }
    pub fn n_edges_compact(&self) -> usize { self.n_edges_compact }
    /// `MultiDbg::get_copy_nums` / `set_copy_nums` (multi_dbg.rs:1041-1066; set panics on copy numbers that do not balance)
    pub fn get_copy_nums(&self) -> Vec<u32> {
        let mut x = vec![0u32; self.n_edges_compact];
        check(unsafe { dbgphmm_dbg_get_copy_nums(self.h, x.as_mut_ptr()) });
        x
    }
    pub fn set_copy_nums(&mut self, copy_nums: &[u32]) {
        assert_eq!(copy_nums.len(), self.n_edges_compact);
        check(unsafe { dbgphmm_dbg_set_copy_nums(self.h, copy_nums.as_ptr()) });
    }
    /// `MultiDbg::genome_size` (multi_dbg.rs:1018-1028)
    pub fn genome_size(&self) -> u64 {
        let mut g = 0u64;
        check(unsafe { dbgphmm_dbg_genome_size(self.h, 1, std::ptr::null(), &mut g) });
        g
    }
    /// `MultiDbg::n_euler_circuits` (multi_dbg.rs:831-837), natural log
    pub fn n_euler_circuits(&self) -> f64 {
        let mut v = 0f64;
        check(unsafe { dbgphmm_dbg_n_euler_circuits(self.h, 1, std::ptr::null(), &mut v) });
        v
    }
    /// `MultiDbg::to_phmm` (multi_dbg.rs:1394-1397; `n_warmup := k`)
    pub fn to_phmm(&self, param: &dbgphmm_params, device: i32) -> CudaPHMM {
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_dbg_to_model(self.h, param, 0, device, 0, &mut h) });
        CudaPHMM { h, n_nodes: self.n_edges_full, n_edges: self.n_phmm_edges }
    }
    /// `MultiDbg::to_score` (posterior.rs:259-277) for every candidate of `sample_posterior_once` (posterior.rs:504-515) at once:
    /// `candidates` is `[n_batch][n_edges_compact]`, flattened.  One expansion to k-mer copy numbers, one on-device derivation of
    /// (init, trans), one batched `to_full_prob_reads` (use_max_ratio = true like `to_likelihood`), then the host-side terms.
    pub fn to_scores(&self, phmm: &mut CudaPHMM, reads: &CudaReads, mappings: Option<&CudaMappings>, n_batch: usize, candidates: &[u32],
                     genome_size_expected: u32, genome_size_sigma: u32) -> Vec<Score> {
        assert_eq!(candidates.len(), n_batch * self.n_edges_compact);
        let mut full = vec![0u32; n_batch * self.n_edges_full];
        check(unsafe { dbgphmm_dbg_expand_copy_nums(self.h, n_batch as u32, candidates.as_ptr(), full.as_mut_ptr()) });
        phmm.set_copy_nums_batch(n_batch, &full, false);
        let like = phmm.to_full_prob_reads(reads, mappings, true);
        let (mut gs, mut ne) = (vec![0u64; n_batch], vec![0f64; n_batch]);
        check(unsafe { dbgphmm_dbg_genome_size(self.h, n_batch as u32, candidates.as_ptr(), gs.as_mut_ptr()) });
        check(unsafe { dbgphmm_dbg_n_euler_circuits(self.h, n_batch as u32, candidates.as_ptr(), ne.as_mut_ptr()) });
        (0..n_batch).map(|b| {
            let mut prior = 0f64;
            check(unsafe { dbgphmm_prior_normal(gs[b] as f64, genome_size_expected as f64, genome_size_sigma as f64, &mut prior) });
            Score { likelihood: like[b], prior, genome_size: gs[b], n_euler_circuits: ne[b] }
        }).collect()
    }
}
impl Drop for CudaDbg { fn drop(&mut self) { unsafe { dbgphmm_dbg_destroy(self.h) } } }
unsafe impl Send for CudaDbg {}

impl Drop for CudaPHMM { fn drop(&mut self) { unsafe { dbgphmm_model_destroy(self.h) } } }
// one host thread at a time per handle (header, "Conventions")
unsafe impl Send for CudaPHMM {}
unsafe impl Send for CudaReads {}
unsafe impl Send for CudaMappings {}
unsafe impl Send for CudaTables {}
