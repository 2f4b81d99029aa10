//! hmmv2_cuda — the reference-side binding of include/dbgphmm_b200.h.
//!
//! SOURCE ONLY: this image has no cargo/rustc, so this file has never been compiled.  It shows the shim a dbgphmm
//! maintainer would add so that `src/hmmv2` callers (multi_dbg/posterior.rs:247-255,609-630, multi_dbg/draft.rs:201)
//! keep their method names.  Graph flattening follows hmmv2/common.rs:61-67 (PModel = DiGraph<PNode, PEdge>).
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
#[repr(C)] pub struct dbgphmm_dbg { _p: [u8; 0] }

extern "C" {
    pub fn dbgphmm_last_error() -> *const c_char;
    pub fn dbgphmm_model_create(n_nodes: u32, n_edges: u32, edge_src: *const u32, edge_dst: *const u32, emission: *const u8,
                                log_init: *const f64, log_trans: *const f64, params: *const dbgphmm_params, device: c_int,
                                mem_budget_bytes: u64, out: *mut *mut dbgphmm_model) -> c_int;
    pub fn dbgphmm_model_destroy(m: *mut dbgphmm_model);
    pub fn dbgphmm_model_set_copy_nums_batch(m: *mut dbgphmm_model, n_batch: u32, copy_nums: *const u32, mode: c_int) -> c_int;
    pub fn dbgphmm_reads_create(n_reads: u64, offsets: *const u64, bases: *const u8, out: *mut *mut dbgphmm_reads) -> c_int;
    pub fn dbgphmm_reads_destroy(r: *mut dbgphmm_reads);
    pub fn dbgphmm_mappings_create(n_reads: u64, read_off: *const u64, row_off: *const u64, nodes: *const u32, logp: *const f64,
                                   out: *mut *mut dbgphmm_mappings) -> c_int;
    pub fn dbgphmm_mappings_destroy(m: *mut dbgphmm_mappings);
    pub fn dbgphmm_mappings_map_nodes(mp: *const dbgphmm_mappings, n_nodes_before: u32, map_off: *const u64, map_to: *const u32,
                                      out: *mut *mut dbgphmm_mappings) -> c_int;
    pub fn dbgphmm_to_full_prob_reads(m: *mut dbgphmm_model, reads: *const dbgphmm_reads, mappings: *const dbgphmm_mappings,
                                      use_max_ratio: c_int, out_logp: *mut f64, out_logp_per_read: *mut f64) -> c_int;
    pub fn dbgphmm_run_node_freqs(m: *mut dbgphmm_model, reads: *const dbgphmm_reads, mode: c_int, use_max_ratio: c_int,
                                  mappings: *const dbgphmm_mappings, node_freqs: *mut f64, logp_fwd: *mut f64, logp_bwd: *mut f64,
                                  cells: *mut u64) -> c_int;
    // file formats either side of the path (multi_dbg/output.rs:155-345, 455-623) and compact-edge copy numbers (multi_dbg.rs:1041-1066)
    pub fn dbgphmm_dbg_from_file(path: *const c_char, out: *mut *mut dbgphmm_dbg) -> c_int;
    pub fn dbgphmm_dbg_destroy(d: *mut dbgphmm_dbg);
    pub fn dbgphmm_dbg_sizes(d: *const dbgphmm_dbg, sizes: *mut u32) -> c_int;
    pub fn dbgphmm_dbg_expand_copy_nums(d: *const dbgphmm_dbg, n_batch: u32, compact: *const u32, full: *mut u32) -> c_int;
    pub fn dbgphmm_dbg_to_model(d: *const dbgphmm_dbg, params: *const dbgphmm_params, mode: c_int, device: c_int, mem_budget_bytes: u64,
                                out: *mut *mut dbgphmm_model) -> c_int;
    pub fn dbgphmm_mappings_from_map_file(path: *const c_char, out: *mut *mut dbgphmm_mappings) -> c_int;
    pub fn dbgphmm_mappings_to_map_file(mp: *const dbgphmm_mappings, reads: *const dbgphmm_reads, d: *const dbgphmm_dbg, path: *const c_char) -> c_int;
    pub fn dbgphmm_generate_mappings(m: *mut dbgphmm_model, reads: *const dbgphmm_reads, mappings: *const dbgphmm_mappings,
                                     use_max_ratio: c_int, out: *mut *mut dbgphmm_mappings) -> c_int;
}

fn check(st: c_int) {
    if st != 0 {
        // the reference's hot path never returns Result: it panics (table.rs:388, prob.rs:275, float.rs:11)
        let msg = unsafe { std::ffi::CStr::from_ptr(dbgphmm_last_error()) }.to_string_lossy().into_owned();
        panic!("dbgphmm_b200 status {st}: {msg}");
    }
}

/// Device-resident PHMMModel.  In dbgphmm this would be built by `impl From<&PModel> for CudaPHMM`:
/// edges in `graph.edge_references()` order, `emission()`/`init_prob().to_log_value()` per node,
/// `trans_prob().to_log_value()` per edge (hmmv2/common.rs:202-261).
pub struct CudaPHMM { h: *mut dbgphmm_model, n_nodes: usize }

impl CudaPHMM {
    pub fn new(edge_src: &[u32], edge_dst: &[u32], emission: &[u8], log_init: &[f64], log_trans: &[f64], param: &dbgphmm_params, device: i32) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { dbgphmm_model_create(emission.len() as u32, edge_src.len() as u32, edge_src.as_ptr(), edge_dst.as_ptr(), emission.as_ptr(),
                                            log_init.as_ptr(), log_trans.as_ptr(), param, device, 0, &mut h) });
        CudaPHMM { h, n_nodes: emission.len() }
    }
    /// PHMMModel::to_full_prob_reads (freq.rs:175-192); `reads` = (offsets, bases), `mappings` as CSR handles.
    pub fn to_full_prob_reads(&self, reads: *const dbgphmm_reads, mappings: *const dbgphmm_mappings, use_max_ratio: bool, n_batch: usize) -> Vec<f64> {
        let mut out = vec![0f64; n_batch];
        check(unsafe { dbgphmm_to_full_prob_reads(self.h, reads, mappings, use_max_ratio as c_int, out.as_mut_ptr(), std::ptr::null_mut()) });
        out
    }
    /// run_sparse + to_node_freqs summed over reads (freq.rs:51-55,245-255); mode: 0 run, 1 run_sparse, 2 run_sparse_adaptive, 3 run_with_mapping
    pub fn to_node_freqs(&self, reads: *const dbgphmm_reads, mode: i32, use_max_ratio: bool, mappings: *const dbgphmm_mappings) -> Vec<f64> {
        let mut f = vec![0f64; self.n_nodes];
        check(unsafe { dbgphmm_run_node_freqs(self.h, reads, mode, use_max_ratio as c_int, mappings, f.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) });
        f
    }
    /// MultiDbg::set_copy_nums + to_phmm for a batch of candidates (multi_dbg.rs:1041-1052,1394-1397; seq_graph.rs:160-223)
    pub fn set_copy_nums_batch(&mut self, n_batch: usize, node_copy_nums: &[u32], non_zero: bool) {
        assert_eq!(node_copy_nums.len(), n_batch * self.n_nodes);
        check(unsafe { dbgphmm_model_set_copy_nums_batch(self.h, n_batch as u32, node_copy_nums.as_ptr(), if non_zero { 1 } else { 0 }) });
    }
}
impl Drop for CudaPHMM { fn drop(&mut self) { unsafe { dbgphmm_model_destroy(self.h) } } }
unsafe impl Send for CudaPHMM {}
