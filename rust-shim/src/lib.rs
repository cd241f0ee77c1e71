//! Rust side of the boundary of include/floria_b200.h.  SOURCE ONLY — not compiled in this environment.
//!
//! `generate_hap_graph_b200` has the signature of `graph_processing::generate_hap_graph`
//! (src/graph_processing.rs:325-330) and replaces the body of its rayon `par_iter` (345-362) by one
//! `fb_phase_blocks` call; HapNode construction (276-303), process_chunks and update_hap_graph are unchanged.
//! Errors follow the reference convention (panic): a non-zero status becomes `panic!`.
use floria::types_structs::{Frag, HapNode, Options, SnpPosition};
use floria::utils_frags;
use fxhash::FxHashSet;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct FbParams {
    pub epsilon: f64,
    pub div_factor: f64,
    pub prob_cutoff_ln: f64,
    pub max_number_solns: u32,
    pub max_ploidy: u32,
    pub num_iter_optimize: u32,
    pub ploidy_sensitivity: u32,
    pub stopping_heuristic: u32,
    pub order_model: u32,
    pub block_length: u32,
    pub reassign_short: u32,
    pub phred_lut: *const f32,
}

#[repr(C)]
pub struct FbFrags {
    pub n_reads: u64,
    pub nnz: u64,
    pub row_ptr: *const u64,
    pub first: *const u32,
    pub last: *const u32,
    pub pos: *const u32,
    pub allele: *const u8,
    pub qual: *const u8,
}

#[repr(C)]
pub struct FbBlockResults {
    pub n_blocks: u64,
    pub max_ploidy: u32,
    _pad: u32,
    pub best_ploidy: *mut u32,
    pub ploidies_run: *mut u32,
    pub mec_vector: *mut f64,
    pub expected_errors: *mut f64,
    pub read_ptr: *mut u64,
    pub read_ids: *mut u32,
    pub hap: *mut u8,
    pub cells_sweep: u64,
    pub cells_hist: u64,
    pub cells_beam: u64,
}

#[repr(C)]
pub struct FbCtx {
    _private: [u8; 0],
}

extern "C" {
    pub fn fb_init(device: c_int, out: *mut *mut FbCtx) -> c_int;
    pub fn fb_destroy(ctx: *mut FbCtx);
    pub fn fb_last_error(ctx: *const FbCtx) -> *const c_char;
    pub fn fb_params_default(p: *mut FbParams);
    pub fn fb_phase_blocks(
        ctx: *mut FbCtx,
        frags: *const FbFrags,
        n_blocks: u64,
        blk_lo: *const u32,
        blk_hi: *const u32,
        params: *const FbParams,
        out: *mut *mut FbBlockResults,
    ) -> c_int;
    pub fn fb_free_block_results(r: *mut FbBlockResults);
}

/// Flat CSR copy of `&Vec<Frag>` (positions ascending within a read).  `counter_id == index` (floria.rs:289-293).
pub struct FlatFrags {
    row_ptr: Vec<u64>,
    first: Vec<u32>,
    last: Vec<u32>,
    pos: Vec<u32>,
    allele: Vec<u8>,
    qual: Vec<u8>,
}

impl FlatFrags {
    pub fn new(all_frags: &Vec<Frag>) -> FlatFrags {
        let mut f = FlatFrags { row_ptr: vec![0], first: vec![], last: vec![], pos: vec![], allele: vec![], qual: vec![] };
        for (i, frag) in all_frags.iter().enumerate() {
            assert_eq!(frag.counter_id, i);
            let mut p: Vec<SnpPosition> = frag.positions.iter().copied().collect();
            p.sort();
            for x in p {
                f.pos.push(x);
                f.allele.push(frag.seq_dict[&x]);
                f.qual.push(frag.qual_dict[&x]);
            }
            f.first.push(frag.first_position);
            f.last.push(frag.last_position);
            f.row_ptr.push(f.pos.len() as u64);
        }
        f
    }
    pub fn as_c(&self) -> FbFrags {
        FbFrags {
            n_reads: self.first.len() as u64,
            nnz: self.pos.len() as u64,
            row_ptr: self.row_ptr.as_ptr(),
            first: self.first.as_ptr(),
            last: self.last.as_ptr(),
            pos: self.pos.as_ptr(),
            allele: self.allele.as_ptr(),
            qual: self.qual.as_ptr(),
        }
    }
}

fn params_from(options: &Options) -> FbParams {
    let mut p: FbParams = unsafe { std::mem::zeroed() };
    unsafe { fb_params_default(&mut p) };
    p.epsilon = options.epsilon;
    p.max_number_solns = options.max_number_solns as u32;
    p.max_ploidy = options.max_ploidy as u32;
    p.ploidy_sensitivity = options.ploidy_sensitivity as u32;
    p.stopping_heuristic = options.stopping_heuristic as u32;
    p.block_length = options.block_length as u32;
    p
}

fn check(ctx: *mut FbCtx, rc: c_int) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(fb_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("floria_b200 error {}: {}", rc, msg); // reference convention: panic / unwrap
    }
}

/// Drop-in for `graph_processing::generate_hap_graph` up to (not including) `update_hap_graph`:
/// returns, per block that has reads, the best-ploidy partition as index sets, in block order.
pub fn phase_blocks_b200<'a>(
    ctx: *mut FbCtx,
    all_frags: &'a Vec<Frag>,
    snp_to_genome_pos: &'a Vec<usize>,
    options: &Options,
) -> Vec<(usize, Vec<FxHashSet<usize>>, (SnpPosition, SnpPosition))> {
    let ranges = utils_frags::get_range_with_lengths(
        snp_to_genome_pos,
        options.block_length,
        options.block_length / 3,
        options.snp_density,
    );
    let lo: Vec<u32> = ranges.iter().map(|x| x.0).collect();
    let hi: Vec<u32> = ranges.iter().map(|x| x.1).collect();
    let flat = FlatFrags::new(all_frags);
    let cfrags = flat.as_c();
    let params = params_from(options);
    let mut res: *mut FbBlockResults = std::ptr::null_mut();
    let rc = unsafe { fb_phase_blocks(ctx, &cfrags, lo.len() as u64, lo.as_ptr(), hi.as_ptr(), &params, &mut res) };
    check(ctx, rc);
    let r = unsafe { &*res };
    let mut out = vec![];
    for j in 0..r.n_blocks as usize {
        let best = unsafe { *r.best_ploidy.add(j) } as usize;
        if best == 0 {
            continue; // get_local_hap_blocks returned None (graph_processing.rs:129-131)
        }
        let (a, b) = unsafe { (*r.read_ptr.add(j) as usize, *r.read_ptr.add(j + 1) as usize) };
        let mut part = vec![FxHashSet::default(); best];
        for k in a..b {
            let (id, h) = unsafe { (*r.read_ids.add(k) as usize, *r.hap.add(k) as usize) };
            part[h].insert(id);
        }
        out.push((j, part, ranges[j]));
    }
    unsafe { fb_free_block_results(res) };
    out
}

/// The HapNode construction of get_local_hap_blocks (graph_processing.rs:276-303), unchanged.
pub fn hap_nodes_from<'a>(
    all_frags: &'a Vec<Frag>,
    blocks: Vec<(usize, Vec<FxHashSet<usize>>, (SnpPosition, SnpPosition))>,
) -> Vec<Vec<HapNode<'a>>> {
    let mut cols = vec![];
    for (_j, part, endpoints) in blocks {
        let mut col = vec![];
        for ind_part in part.iter() {
            let frag_set: FxHashSet<&Frag> = ind_part.iter().map(|x| &all_frags[*x]).collect();
            let mut node = HapNode::new(frag_set, endpoints);
            node.row = col.len();
            col.push(node);
        }
        cols.push(col);
    }
    cols
}
