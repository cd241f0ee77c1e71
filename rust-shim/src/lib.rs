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
    pub block_cells: *mut u64,
}

#[repr(C)]
pub struct FbParts {
    pub n_parts: u64,
    pub part_ptr: *mut u64,
    pub read_ids: *mut u32,
    pub range_lo: *mut u32,
    pub range_hi: *mut u32,
}

#[repr(C)]
pub struct FbMulti {
    _private: [u8; 0],
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
    pub fn fb_beam_search_phasing(
        ctx: *mut FbCtx, frags: *const FbFrags, n_sel: u64, sel: *const u32, ploidy: u32, params: *const FbParams,
        hap_out: *mut u8, best_score: *mut f64, tap_same: *mut f64, tap_diff: *mut f64, tap_logp: *mut f64, tap_cap: u64,
        tap_n: *mut u64,
    ) -> c_int;
    pub fn fb_optimize_clustering(
        ctx: *mut FbCtx, frags: *const FbFrags, n_sel: u64, sel: *const u32, hap_in: *const u8, ploidy: u32,
        params: *const FbParams, hap_out: *mut u8, score: *mut f64, n_rounds: *mut u32,
    ) -> c_int;
    pub fn fb_process_reads_for_final_parts(
        ctx: *mut FbCtx, frags: *const FbFrags, n_parts: u64, part_ptr: *const u64, part_reads: *const u32,
        range_lo: *const u32, range_hi: *const u32, params: *const FbParams, out: *mut *mut FbParts,
    ) -> c_int;
    pub fn fb_free_parts(p: *mut FbParts);
    pub fn fb_get_hapq(
        ctx: *mut FbCtx, frags: *const FbFrags, n_parts: u64, part_ptr: *const u64, part_reads: *const u32,
        range_lo: *const u32, range_hi: *const u32, snp_to_genome_pos: *const u64, n_snps: u64, params: *const FbParams,
        hapq: *mut u8, rel_err: *mut f64, avg_err: *mut f64,
    ) -> c_int;
    // several devices in one process: contigs dealt to the devices by the library's static LPT queue
    pub fn fb_init_multi(n_devices: c_int, device_ids: *const c_int, out: *mut *mut FbMulti) -> c_int;
    pub fn fb_destroy_multi(m: *mut FbMulti);
    pub fn fb_multi_last_error(m: *const FbMulti) -> *const c_char;
    pub fn fb_phase_contigs(
        m: *mut FbMulti, n_contigs: u64, contigs: *const FbFrags, blk_ptr: *const u64, blk_lo: *const u32,
        blk_hi: *const u32, params: *const FbParams, out: *mut *mut FbBlockResults, device_of: *mut u32,
        device_ms: *mut f32,
    ) -> c_int;
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


// =====================================================================================================================
// Wrappers with the PRESERVED signatures of the reference's pub fns (north star: "the local_clustering / part_block_manip
// function signatures preserved so the new path drops in").  They need no context argument: a per-thread context on
// device FLORIA_B200_DEVICE (default 0) is opened on first use.  Reads are identified by counter_id; every wrapper flattens
// exactly the reads it is given (sorted by Frag::cmp, renumbered 0..n) and maps the answer back to `&Frag`.
// =====================================================================================================================
use floria::types_structs::HapBlock;
use fxhash::FxHashMap;
use std::cell::RefCell;

thread_local! {
    static CTX: RefCell<*mut FbCtx> = RefCell::new(std::ptr::null_mut());
}

fn ctx() -> *mut FbCtx {
    CTX.with(|c| {
        if c.borrow().is_null() {
            let dev: c_int = std::env::var("FLORIA_B200_DEVICE").ok().and_then(|x| x.parse().ok()).unwrap_or(0);
            let mut p: *mut FbCtx = std::ptr::null_mut();
            let rc = unsafe { fb_init(dev, &mut p) };
            if rc != 0 {
                let msg = unsafe { CStr::from_ptr(fb_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
                panic!("floria_b200: fb_init({}) failed ({}): {} (there is no CPU fallback)", dev, rc, msg);
            }
            *c.borrow_mut() = p;
        }
        *c.borrow()
    })
}

/// the given reads sorted by Frag::cmp (types_structs.rs:87-93) + their flat CSR with local ids 0..n
fn flatten<'a>(reads: impl Iterator<Item = &'a Frag>) -> (Vec<&'a Frag>, FlatFrags, FxHashMap<usize, u32>) {
    let mut v: Vec<&Frag> = reads.collect();
    v.sort();
    v.dedup_by(|a, b| a.counter_id == b.counter_id);
    let mut f = FlatFrags { row_ptr: vec![0], first: vec![], last: vec![], pos: vec![], allele: vec![], qual: vec![] };
    let mut local = FxHashMap::default();
    for (i, frag) in v.iter().enumerate() {
        local.insert(frag.counter_id, i as u32);
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
    (v, f, local)
}

fn params_eps(epsilon: f64) -> FbParams {
    let mut p: FbParams = unsafe { std::mem::zeroed() };
    unsafe { fb_params_default(&mut p) };
    p.epsilon = epsilon;
    p
}

/// global_clustering::beam_search_phasing (src/global_clustering.rs:10-22), same signature.  `clique` must hold empty
/// sets (the only way the reference calls it, graph_processing.rs:140-150); break_positions is returned empty (unused
/// downstream: WEIRD_SPLIT = false, graph_processing.rs:166).
pub fn beam_search_phasing<'a>(
    clique: Vec<FxHashSet<&'a Frag>>,
    all_reads: &'a Vec<&Frag>,
    epsilon: f64,
    div_factor: f64,
    cutoff_value: f64,
    max_number_solns: usize,
    _use_mec: bool,
    _use_ref_bias: bool,
) -> (FxHashMap<SnpPosition, FxHashSet<usize>>, Vec<FxHashSet<&'a Frag>>) {
    if all_reads.is_empty() {
        return (FxHashMap::default(), vec![]);
    }
    assert!(clique.iter().all(|s| s.is_empty()), "floria_b200: a non-empty starting clique is not supported");
    let ploidy = clique.len();
    let (sorted, flat, _) = flatten(all_reads.iter().copied());
    let mut p = params_eps(epsilon);
    p.div_factor = div_factor;
    p.prob_cutoff_ln = cutoff_value;
    p.max_number_solns = max_number_solns as u32;
    let sel: Vec<u32> = (0..sorted.len() as u32).collect();
    let mut hap = vec![0u8; sorted.len()];
    let mut score = 0f64;
    let c = ctx();
    let rc = unsafe {
        fb_beam_search_phasing(c, &flat.as_c(), sel.len() as u64, sel.as_ptr(), ploidy as u32, &p, hap.as_mut_ptr(),
                               &mut score, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), 0,
                               std::ptr::null_mut())
    };
    check(c, rc);
    let mut partition = vec![FxHashSet::default(); ploidy];
    for (i, f) in sorted.iter().enumerate() {
        partition[hap[i] as usize].insert(*f);
    }
    (FxHashMap::default(), partition)
}

/// local_clustering::optimize_clustering (src/local_clustering.rs:71-75), same signature.  The returned HapBlock is
/// rebuilt on the host from the returned partition (utils_frags::hap_block_from_partition, what the reference returns).
pub fn optimize_clustering<'a>(
    partition: Vec<FxHashSet<&'a Frag>>,
    epsilon: f64,
    max_iters: usize,
) -> (f64, Vec<FxHashSet<&'a Frag>>, HapBlock) {
    let ploidy = partition.len();
    let (sorted, flat, local) = flatten(partition.iter().flat_map(|s| s.iter().copied()));
    let mut hap_in = vec![0u8; sorted.len()];
    for (h, set) in partition.iter().enumerate() {
        for f in set.iter() {
            hap_in[local[&f.counter_id] as usize] = h as u8;
        }
    }
    let mut p = params_eps(epsilon);
    p.num_iter_optimize = max_iters as u32;
    let sel: Vec<u32> = (0..sorted.len() as u32).collect();
    let mut hap = vec![0u8; sorted.len()];
    let (mut score, mut rounds) = (0f64, 0u32);
    let c = ctx();
    let rc = unsafe {
        fb_optimize_clustering(c, &flat.as_c(), sel.len() as u64, sel.as_ptr(), hap_in.as_ptr(), ploidy as u32, &p,
                               hap.as_mut_ptr(), &mut score, &mut rounds)
    };
    check(c, rc);
    let mut out = vec![FxHashSet::default(); ploidy];
    for (i, f) in sorted.iter().enumerate() {
        out[hap[i] as usize].insert(*f);
    }
    let block = utils_frags::hap_block_from_partition(&out, true);
    (score, out, block)
}

fn parts_csr<'a>(parts: &Vec<FxHashSet<&'a Frag>>, local: &FxHashMap<usize, u32>) -> (Vec<u64>, Vec<u32>) {
    let mut ptr = vec![0u64];
    let mut ids = vec![];
    for set in parts.iter() {
        let mut v: Vec<u32> = set.iter().map(|f| local[&f.counter_id]).collect();
        v.sort();
        ids.extend(v);
        ptr.push(ids.len() as u64);
    }
    (ptr, ids)
}

/// part_block_manip::process_reads_for_final_parts (src/part_block_manip.rs:174-180), same signature.
/// `--reassign-short` (a hidden flag) is not implemented on the device path: the library rejects it.
pub fn process_reads_for_final_parts<'a>(
    all_joined_path_parts: Vec<FxHashSet<&'a Frag>>,
    _short_frags: &'a Vec<Frag>,
    snp_range_parts_vec: Vec<(SnpPosition, SnpPosition)>,
    options: &Options,
    _snp_to_genome_pos: &'a Vec<usize>,
) -> (Vec<FxHashSet<&'a Frag>>, Vec<(SnpPosition, SnpPosition)>) {
    let (sorted, flat, local) = flatten(all_joined_path_parts.iter().flat_map(|s| s.iter().copied()));
    let (ptr, ids) = parts_csr(&all_joined_path_parts, &local);
    let lo: Vec<u32> = snp_range_parts_vec.iter().map(|x| x.0).collect();
    let hi: Vec<u32> = snp_range_parts_vec.iter().map(|x| x.1).collect();
    let mut p = params_from(options);
    p.reassign_short = options.reassign_short as u32;
    let mut res: *mut FbParts = std::ptr::null_mut();
    let c = ctx();
    let rc = unsafe {
        fb_process_reads_for_final_parts(c, &flat.as_c(), all_joined_path_parts.len() as u64, ptr.as_ptr(), ids.as_ptr(),
                                         lo.as_ptr(), hi.as_ptr(), &p, &mut res)
    };
    check(c, rc);
    let r = unsafe { &*res };
    let mut parts = vec![];
    let mut ranges = vec![];
    for i in 0..r.n_parts as usize {
        let (a, b) = unsafe { (*r.part_ptr.add(i) as usize, *r.part_ptr.add(i + 1) as usize) };
        let mut set = FxHashSet::default();
        for k in a..b {
            set.insert(sorted[unsafe { *r.read_ids.add(k) } as usize]);
        }
        parts.push(set);
        ranges.push(unsafe { (*r.range_lo.add(i), *r.range_hi.add(i)) });
    }
    unsafe { fb_free_parts(res) };
    (parts, ranges)
}

/// part_block_manip::get_hapq (src/part_block_manip.rs:517-522), same signature.
pub fn get_hapq<'a>(
    parts: &Vec<FxHashSet<&'a Frag>>,
    snp_to_genome_pos: &'a Vec<usize>,
    snp_range_parts_vec: &Vec<(SnpPosition, SnpPosition)>,
    options: &Options,
) -> (Vec<u8>, Vec<f64>, f64) {
    let (_sorted, flat, local) = flatten(parts.iter().flat_map(|s| s.iter().copied()));
    let (ptr, ids) = parts_csr(parts, &local);
    let lo: Vec<u32> = snp_range_parts_vec.iter().map(|x| x.0).collect();
    let hi: Vec<u32> = snp_range_parts_vec.iter().map(|x| x.1).collect();
    let g: Vec<u64> = snp_to_genome_pos.iter().map(|x| *x as u64).collect();
    let p = params_from(options);
    let n = parts.len();
    let mut hapq = vec![0u8; n];
    let mut rel = vec![0f64; n];
    let mut avg = 0f64;
    let c = ctx();
    let rc = unsafe {
        fb_get_hapq(c, &flat.as_c(), n as u64, ptr.as_ptr(), ids.as_ptr(), lo.as_ptr(), hi.as_ptr(), g.as_ptr(),
                    g.len() as u64, &p, hapq.as_mut_ptr(), rel.as_mut_ptr(), &mut avg)
    };
    check(c, rc);
    (hapq, rel, avg)
}
