"""ctypes host binding of libfloria_b200.so (the C-ABI of include/floria_b200.h).

Method names mirror the reference's pub fns (src/lib.rs:1-23) for the hot path so parity tests read like calls
into floria itself.  The library is the only compute path: loading fails loudly when the CUDA build is missing
and fb_init fails when there is no device (no CPU fallback).
"""
import ctypes as C
import os

import numpy as np

from ._cdefs import (FbReaderOptions, FbFragSet, 
    BlockResults,
    FbBlockPhase,
    FbBlockResults,
    FbFrags,
    FbParams,
    FbParts,
    FbTimings,
    Parts,
    default_params,
    f64p,
    i64p,
    ptr,
    u8p,
    u32p,
    u64p,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FB_LIB") or os.path.join(_HERE, "libfloria_b200.so")  # FB_LIB: A/B builds

EXPORTS = [
    "fb_init", "fb_destroy", "fb_last_error", "fb_params_default", "fb_last_timings", "fb_stream",
    "fb_frags_upload", "fb_frags_upload_parts", "fb_frags_free", "fb_dfrags_bytes", "fb_get_range_with_lengths",
    "fb_find_reads_in_interval", "fb_phase_blocks", "fb_phase_blocks_resident", "fb_free_block_results",
    "fb_phase_block", "fb_phase_block_resident",
    "fb_init_multi", "fb_destroy_multi", "fb_multi_size", "fb_multi_ctx", "fb_multi_last_error", "fb_lpt_assign",
    "fb_contigs_upload", "fb_contigs_free", "fb_phase_contigs_resident", "fb_phase_contigs",
    "fb_score_reads", "fb_hap_block_from_partition", "fb_get_mec_stats_epsilon", "fb_beam_search_phasing",
    "fb_optimize_clustering", "fb_process_reads_for_final_parts", "fb_free_parts", "fb_get_hapq",
    "fb_update_hap_graph", "fb_process_reads_for_final_parts_resident", "fb_get_hapq_resident",
    "fb_update_hap_graph_resident",
]
READER_EXPORTS = ["fb_reader_options_default", "fb_read_frags", "fb_free_frag_set", "fb_reader_last_error"]  # floria_b200_reader.h

_lib = None


class FloriaB200Error(RuntimeError):
    pass


def load_library():
    """dlopen the CUDA library; raises if it has not been built (python -m floria_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FloriaB200Error(
            f"{LIB_PATH} is missing: build it with `python -m floria_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    L.fb_last_error.restype = C.c_char_p
    L.fb_last_error.argtypes = [C.c_void_p]
    L.fb_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.fb_destroy.argtypes = [C.c_void_p]
    L.fb_stream.restype = C.c_void_p
    L.fb_stream.argtypes = [C.c_void_p]
    L.fb_last_timings.argtypes = [C.c_void_p, C.POINTER(FbTimings)]
    L.fb_reader_options_default.argtypes = [C.POINTER(FbReaderOptions)]
    L.fb_read_frags.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(FbReaderOptions),
                                C.POINTER(C.POINTER(FbFragSet))]
    L.fb_free_frag_set.argtypes = [C.POINTER(FbFragSet)]
    L.fb_reader_last_error.restype = C.c_char_p
    L.fb_frags_upload.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.POINTER(C.c_void_p)]
    L.fb_frags_free.argtypes = [C.c_void_p, C.c_void_p]
    L.fb_dfrags_bytes.restype = C.c_uint64
    L.fb_dfrags_bytes.argtypes = [C.c_void_p]
    L.fb_get_range_with_lengths.restype = C.c_int64
    L.fb_get_range_with_lengths.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, u32p, u32p,
                                            C.c_uint64]
    L.fb_find_reads_in_interval.restype = C.c_int64
    L.fb_find_reads_in_interval.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, u32p, u32p, u32p, C.c_uint64]
    L.fb_phase_blocks.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u32p, u32p, C.POINTER(FbParams),
                                  C.POINTER(C.POINTER(FbBlockResults))]
    L.fb_phase_blocks_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, u32p, u32p, C.POINTER(FbParams),
                                           C.POINTER(C.POINTER(FbBlockResults))]
    L.fb_free_block_results.argtypes = [C.POINTER(FbBlockResults)]
    L.fb_phase_block.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u32p, C.c_uint32, C.POINTER(FbParams), u8p,
                                 f64p, f64p, C.POINTER(FbBlockPhase)]
    L.fb_phase_block_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, u32p, C.c_uint32, C.POINTER(FbParams),
                                          u8p, f64p, f64p, C.POINTER(FbBlockPhase)]
    L.fb_score_reads.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u32p, u8p, C.c_uint32,
                                 C.POINTER(FbParams), f64p, f64p, i64p, i64p, u32p]
    L.fb_hap_block_from_partition.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u32p, u8p, C.c_uint32,
                                              C.c_int, C.POINTER(FbParams), C.c_uint32, C.c_uint32, f64p, u8p]
    L.fb_get_mec_stats_epsilon.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u32p, u8p, C.c_uint32,
                                           C.c_int, C.POINTER(FbParams), f64p, f64p]
    L.fb_beam_search_phasing.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u32p, C.c_uint32,
                                         C.POINTER(FbParams), u8p, f64p, f64p, f64p, f64p, C.c_uint64, u64p]
    L.fb_optimize_clustering.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u32p, u8p, C.c_uint32,
                                         C.POINTER(FbParams), u8p, f64p, u32p]
    L.fb_process_reads_for_final_parts.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u64p, u32p, u32p,
                                                   u32p, C.POINTER(FbParams), C.POINTER(C.POINTER(FbParts))]
    L.fb_free_parts.argtypes = [C.POINTER(FbParts)]
    L.fb_get_hapq.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u64p, u32p, u32p, u32p, u64p, C.c_uint64,
                              C.POINTER(FbParams), u8p, f64p, f64p]
    L.fb_update_hap_graph.argtypes = [C.c_void_p, C.POINTER(FbFrags), C.c_uint64, u64p, u64p, u32p, u32p, u32p,
                                      C.POINTER(FbParams), f64p]
    L.fb_process_reads_for_final_parts_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, u64p, u32p, u32p, u32p,
                                                            C.POINTER(FbParams), C.POINTER(C.POINTER(FbParts))]
    L.fb_get_hapq_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, u64p, u32p, u32p, u32p, u64p, C.c_uint64,
                                       C.POINTER(FbParams), u8p, f64p, f64p]
    L.fb_update_hap_graph_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, u64p, u64p, u32p, u32p, u32p,
                                               C.POINTER(FbParams), f64p]
    # several devices in one process
    i32p = C.POINTER(C.c_int)
    L.fb_init_multi.argtypes = [C.c_int, i32p, C.POINTER(C.c_void_p)]
    L.fb_destroy_multi.argtypes = [C.c_void_p]
    L.fb_multi_size.argtypes = [C.c_void_p]
    L.fb_multi_ctx.restype = C.c_void_p
    L.fb_multi_ctx.argtypes = [C.c_void_p, C.c_int]
    L.fb_multi_last_error.restype = C.c_char_p
    L.fb_multi_last_error.argtypes = [C.c_void_p]
    L.fb_lpt_assign.restype = None
    L.fb_lpt_assign.argtypes = [f64p, C.c_uint64, C.c_uint32, u32p]
    L.fb_contigs_upload.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(FbFrags), u64p, u32p, u32p, C.POINTER(C.c_void_p)]
    L.fb_contigs_free.argtypes = [C.c_void_p, C.c_void_p]
    L.fb_phase_contigs_resident.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(FbParams),
                                            C.POINTER(C.POINTER(FbBlockResults)), u32p, C.POINTER(C.c_float)]
    L.fb_phase_contigs.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(FbFrags), u64p, u32p, u32p, C.POINTER(FbParams),
                                   C.POINTER(C.POINTER(FbBlockResults)), u32p, C.POINTER(C.c_float)]
    # measurement helpers (include/floria_b200_bench.h)
    L.fb_bench_synth_dense.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_double,
                                       C.c_double, u8p, u8p, u8p, C.POINTER(C.c_void_p)]
    L.fb_bench_sweep_hist.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, u8p, C.POINTER(FbParams), C.c_uint32,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), u64p]
    L.fb_bench_block_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, u8p, C.POINTER(FbParams), u64p, u64p, i64p,
                                        i64p, u32p]
    L.fb_bench_export_csr.argtypes = [C.c_void_p, C.c_void_p, u64p, u32p, u32p, u32p, u8p, u8p]
    L.fb_dfrags_nnz.restype = C.c_uint64
    L.fb_dfrags_nnz.argtypes = [C.c_void_p]
    L.fb_dfrags_n_reads.restype = C.c_uint64
    L.fb_dfrags_n_reads.argtypes = [C.c_void_p]
    L.fb_bench_download_planes.argtypes = [C.c_void_p, C.c_void_p, u64p, u8p, u32p, C.POINTER(C.c_uint16)]
    _lib = L
    return L


def get_range_with_lengths(snp_to_genome_pos, block_length, overlap_len, minimal_density):
    """utils_frags.rs:405-463 (host-side; defines the work units)."""
    L = load_library()
    g = np.ascontiguousarray(snp_to_genome_pos, dtype=np.uint64)
    cap = len(g) + 1
    lo = np.zeros(cap, np.uint32)
    hi = np.zeros(cap, np.uint32)
    n = L.fb_get_range_with_lengths(ptr(g, u64p), len(g), block_length, overlap_len, minimal_density, ptr(lo, u32p),
                                    ptr(hi, u32p), cap)
    if n < 0:
        raise FloriaB200Error("VCF malformed. Positions are not increasing")
    return lo[:n].copy(), hi[:n].copy()


def find_reads_in_interval(start, end, frags):
    """local_clustering.rs:12-59."""
    L = load_library()
    out = np.zeros(max(frags.n_reads, 1), np.uint32)
    n = L.fb_find_reads_in_interval(start, end, frags.n_reads, ptr(frags.first, u32p), ptr(frags.last, u32p),
                                    ptr(out, u32p), len(out))
    return out[:n].copy()


def read_frags(bam_path, vcf_path, contig=None, mapq_cutoff=15, use_supp_aln=True, supp_aln_dist_cutoff=40000):
    """fb_read_frags (include/floria_b200_reader.h): BAM + VCF -> (Frags in Frag::cmp order, snp_to_genome_pos, info).
    Host-only; restates get_vcf_profile / alignment_passed_check / frag_from_record / combine_frags of file_reader.rs."""
    from .frags import Frags

    L = load_library()
    o = FbReaderOptions()
    L.fb_reader_options_default(C.byref(o))
    o.mapq_cutoff, o.use_supp_aln, o.supp_aln_dist_cutoff = int(mapq_cutoff), int(bool(use_supp_aln)), int(supp_aln_dist_cutoff)
    out = C.POINTER(FbFragSet)()
    rc = L.fb_read_frags(os.fsencode(bam_path), os.fsencode(vcf_path), contig.encode() if contig else None, C.byref(o),
                         C.byref(out))
    if rc != 0:
        raise FloriaB200Error(f"fb_read_frags rc={rc}: {L.fb_reader_last_error().decode()}")
    try:
        s = out.contents
        f = s.frags
        n, nnz = int(f.n_reads), int(f.nnz)
        take = lambda p, cnt, dt: np.ctypeslib.as_array(p, shape=(max(cnt, 1),))[:cnt].astype(dt, copy=True)
        fr = Frags(take(f.row_ptr, n + 1, np.uint64), take(f.pos, nnz, np.uint32), take(f.allele, nnz, np.uint8),
                   take(f.qual, nnz, np.uint8), take(f.first, n, np.uint32), take(f.last, n, np.uint32))
        g2p = take(s.snp_to_genome_pos, int(s.n_snps), np.uint64)
        info = {"contig": s.contig.decode(), "n_records": int(s.n_records), "n_passed": int(s.n_passed),
                "n_without_snps": int(s.n_without_snps), "read_len_p66": int(s.read_len_p66)}
    finally:
        L.fb_free_frag_set(out)
    return fr, g2p, info


def contig_cost(frags, n_blocks):
    """the cost estimate fb_contigs_upload deals contigs to devices by (stored cells, block count as tie-breaker)"""
    return float(frags.nnz) + 1e-3 * float(n_blocks)


def lpt_assign(costs, n_bins):
    """fb_lpt_assign: deterministic longest-processing-time-first assignment (the library's own, so that ranks of a
    process-per-GPU job agree with it)"""
    L = load_library()
    c = np.ascontiguousarray(costs, dtype=np.float64)
    owner = np.zeros(max(len(c), 1), np.uint32)
    L.fb_lpt_assign(ptr(c, f64p), len(c), n_bins, ptr(owner, u32p))
    return owner[: len(c)].astype(np.int64)


class DeviceContigs:
    def __init__(self, multi, handle, n_contigs):
        self.multi, self.handle, self.n_contigs = multi, handle, n_contigs

    def free(self):
        if self.handle:
            self.multi.L.fb_contigs_free(self.multi.h, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MultiContext:
    """fb_init_multi: several devices driven from this process (one host thread + stream per device inside the
    library).  phase_contigs deals a contig list to the devices (static LPT queue) and returns per-contig results."""

    def __init__(self, device_ids):
        self.L = load_library()
        ids = (C.c_int * len(device_ids))(*device_ids)
        h = C.c_void_p()
        rc = self.L.fb_init_multi(len(device_ids), ids, C.byref(h))
        if rc != 0:
            raise FloriaB200Error(f"fb_init_multi failed ({rc}): " + self.L.fb_last_error(None).decode())
        self.h = h
        self.n_devices = len(device_ids)

    def close(self):
        if getattr(self, "h", None):
            self.L.fb_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def timings(self, i=0):
        """fb_last_timings of device i's context (launch counters, per-kernel times of the last call)"""
        t = FbTimings()
        self.L.fb_last_timings(C.c_void_p(self.L.fb_multi_ctx(self.h, i)), C.byref(t))
        return {k: getattr(t, k) for k, _ in FbTimings._fields_}

    def _chk(self, rc):
        if rc != 0:
            raise FloriaB200Error(f"floria_b200 error {rc}: " + self.L.fb_multi_last_error(self.h).decode())

    @staticmethod
    def _pack(contigs, blocks):
        n = len(contigs)
        arr = (FbFrags * max(n, 1))()
        keep = []
        for k, fr in enumerate(contigs):
            st = fr.as_struct()
            keep.append(st)
            arr[k] = st
        bp = np.zeros(n + 1, np.uint64)
        for k, (lo, _) in enumerate(blocks):
            bp[k + 1] = bp[k] + len(lo)
        lo = np.concatenate([np.asarray(b[0], np.uint32) for b in blocks]) if n else np.zeros(0, np.uint32)
        hi = np.concatenate([np.asarray(b[1], np.uint32) for b in blocks]) if n else np.zeros(0, np.uint32)
        return arr, bp, np.ascontiguousarray(lo), np.ascontiguousarray(hi), keep

    def upload(self, contigs, blocks):
        arr, bp, lo, hi, keep = self._pack(contigs, blocks)
        out = C.c_void_p()
        self._chk(self.L.fb_contigs_upload(self.h, len(contigs), arr, ptr(bp, u64p), ptr(lo, u32p), ptr(hi, u32p),
                                           C.byref(out)))
        return DeviceContigs(self, out, len(contigs))

    def _collect(self, n, outs, dev, ms):
        res = []
        for k in range(n):
            res.append(BlockResults(outs[k].contents))
            self.L.fb_free_block_results(outs[k])
        return res, dev[:n].astype(np.int64), ms.copy()

    def phase_contigs_resident(self, dcontigs, params):
        n = dcontigs.n_contigs
        outs = (C.POINTER(FbBlockResults) * max(n, 1))()
        dev = np.zeros(max(n, 1), np.uint32)
        ms = np.zeros(self.n_devices, np.float32)
        self._chk(self.L.fb_phase_contigs_resident(self.h, dcontigs.handle, C.byref(params), outs, ptr(dev, u32p),
                                                   ms.ctypes.data_as(C.POINTER(C.c_float))))
        return self._collect(n, outs, dev, ms)

    def phase_contigs(self, contigs, blocks, params):
        """contigs: list of Frags; blocks: list of (blk_lo, blk_hi) per contig.  Returns (list of BlockResults in contig
        order, device index per contig, CUDA-event ms per device)."""
        arr, bp, lo, hi, keep = self._pack(contigs, blocks)
        n = len(contigs)
        outs = (C.POINTER(FbBlockResults) * max(n, 1))()
        dev = np.zeros(max(n, 1), np.uint32)
        ms = np.zeros(self.n_devices, np.float32)
        self._chk(self.L.fb_phase_contigs(self.h, n, arr, ptr(bp, u64p), ptr(lo, u32p), ptr(hi, u32p), C.byref(params),
                                          outs, ptr(dev, u32p), ms.ctypes.data_as(C.POINTER(C.c_float))))
        return self._collect(n, outs, dev, ms)


class DeviceFrags:
    def __init__(self, ctx, handle, frags):
        self.ctx = ctx
        self.handle = handle
        self.frags = frags
        self.n_reads = frags.n_reads if frags is not None else None

    @property
    def nbytes(self):
        return int(self.ctx.L.fb_dfrags_bytes(self.handle))

    def free(self):
        if self.handle:
            self.ctx.L.fb_frags_free(self.ctx.h, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.fb_init(device, C.byref(h))
        if rc != 0:
            raise FloriaB200Error(f"fb_init failed ({rc}): " + self.L.fb_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.fb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise FloriaB200Error(f"floria_b200 error {rc}: " + self.L.fb_last_error(self.h).decode())

    @property
    def stream(self):
        return self.L.fb_stream(self.h)

    def timings(self):
        t = FbTimings()
        self._chk(self.L.fb_last_timings(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in FbTimings._fields_}

    # ---- data movement ----
    def upload(self, frags):
        fs = frags.as_struct()
        out = C.c_void_p()
        self._chk(self.L.fb_frags_upload(self.h, C.byref(fs), C.byref(out)))
        return DeviceFrags(self, out, frags)

    # ---- measurement helpers (include/floria_b200_bench.h) ----
    def bench_synth_dense(self, n_reads, n_snps, ploidy, config_seed, present=0.98, flip=0.04):
        """configs[2]-style full-span block generated in HBM; same cells as synth.make_contig(.., full_span=True)."""
        from . import synth

        seed = synth.SEED_BASE + config_seed
        truth, nall = synth.make_truth(seed, ploidy, n_snps)
        rid = np.arange(n_reads, dtype=np.uint64)
        w = 1.0 / np.arange(1, ploidy + 1)
        cdf = np.cumsum(w / w.sum())
        src = np.searchsorted(cdf, synth.rng_u01(seed, 100, rid), side="right").clip(0, ploidy - 1).astype(np.uint8)
        truth = np.ascontiguousarray(truth, np.uint8)
        nall8 = np.ascontiguousarray(nall, np.uint8)
        out = C.c_void_p()
        self._chk(self.L.fb_bench_synth_dense(self.h, n_reads, n_snps, ploidy, seed, present, flip, ptr(truth, u8p),
                                              ptr(nall8, u8p), ptr(src, u8p), C.byref(out)))
        d = DeviceFrags(self, out, None)
        d.src = src
        d.n_reads = n_reads
        return d

    def bench_export_csr(self, dfrags, alloc=None):
        """host CSR (a Frags) of a resident contig; alloc(nbytes) -> uint8 numpy array lets the caller supply pinned memory"""
        from .frags import Frags

        alloc = alloc or (lambda n: np.zeros(max(n, 1), np.uint8))
        R = int(self.L.fb_dfrags_n_reads(dfrags.handle))
        nnz = int(self.L.fb_dfrags_nnz(dfrags.handle))
        row_ptr = alloc(8 * (R + 1))[: 8 * (R + 1)].view(np.uint64)
        first = alloc(4 * R)[: 4 * R].view(np.uint32)
        last = alloc(4 * R)[: 4 * R].view(np.uint32)
        pos = alloc(4 * nnz)[: 4 * nnz].view(np.uint32)
        allele = alloc(nnz)[:nnz]
        qual = alloc(nnz)[:nnz]
        self._chk(self.L.fb_bench_export_csr(self.h, dfrags.handle, ptr(row_ptr, u64p), ptr(first, u32p),
                                             ptr(last, u32p), ptr(pos, u32p), ptr(allele, u8p), ptr(qual, u8p)))
        f = Frags.__new__(Frags)
        f.row_ptr, f.pos, f.allele, f.qual, f.first, f.last = row_ptr, pos, allele, qual, first, last
        f.n_reads, f.nnz = R, nnz
        return f

    def bench_sweep_hist(self, dfrags, ploidy, hap, params, iters):
        hap = np.ascontiguousarray(hap, np.uint8)
        sw = np.zeros(iters, np.float32)
        hs = np.zeros(iters, np.float32)
        cells = C.c_uint64(0)
        self._chk(self.L.fb_bench_sweep_hist(self.h, dfrags.handle, ploidy, ptr(hap, u8p), C.byref(params), iters,
                                             sw.ctypes.data_as(C.POINTER(C.c_float)),
                                             hs.ctypes.data_as(C.POINTER(C.c_float)), C.byref(cells)))
        return sw, hs, int(cells.value)

    def bench_block_tables(self, dfrags, ploidy, hap, params):
        """raw k_hist count words [ploidy, n_pos, 4] and SCORE-mode sweep sums (same_q26, diff_q26, n_empty) of one block
        made of all reads of a resident contig (full-size property tests)."""
        hap = np.ascontiguousarray(hap, np.uint8)
        npos = C.c_uint64(0)
        self._chk(self.L.fb_bench_block_tables(self.h, dfrags.handle, ploidy, ptr(hap, u8p), C.byref(params),
                                               C.byref(npos), None, None, None, None))
        n_pos, n = int(npos.value), len(hap)
        counts = np.zeros((ploidy, n_pos, 4), np.uint64)
        sq = np.zeros((n, ploidy), np.int64)
        dq = np.zeros((n, ploidy), np.int64)
        ne = np.zeros((n, ploidy), np.uint32)
        self._chk(self.L.fb_bench_block_tables(self.h, dfrags.handle, ploidy, ptr(hap, u8p), C.byref(params),
                                               C.byref(npos), ptr(counts, u64p), ptr(sq, i64p), ptr(dq, i64p),
                                               ptr(ne, u32p)))
        return counts, sq, dq, ne

    def download_planes(self, dfrags):
        ng = C.c_uint64(0)
        self._chk(self.L.fb_bench_download_planes(self.h, dfrags.handle, C.byref(ng), None, None, None))
        n = int(ng.value)
        q = np.zeros(n * 16, np.uint8)
        a = np.zeros(n, np.uint32)
        p = np.zeros(n, np.uint16)
        self._chk(self.L.fb_bench_download_planes(self.h, dfrags.handle, C.byref(ng), ptr(q, u8p), ptr(a, u32p),
                                                  p.ctypes.data_as(C.POINTER(C.c_uint16))))
        return q.reshape(n, 16), a, p

    # ---- batched hot path ----
    def phase_blocks(self, frags, blk_lo, blk_hi, params):
        lo = np.ascontiguousarray(blk_lo, dtype=np.uint32)
        hi = np.ascontiguousarray(blk_hi, dtype=np.uint32)
        out = C.POINTER(FbBlockResults)()
        fs = frags.as_struct()
        self._chk(self.L.fb_phase_blocks(self.h, C.byref(fs), len(lo), ptr(lo, u32p), ptr(hi, u32p), C.byref(params),
                                         C.byref(out)))
        res = BlockResults(out.contents)
        self.L.fb_free_block_results(out)
        return res

    def phase_blocks_resident(self, dfrags, blk_lo, blk_hi, params):
        lo = np.ascontiguousarray(blk_lo, dtype=np.uint32)
        hi = np.ascontiguousarray(blk_hi, dtype=np.uint32)
        out = C.POINTER(FbBlockResults)()
        self._chk(self.L.fb_phase_blocks_resident(self.h, dfrags.handle, len(lo), ptr(lo, u32p), ptr(hi, u32p),
                                                  C.byref(params), C.byref(out)))
        res = BlockResults(out.contents)
        self.L.fb_free_block_results(out)
        return res

    # ---- fine-grained entry points ----
    def phase_block(self, frags, sel, ploidy, params):
        """graph_processing.rs:140-162 for one block at a fixed ploidy: beam_search_phasing -> optimize_clustering ->
        get_mec_stats_epsilon_no_phred.  `frags` = host Frags (uploaded and packed by the call) or DeviceFrags;
        sel=None takes every read.  Returns (hap, mec_bases, mec_errors, info dict)."""
        resident = isinstance(frags, DeviceFrags)
        if sel is None:
            n, selp = int(frags.n_reads), None
        else:
            sel = np.ascontiguousarray(sel, dtype=np.uint32)
            n, selp = len(sel), ptr(sel, u32p)
        hap = np.zeros(max(n, 1), np.uint8)
        bases = np.zeros(ploidy)
        errors = np.zeros(ploidy)
        info = FbBlockPhase()
        if resident:
            self._chk(self.L.fb_phase_block_resident(self.h, frags.handle, n, selp, ploidy, C.byref(params),
                                                     ptr(hap, u8p), ptr(bases, f64p), ptr(errors, f64p), C.byref(info)))
        else:
            fs = frags.as_struct()
            self._chk(self.L.fb_phase_block(self.h, C.byref(fs), n, selp, ploidy, C.byref(params), ptr(hap, u8p),
                                            ptr(bases, f64p), ptr(errors, f64p), C.byref(info)))
        return hap[:n], bases, errors, {k: getattr(info, k) for k, _ in FbBlockPhase._fields_}

    def score_reads(self, frags, sel, hap, ploidy, params):
        sel = np.ascontiguousarray(sel, dtype=np.uint32)
        hap = np.ascontiguousarray(hap, dtype=np.uint8)
        n = len(sel)
        same = np.zeros((n, ploidy))
        diff = np.zeros((n, ploidy))
        sq = np.zeros((n, ploidy), np.int64)
        dq = np.zeros((n, ploidy), np.int64)
        ne = np.zeros((n, ploidy), np.uint32)
        fs = frags.as_struct()
        self._chk(self.L.fb_score_reads(self.h, C.byref(fs), n, ptr(sel, u32p), ptr(hap, u8p), ploidy,
                                        C.byref(params), ptr(same, f64p), ptr(diff, f64p), ptr(sq, i64p),
                                        ptr(dq, i64p), ptr(ne, u32p)))
        return same, diff, sq, dq, ne

    def hap_block_from_partition(self, frags, sel, hap, ploidy, use_qual, params, pos_lo, n_pos):
        sel = np.ascontiguousarray(sel, dtype=np.uint32)
        hap = np.ascontiguousarray(hap, dtype=np.uint8)
        counts = np.zeros((ploidy, n_pos, 4))
        mask = np.zeros((ploidy, n_pos), np.uint8)
        fs = frags.as_struct()
        self._chk(self.L.fb_hap_block_from_partition(self.h, C.byref(fs), len(sel), ptr(sel, u32p), ptr(hap, u8p),
                                                     ploidy, int(use_qual), C.byref(params), pos_lo, n_pos,
                                                     ptr(counts, f64p), ptr(mask, u8p)))
        return counts, mask

    def get_mec_stats_epsilon(self, frags, sel, hap, ploidy, use_phred, params):
        sel = np.ascontiguousarray(sel, dtype=np.uint32)
        hap = np.ascontiguousarray(hap, dtype=np.uint8)
        bases = np.zeros(ploidy)
        errors = np.zeros(ploidy)
        fs = frags.as_struct()
        self._chk(self.L.fb_get_mec_stats_epsilon(self.h, C.byref(fs), len(sel), ptr(sel, u32p), ptr(hap, u8p),
                                                  ploidy, int(use_phred), C.byref(params), ptr(bases, f64p),
                                                  ptr(errors, f64p)))
        return bases, errors

    def beam_search_phasing(self, frags, sel, ploidy, params, tap_cap=0):
        sel = np.ascontiguousarray(sel, dtype=np.uint32)
        hap = np.zeros(max(len(sel), 1), np.uint8)
        score = C.c_double(0)
        tn = C.c_uint64(0)
        ts = np.zeros(max(tap_cap, 1))
        td = np.zeros(max(tap_cap, 1))
        tp = np.zeros(max(tap_cap, 1))
        fs = frags.as_struct()
        self._chk(self.L.fb_beam_search_phasing(self.h, C.byref(fs), len(sel), ptr(sel, u32p), ploidy,
                                                C.byref(params), ptr(hap, u8p), C.byref(score), ptr(ts, f64p),
                                                ptr(td, f64p), ptr(tp, f64p), tap_cap, C.byref(tn)))
        n = min(int(tn.value), tap_cap)
        return hap[: len(sel)], score.value, (ts[:n], td[:n], tp[:n], int(tn.value))

    def optimize_clustering(self, frags, sel, hap_in, ploidy, params):
        sel = np.ascontiguousarray(sel, dtype=np.uint32)
        hap_in = np.ascontiguousarray(hap_in, dtype=np.uint8)
        hap = np.zeros(max(len(sel), 1), np.uint8)
        score = C.c_double(0)
        nr = C.c_uint32(0)
        fs = frags.as_struct()
        self._chk(self.L.fb_optimize_clustering(self.h, C.byref(fs), len(sel), ptr(sel, u32p), ptr(hap_in, u8p),
                                                ploidy, C.byref(params), ptr(hap, u8p), C.byref(score),
                                                C.byref(nr)))
        return hap[: len(sel)], score.value, int(nr.value)

    def process_reads_for_final_parts(self, frags, part_ptr, part_reads, range_lo, range_hi, params):
        pp = np.ascontiguousarray(part_ptr, np.uint64)
        pr = np.ascontiguousarray(part_reads, np.uint32)
        rl = np.ascontiguousarray(range_lo, np.uint32)
        rh = np.ascontiguousarray(range_hi, np.uint32)
        out = C.POINTER(FbParts)()
        if isinstance(frags, DeviceFrags):  # contig already resident in HBM
            self._chk(self.L.fb_process_reads_for_final_parts_resident(self.h, frags.handle, len(pp) - 1, ptr(pp, u64p),
                                                                       ptr(pr, u32p), ptr(rl, u32p), ptr(rh, u32p),
                                                                       C.byref(params), C.byref(out)))
        else:
            fs = frags.as_struct()
            self._chk(self.L.fb_process_reads_for_final_parts(self.h, C.byref(fs), len(pp) - 1, ptr(pp, u64p),
                                                              ptr(pr, u32p), ptr(rl, u32p), ptr(rh, u32p),
                                                              C.byref(params), C.byref(out)))
        res = Parts(out.contents)
        self.L.fb_free_parts(out)
        return res

    def get_hapq(self, frags, part_ptr, part_reads, range_lo, range_hi, snp_to_genome_pos, params):
        pp = np.ascontiguousarray(part_ptr, np.uint64)
        pr = np.ascontiguousarray(part_reads, np.uint32)
        rl = np.ascontiguousarray(range_lo, np.uint32)
        rh = np.ascontiguousarray(range_hi, np.uint32)
        g = np.ascontiguousarray(snp_to_genome_pos, np.uint64)
        n = len(pp) - 1
        hapq = np.zeros(max(n, 1), np.uint8)
        rel = np.zeros(max(n, 1))
        avg = C.c_double(0)
        if isinstance(frags, DeviceFrags):
            self._chk(self.L.fb_get_hapq_resident(self.h, frags.handle, n, ptr(pp, u64p), ptr(pr, u32p), ptr(rl, u32p),
                                                  ptr(rh, u32p), ptr(g, u64p), len(g), C.byref(params), ptr(hapq, u8p),
                                                  ptr(rel, f64p), C.byref(avg)))
        else:
            fs = frags.as_struct()
            self._chk(self.L.fb_get_hapq(self.h, C.byref(fs), n, ptr(pp, u64p), ptr(pr, u32p), ptr(rl, u32p),
                                         ptr(rh, u32p), ptr(g, u64p), len(g), C.byref(params), ptr(hapq, u8p),
                                         ptr(rel, f64p), C.byref(avg)))
        return hapq[:n], rel[:n], avg.value

    def update_hap_graph(self, frags, col_ptr, node_ptr, node_reads, node_lo, node_hi, params):
        cp = np.ascontiguousarray(col_ptr, np.uint64)
        npt = np.ascontiguousarray(node_ptr, np.uint64)
        nr = np.ascontiguousarray(node_reads, np.uint32)
        nl = np.ascontiguousarray(node_lo, np.uint32)
        nh = np.ascontiguousarray(node_hi, np.uint32)
        n_cols = len(cp) - 1
        tot = sum(int(cp[i + 1] - cp[i]) * int(cp[i + 2] - cp[i + 1]) for i in range(n_cols - 1))
        out = np.zeros(max(tot, 1))
        if isinstance(frags, DeviceFrags):
            self._chk(self.L.fb_update_hap_graph_resident(self.h, frags.handle, n_cols, ptr(cp, u64p), ptr(npt, u64p),
                                                          ptr(nr, u32p), ptr(nl, u32p), ptr(nh, u32p), C.byref(params),
                                                          ptr(out, f64p)))
        else:
            fs = frags.as_struct()
            self._chk(self.L.fb_update_hap_graph(self.h, C.byref(fs), n_cols, ptr(cp, u64p), ptr(npt, u64p),
                                                 ptr(nr, u32p), ptr(nl, u32p), ptr(nh, u32p), C.byref(params),
                                                 ptr(out, f64p)))
        return out[:tot]
