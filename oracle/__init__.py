"""ctypes binding of the CPU ORACLE (oracle/floria_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; nothing under floria_b200/ does.  PARITY UNPINNED: see floria_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from floria_b200._cdefs import (
    BlockResults,
    FbBlockPhase,
    FbBlockResults,
    FbFrags,
    FbParams,
    FbParts,
    Parts,
    default_params,
    f32p,
    f64p,
    ptr,
    u8p,
    u32p,
    u64p,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "floria_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_stable_binom_cdf_p_rev.restype = C.c_double
        L.orc_stable_binom_cdf_p_rev.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double]
        L.orc_log_sum_exp.restype = C.c_double
        L.orc_log_sum_exp.argtypes = [f64p, C.c_uint64]
        L.orc_mec_threshold.restype = C.c_double
        L.orc_mec_threshold.argtypes = [C.c_uint32, C.c_double, C.c_uint32]
        L.orc_get_range_with_lengths.restype = C.c_int64
        L.orc_get_range_with_lengths.argtypes = [u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, u32p, u32p,
                                                 C.c_uint64]
        L.orc_find_reads_in_interval.restype = C.c_int64
        L.orc_find_reads_in_interval.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, u32p, u32p, u32p, C.c_uint64]
        L.orc_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())


def phred_lut():
    out = np.zeros(256, dtype=np.float32)
    lib().orc_phred_lut(ptr(out, f32p))
    return out


def stable_binom_cdf_p_rev(n, k, p, div):
    return lib().orc_stable_binom_cdf_p_rev(n, k, p, div)


def log_sum_exp(xs):
    a = np.ascontiguousarray(xs, dtype=np.float64)
    return lib().orc_log_sum_exp(ptr(a, f64p), len(a))


def mec_threshold(ploidy, eps, sens):
    return lib().orc_mec_threshold(ploidy, eps, sens)


def get_range_with_lengths(snp_to_genome_pos, block_length, overlap_len, minimal_density):
    g = np.ascontiguousarray(snp_to_genome_pos, dtype=np.uint64)
    cap = len(g) + 1
    lo = np.zeros(cap, np.uint32)
    hi = np.zeros(cap, np.uint32)
    n = lib().orc_get_range_with_lengths(ptr(g, u64p), len(g), block_length, overlap_len, minimal_density,
                                         ptr(lo, u32p), ptr(hi, u32p), cap)
    if n < 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())
    return lo[:n].copy(), hi[:n].copy()


def find_reads_in_interval(start, end, frags):
    out = np.zeros(max(frags.n_reads, 1), np.uint32)
    n = lib().orc_find_reads_in_interval(start, end, frags.n_reads, ptr(frags.first, u32p), ptr(frags.last, u32p),
                                         ptr(out, u32p), len(out))
    return out[:n].copy()


def _sel(sel):
    return np.ascontiguousarray(sel, dtype=np.uint32)


def score_reads(frags, sel, hap, ploidy, params):
    sel = _sel(sel)
    hap = np.ascontiguousarray(hap, dtype=np.uint8)
    same = np.zeros((len(sel), ploidy))
    diff = np.zeros((len(sel), ploidy))
    fs = frags.as_struct()
    _chk(lib().orc_score_reads(C.byref(fs), len(sel), ptr(sel, u32p), ptr(hap, u8p), ploidy, C.byref(params),
                               ptr(same, f64p), ptr(diff, f64p)))
    return same, diff


def score_reads_noeps(frags, sel, hap, ploidy, params):
    sel = _sel(sel)
    hap = np.ascontiguousarray(hap, dtype=np.uint8)
    same = np.zeros((len(sel), ploidy), np.uint64)
    diff = np.zeros((len(sel), ploidy), np.uint64)
    fs = frags.as_struct()
    _chk(lib().orc_score_reads_noeps(C.byref(fs), len(sel), ptr(sel, u32p), ptr(hap, u8p), ploidy, C.byref(params),
                                     ptr(same, u64p), ptr(diff, u64p)))
    return same, diff


def hap_block_from_partition(frags, sel, hap, ploidy, use_qual, params, pos_lo, n_pos):
    sel = _sel(sel)
    hap = np.ascontiguousarray(hap, dtype=np.uint8)
    counts = np.zeros((ploidy, n_pos, 4))
    mask = np.zeros((ploidy, n_pos), np.uint8)
    fs = frags.as_struct()
    _chk(lib().orc_hap_block_from_partition(C.byref(fs), len(sel), ptr(sel, u32p), ptr(hap, u8p), ploidy,
                                            int(use_qual), C.byref(params), pos_lo, n_pos, ptr(counts, f64p),
                                            ptr(mask, u8p)))
    return counts, mask


def get_mec_stats_epsilon(frags, sel, hap, ploidy, use_phred, params):
    sel = _sel(sel)
    hap = np.ascontiguousarray(hap, dtype=np.uint8)
    bases = np.zeros(ploidy)
    errors = np.zeros(ploidy)
    fs = frags.as_struct()
    _chk(lib().orc_get_mec_stats_epsilon(C.byref(fs), len(sel), ptr(sel, u32p), ptr(hap, u8p), ploidy,
                                         int(use_phred), C.byref(params), ptr(bases, f64p), ptr(errors, f64p)))
    return bases, errors


def beam_search_phasing(frags, sel, ploidy, params, tap_cap=0):
    sel = _sel(sel)
    hap = np.zeros(len(sel), np.uint8)
    score = C.c_double(0)
    tn = C.c_uint64(0)
    ts = np.zeros(max(tap_cap, 1))
    td = np.zeros(max(tap_cap, 1))
    tp = np.zeros(max(tap_cap, 1))
    fs = frags.as_struct()
    _chk(lib().orc_beam_search_phasing(C.byref(fs), C.c_uint64(len(sel)), ptr(sel, u32p), C.c_uint32(ploidy),
                                       C.byref(params), ptr(hap, u8p), C.byref(score), ptr(ts, f64p), ptr(td, f64p),
                                       ptr(tp, f64p), C.c_uint64(tap_cap), C.byref(tn)))
    n = min(int(tn.value), tap_cap)
    return hap, score.value, (ts[:n], td[:n], tp[:n], int(tn.value))


def optimize_clustering(frags, sel, hap_in, ploidy, params):
    sel = _sel(sel)
    hap_in = np.ascontiguousarray(hap_in, dtype=np.uint8)
    hap = np.zeros(len(sel), np.uint8)
    score = C.c_double(0)
    nr = C.c_uint32(0)
    fs = frags.as_struct()
    _chk(lib().orc_optimize_clustering(C.byref(fs), C.c_uint64(len(sel)), ptr(sel, u32p), ptr(hap_in, u8p),
                                       C.c_uint32(ploidy), C.byref(params), ptr(hap, u8p), C.byref(score),
                                       C.byref(nr)))
    return hap, score.value, int(nr.value)


def phase_block(frags, sel, ploidy, params):
    """mirror of api.Context.phase_block: beam -> optimize -> no-phred MEC of one block at a fixed ploidy"""
    if sel is None:
        n, selp = frags.n_reads, None
    else:
        sel = _sel(sel)
        n, selp = len(sel), ptr(sel, u32p)
    hap = np.zeros(max(n, 1), np.uint8)
    bases = np.zeros(ploidy)
    errors = np.zeros(ploidy)
    info = FbBlockPhase()
    fs = frags.as_struct()
    _chk(lib().orc_phase_block(C.byref(fs), C.c_uint64(n), selp, C.c_uint32(ploidy), C.byref(params), ptr(hap, u8p),
                               ptr(bases, f64p), ptr(errors, f64p), C.byref(info)))
    return hap[:n], bases, errors, {k: getattr(info, k) for k, _ in FbBlockPhase._fields_}


def phase_blocks(frags, blk_lo, blk_hi, params, n_threads=1):
    lo = np.ascontiguousarray(blk_lo, dtype=np.uint32)
    hi = np.ascontiguousarray(blk_hi, dtype=np.uint32)
    out = C.POINTER(FbBlockResults)()
    fs = frags.as_struct()
    _chk(lib().orc_phase_blocks(C.byref(fs), C.c_uint64(len(lo)), ptr(lo, u32p), ptr(hi, u32p), C.byref(params),
                                C.c_uint32(n_threads), C.byref(out)))
    res = BlockResults(out.contents)
    lib().orc_free_block_results(out)
    return res


def _parts_args(part_ptr, part_reads, range_lo, range_hi):
    return (np.ascontiguousarray(part_ptr, np.uint64), np.ascontiguousarray(part_reads, np.uint32),
            np.ascontiguousarray(range_lo, np.uint32), np.ascontiguousarray(range_hi, np.uint32))


def process_reads_for_final_parts(frags, part_ptr, part_reads, range_lo, range_hi, params):
    pp, pr, rl, rh = _parts_args(part_ptr, part_reads, range_lo, range_hi)
    out = C.POINTER(FbParts)()
    fs = frags.as_struct()
    _chk(lib().orc_process_reads_for_final_parts(C.byref(fs), C.c_uint64(len(pp) - 1), ptr(pp, u64p), ptr(pr, u32p),
                                                 ptr(rl, u32p), ptr(rh, u32p), C.byref(params), C.byref(out)))
    res = Parts(out.contents)
    lib().orc_free_parts(out)
    return res


def get_hapq(frags, part_ptr, part_reads, range_lo, range_hi, snp_to_genome_pos, params):
    pp, pr, rl, rh = _parts_args(part_ptr, part_reads, range_lo, range_hi)
    g = np.ascontiguousarray(snp_to_genome_pos, np.uint64)
    n = len(pp) - 1
    hapq = np.zeros(max(n, 1), np.uint8)
    rel = np.zeros(max(n, 1))
    avg = C.c_double(0)
    fs = frags.as_struct()
    _chk(lib().orc_get_hapq(C.byref(fs), C.c_uint64(n), ptr(pp, u64p), ptr(pr, u32p), ptr(rl, u32p), ptr(rh, u32p),
                            ptr(g, u64p), C.c_uint64(len(g)), C.byref(params), ptr(hapq, u8p), ptr(rel, f64p),
                            C.byref(avg)))
    return hapq[:n], rel[:n], avg.value


def update_hap_graph(frags, col_ptr, node_ptr, node_reads, node_lo, node_hi, params):
    cp = np.ascontiguousarray(col_ptr, np.uint64)
    npt = np.ascontiguousarray(node_ptr, np.uint64)
    nr = np.ascontiguousarray(node_reads, np.uint32)
    nl = np.ascontiguousarray(node_lo, np.uint32)
    nh = np.ascontiguousarray(node_hi, np.uint32)
    n_cols = len(cp) - 1
    tot = 0
    for i in range(n_cols - 1):
        tot += int(cp[i + 1] - cp[i]) * int(cp[i + 2] - cp[i + 1])
    out = np.zeros(max(tot, 1))
    fs = frags.as_struct()
    _chk(lib().orc_update_hap_graph(C.byref(fs), C.c_uint64(n_cols), ptr(cp, u64p), ptr(npt, u64p), ptr(nr, u32p),
                                    ptr(nl, u32p), ptr(nh, u32p), C.byref(params), ptr(out, f64p)))
    return out[:tot]
