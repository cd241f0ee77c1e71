"""Pins the CPU oracle against the only fixed points that exist for this path.

The reference ships no tests / golden vectors, so "parity unpinned": what can be pinned are the
hand-derived known answers of SURVEY.md Appendix D (formulas restated from utils_frags.rs:211-258, 702-711 and
graph_processing.rs:205-222) and the behavioural invariants derivable from the cited code.
"""
import math

import numpy as np
import pytest

import oracle
from floria_b200 import default_params, synth
from floria_b200.frags import Frags


def test_stable_binom_cdf_p_rev_known_answers():
    # utils_frags.rs:211-248; values from SURVEY.md Appendix D (hex = exact f64)
    cases = [
        (0, 0, 0.04, "0x0.0p+0"),
        (10, 0, 0.04, "0x1.a200c148c9140p+0"),
        (10, 1, 0.04, "-0x1.577ea491dbe76p+0"),
        (10, 2, 0.04, "-0x1.c2a33ddf683f2p+2"),
        (10, 10, 0.04, "-0x1.0182890b2d21bp+7"),
        (100, 3, 0.04, "0x1.231c7584bb9c4p-1"),
        (57, 9, 0.03125, "-0x1.f6acfbdeb024ap+4"),
        (1, 1, 0.01, "-0x1.26bb129fe4ff1p+4"),
    ]
    for n, k, eps, hx in cases:
        got = oracle.stable_binom_cdf_p_rev(n, k, eps, 0.25)
        assert got == float.fromhex(hx), (n, k, eps, got.hex(), hx)
    z = oracle.stable_binom_cdf_p_rev(100, 4, 0.04, 0.25)  # a == eps -> KL = 0 -> negative zero
    assert z == 0.0 and math.copysign(1.0, z) == -1.0


def test_log_sum_exp_known_answer():
    ps = [oracle.stable_binom_cdf_p_rev(10, k, 0.04, 0.25) for k in (0, 2, 5)]
    lse = oracle.log_sum_exp(ps)
    assert lse == float.fromhex("0x1.a20bf56588024p+0")
    cutoff = math.log(0.01)
    assert [p - lse > cutoff for p in ps] == [True, False, False]  # global_clustering.rs:98


def test_phred_scale_table():
    # utils_frags.rs:702-711: (1f32 - 10f32.powf(q/-10)) as f64; every weight is a multiple of 2^-26
    lut = oracle.phred_lut().astype(np.float64)
    num = lut * 2.0 ** 26
    assert np.all(num == np.floor(num))
    known = {0: 0, 1: 13802400, 3: 33474760, 10: 60397976, 20: 66437776, 30: 67041756, 40: 67102152,
             60: 67108796, 75: 67108860, 76: 67108864, 255: 67108864}
    for q, v in known.items():
        assert int(num[q]) == v, (q, int(num[q]), v)
    # w(2): SURVEY Appendix D computed 24766036 with numpy's float32 power; the host libm powf (what Rust's
    # f32::powf calls on a gnu target) gives 24766032 -- a 1-ulp(f32) libm difference, recorded in DESIGN.md.
    assert int(num[2]) in (24766032, 24766036)
    assert np.all(np.diff(lut[:77]) >= 0) and np.all(lut[76:] == 1.0)


def test_mec_thresholds():
    # graph_processing.rs:205-222 at eps = 0.04
    exp = {1: [0.736570, 0.762553, 0.781250, 0.795763], 2: [0.729167, 0.801282, 0.846354, 0.877193],
           3: [0.781250, 0.833333, 0.868056, 0.892857]}
    for s, vals in exp.items():
        for p, v in zip((2, 3, 4, 5), vals):
            assert abs(oracle.mec_threshold(p, 0.04, s) - v) < 5e-7


def _tiny():
    # three reads over SNPs 1..4
    # already in Frag::cmp order (first asc, last desc)
    reads = [
        ([1, 2, 3, 4], [0, 0, 1, 1], [20, 20, 20, 20]),
        ([1, 2, 3], [0, 0, 1], [30, 30, 30]),
        ([2, 3, 4], [1, 1, 0], [10, 10, 10]),
    ]
    fr = Frags.from_reads(reads)
    assert list(fr.order) == [0, 1, 2]
    return fr


def test_distance_tie_counts_as_same_and_empty_adds_epsilon():
    fr = _tiny()
    prm = default_params(epsilon=0.03125)
    lut = oracle.phred_lut().astype(np.float64)
    # hap0 = {read1}, hap1 = {read2}; score read0 (4 cells, q20) against both
    same, diff = oracle.score_reads(fr, [0, 1, 2], [255, 0, 1], 2, prm)
    # vs hap0: positions 1,2,3 agree with read1 (w(20) each), position 4 is empty in hap0 -> +eps
    assert same[0, 0] == lut[20] + lut[20] + lut[20] and diff[0, 0] == 0.03125
    # vs hap1 (read2 covers 2,3,4 with alleles 1,1,0): pos1 empty (+eps), pos2 mismatch, pos3 match, pos4 mismatch
    assert same[0, 1] == lut[20] and diff[0, 1] == 0.03125 + lut[20] + lut[20]
    # a tie between the read's allele and the consensus counts as same (utils_frags.rs:63-69):
    fr2 = Frags.from_reads([([1], [0], [30]), ([1], [1], [30]), ([1], [1], [5]), ([1], [0], [5])])
    same, diff = oracle.score_reads(fr2, [0, 1, 2, 3], [0, 0, 255, 255], 1, prm)
    assert same[2, 0] == lut[5] and diff[2, 0] == 0.0
    assert same[3, 0] == lut[5] and diff[3, 0] == 0.0
    # ... in distance_read_haplo (utils_frags.rs:96-102) a tie with a DIFFERENT consensus allele counts as
    # neither; the consensus on a tie is the last maximum in (canonical = ascending allele) iteration order.
    s2, d2 = oracle.score_reads_noeps(fr2, [0, 1, 2, 3], [0, 0, 255, 255], 1, prm)
    assert s2[2, 0] == 1 and d2[2, 0] == 0  # allele 1 == consensus: same += w(5)=0.68 -> rounds to 1
    assert s2[3, 0] == 0 and d2[3, 0] == 0  # allele 0 ties with the consensus count: neither


def test_all_empty_haplotype_scores_L_epsilons_sequentially():
    # (same, diff) = (0, L sequential adds of eps) (utils_frags.rs:45-48)
    L = 25
    fr = Frags.from_reads([(list(range(1, L + 1)), [0] * L, [30] * L)])
    prm = default_params(epsilon=0.04)
    same, diff = oracle.score_reads(fr, [0], [255], 1, prm)
    acc = 0.0
    for _ in range(L):
        acc += 0.04
    assert same[0, 0] == 0.0 and diff[0, 0] == acc and acc != L * 0.04 or acc == L * 0.04


def test_ploidy_one_puts_every_read_in_hap0_and_optimize_is_identity():
    c = synth.make_contig(11, 60, 80, 2, span_mean=30)
    prm = default_params(epsilon=0.03125)
    sel = np.arange(c.frags.n_reads, dtype=np.uint32)
    hap, score, _ = oracle.beam_search_phasing(c.frags, sel, 1, prm)
    assert np.all(hap == 0)
    hap2, s2, nr = oracle.optimize_clustering(c.frags, sel, hap, 1, prm)
    assert np.all(hap2 == 0) and nr == 0


@pytest.mark.parametrize("ploidy", [2, 3])
def test_beam_search_assigns_every_read_exactly_once_and_optimize_is_monotone(ploidy):
    c = synth.make_contig(12 + ploidy, 120, 100, ploidy, span_mean=40)
    prm = default_params(epsilon=0.03125)
    sel = np.arange(c.frags.n_reads, dtype=np.uint32)
    hap, score, _ = oracle.beam_search_phasing(c.frags, sel, ploidy, prm)
    assert np.all(hap < ploidy)
    _, err0 = oracle.get_mec_stats_epsilon(c.frags, sel, hap, ploidy, 1, prm)
    hap2, s2, nr = oracle.optimize_clustering(c.frags, sel, hap, ploidy, prm)
    _, err1 = oracle.get_mec_stats_epsilon(c.frags, sel, hap2, ploidy, 1, prm)
    assert s2 == -sum(err1) or abs(s2 + sum(err1)) < 1e-9
    assert -sum(err1) >= -sum(err0)


def test_dyadic_epsilon_makes_scores_exact_multiples():
    c = synth.make_contig(15, 80, 60, 2, span_mean=25)
    prm = default_params(epsilon=0.03125)
    sel = np.arange(c.frags.n_reads, dtype=np.uint32)
    hap = (np.arange(c.frags.n_reads) % 2).astype(np.uint8)
    same, diff = oracle.score_reads(c.frags, sel, hap, 2, prm)
    for a in (same, diff):
        x = a * 2.0 ** 26
        assert np.all(x == np.floor(x))


def test_get_range_with_lengths_shape():
    g = np.arange(0, 1000 * 100, 100, dtype=np.uint64)
    lo, hi = oracle.get_range_with_lengths(g, 10000, 3333, 0.0005)
    assert lo[0] == 1 and hi[-1] == 1000  # 1-indexed, last block closes at the last SNP (utils_frags.rs:420-423, 461)
    assert np.all(lo[1:] > lo[:-1]) and np.all(lo[1:] <= hi[:-1])  # overlapping by ~1/3
    assert np.all(hi - lo <= 101)


def test_find_reads_in_interval_skips_long_spans():
    # local_clustering.rs:36-46
    reads = [([1, 20000], [0, 0], [30, 30]), ([5, 6], [0, 1], [30, 30]), ([50, 60], [0, 1], [30, 30])]
    fr = Frags.from_reads(reads)
    ids = oracle.find_reads_in_interval(1, 10, fr)
    assert list(ids) == [1]


def test_hpop_roundtrip(tmp_path):
    c = synth.make_contig(16, 30, 40, 2, span_mean=10)
    p = tmp_path / "frags.txt"
    c.frags.write_hpop(str(p))
    back = Frags.read_hpop(str(p))
    assert np.array_equal(back.pos, c.frags.pos) and np.array_equal(back.allele, c.frags.allele)
    assert np.array_equal(back.qual, c.frags.qual) and np.array_equal(back.row_ptr, c.frags.row_ptr)
