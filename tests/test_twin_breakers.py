"""Hand-derived cases for the host logic that exists twice in this repository (once in the library, once in the oracle, both
written from the same reading of the Rust).  The expected values below were traced BY HAND through the Rust source, so a
shared misreading of the reference fails here even though library and oracle agree with each other.
CPU tests exercise the oracle; the gpu-marked ones the library through the C-ABI on the same inputs."""
import math

import numpy as np
import pytest

import oracle
from floria_b200 import api, default_params
from floria_b200.frags import Frags


def reads_from_spans(spans, allele=0):
    """one read per (first, last): every position covered, constant allele, q = 30"""
    return Frags.from_reads([(list(range(f, l + 1)), [allele] * (l - f + 1), [30] * (l - f + 1)) for f, l in spans])


# ---- utils_frags.rs:405-463 get_range_with_lengths -------------------------------------------------------------------------------
RANGE_CASES = [
    # 11 SNPs 200 bases apart, block 1000, overlap 333: cum_pos passes 667 at index 4 (new_left_end) and 1000 at index 6:
    # block (0,5); g[4] + 1000 < g[5] is false -> left_endpoint = 5; the last SNP closes (5,10).  1-based: (1,6), (6,11)
    ([10, 200, 400, 600, 800, 1000, 1200, 1400, 1600, 1800, 2000], 1000, 333, 0.0005, [(1, 6), (6, 11)]),
    # a 4800-base hole: at index 3 both thresholds trip at once: new_left_end = 3, block (0,2) is pushed, and because
    # g[3] + 1000 < g[4] is false the next block starts at index 4: SNP index 3 ends up in NO block.  1-based: (1,3), (5,6)
    ([0, 100, 200, 5000, 5100, 5200], 1000, 333, 0.0005, [(1, 3), (5, 6)]),
    # same, but the first block's density 3/1000 is below the minimum: it is dropped
    ([0, 100, 200, 5000, 5100, 5200], 1000, 333, 0.01, [(5, 6)]),
    # a single SNP: the loop's first iteration is the last index
    ([42], 1000, 333, 0.0005, [(1, 1)]),
]


@pytest.mark.parametrize("g,bl,ov,dens,want", RANGE_CASES)
def test_get_range_with_lengths_hand_traced(g, bl, ov, dens, want):
    lo, hi = oracle.get_range_with_lengths(np.array(g, np.uint64), bl, ov, dens)
    assert list(zip(lo.tolist(), hi.tolist())) == want
    lo, hi = api.get_range_with_lengths(np.array(g, np.uint64), bl, ov, dens)  # host code of the library: no GPU needed
    assert list(zip(lo.tolist(), hi.tolist())) == want


# ---- part_block_manip.rs:27-98 separate_broken_haplogroups (through process_reads_for_final_parts) ------------------------------
def broken_case():
    # one haploset, range (1,30), reads A(1,5) B(3,8) | C(12,15) D(14,20) | E(25,30): coverage gaps after 8 and after 20.
    # Pass 1 finds breaks [8, 20].  Pass 2: A, B end <= 8 -> first piece; C ends beyond 8: the piece {A,B} (1,8) is closed,
    # the next piece starts at 9 with end spot 20, and C ITSELF IS NOT INSERTED ANYWHERE (lines 71-84); D ends <= 20 -> second
    # piece; E ends beyond 20: {D} (9,20) is closed, E is dropped too; the trailing piece {} gets (21,30).  The original
    # haploset is cleared but keeps its slot and range.  sort_parts orders by range: (1,8) (1,30) (9,20) (21,30).
    fr = reads_from_spans([(1, 5), (3, 8), (12, 15), (14, 20), (25, 30)])
    want = [((1, 8), [0, 1]), ((1, 30), []), ((9, 20), [3]), ((21, 30), [])]
    return fr, [0, 5], [0, 1, 2, 3, 4], [1], [30], want


def _check_parts(p, want):
    assert p.n_parts == len(want)
    for i, (rng, ids) in enumerate(want):
        assert (int(p.range_lo[i]), int(p.range_hi[i])) == rng
        assert p.read_ids[int(p.part_ptr[i]):int(p.part_ptr[i + 1])].tolist() == ids


def test_separate_broken_haplogroups_drops_the_switching_fragment():
    fr, pp, pr, rl, rh, want = broken_case()
    _check_parts(oracle.process_reads_for_final_parts(fr, pp, pr, rl, rh, default_params()), want)


@pytest.mark.gpu
def test_separate_broken_haplogroups_drops_the_switching_fragment_gpu():
    fr, pp, pr, rl, rh, want = broken_case()
    ctx = api.Context(0)
    _check_parts(ctx.process_reads_for_final_parts(fr, pp, pr, rl, rh, default_params()), want)
    ctx.close()


# ---- part_block_manip.rs:517-620 get_hapq ------------------------------------------------------------------------------------------
def hapq_case():
    # SNP s (1-based) sits at base 100 * (s - 1); block_length 100.
    # part 0: three reads over SNPs 1..4, one of them carries allele 1 at SNP 2 -> 1 error in 12 cells; range (1,4)
    # part 1: three identical reads over SNPs 10..13 -> 0 errors in 12 cells; range (10,13)
    # part 2: ONE read over SNPs 20..23 -> hapq forced to 0 (line 613)
    # no two ranges overlap -> max_penalty = 0, t1 = 40; t2 = min(1, 3/3) = 1; base_range = 300 -> t3 = ln(300/100 + 1) = ln 4
    # hapq = (40 * 1 * 1.3863) as usize = 55.  avg_err = (1 + 0 + 0) / (12 + 12 + 4) = 1/28; rel_err = err_i / avg_err
    reads = [([1, 2, 3, 4], [0, 0, 0, 0], [30] * 4), ([1, 2, 3, 4], [0, 1, 0, 0], [30] * 4), ([1, 2, 3, 4], [0, 0, 0, 0], [30] * 4),
             ([10, 11, 12, 13], [1, 1, 0, 0], [30] * 4), ([10, 11, 12, 13], [1, 1, 0, 0], [30] * 4),
             ([10, 11, 12, 13], [1, 1, 0, 0], [30] * 4), ([20, 21, 22, 23], [0, 1, 0, 1], [30] * 4)]
    fr = Frags.from_reads(reads)
    g = np.arange(30, dtype=np.uint64) * 100
    want_hapq = [int(40.0 * 1.0 * math.log(4.0)), 55, 0]
    avg = 1.0 / 28.0
    want_rel = [(1.0 / 12.0) / avg, 0.0, 0.0]
    return fr, [0, 3, 6, 7], list(range(7)), [1, 10, 20], [4, 13, 23], g, want_hapq, want_rel, avg


def _check_hapq(res, want_hapq, want_rel, avg):
    hapq, rel, a = res
    assert hapq.tolist() == want_hapq
    assert a == avg and rel.tolist() == want_rel


def test_hapq_formula_hand_computed():
    fr, pp, pr, rl, rh, g, wh, wr, avg = hapq_case()
    _check_hapq(oracle.get_hapq(fr, pp, pr, rl, rh, g, default_params(block_length=100)), wh, wr, avg)


@pytest.mark.gpu
def test_hapq_formula_hand_computed_gpu():
    fr, pp, pr, rl, rh, g, wh, wr, avg = hapq_case()
    ctx = api.Context(0)
    _check_hapq(ctx.get_hapq(fr, pp, pr, rl, rh, g, default_params(block_length=100)), wh, wr, avg)
    ctx.close()


def test_hapq_overlap_penalty_hand_computed():
    # two haplosets over the SAME range with the SAME consensus: overlap_percent = 1 > 0.05, distance 0 -> max_penalty = 1,
    # t1 = 40 * (1 - 1) = 0 -> hapq 0 for both (part_block_manip.rs:556-571, 601)
    reads = [([1, 2, 3, 4], [0, 1, 0, 1], [30] * 4)] * 6
    fr = Frags.from_reads(reads)
    g = np.arange(10, dtype=np.uint64) * 100
    hapq, rel, avg = oracle.get_hapq(fr, [0, 3, 6], list(range(6)), [1, 1], [4, 4], g, default_params(block_length=100))
    assert hapq.tolist() == [0, 0]


# ---- local_clustering.rs:12-59 find_reads_in_interval ------------------------------------------------------------------------------
def test_find_reads_in_interval_hand_traced():
    # reads sorted by first position; interval [10, 20]: last < 10 skipped, first > 20 stops the scan (`break`), spans above
    # 10000 SNPs are skipped (line 44)
    # Frag::cmp order (first asc, last desc): (1,10)=0 (1,9)=1 (5,30)=2 (10,10)=3 (20,25)=4 (21,40)=5
    fr = reads_from_spans([(1, 9), (1, 10), (5, 30), (10, 10), (20, 25), (21, 40)])
    assert oracle.find_reads_in_interval(10, 20, fr).tolist() == [0, 2, 3, 4]
    assert api.find_reads_in_interval(10, 20, fr).tolist() == [0, 2, 3, 4]
    first = np.array([1, 2], np.uint32)
    last = np.array([10002, 10002], np.uint32)  # spans 10001 (> 10000: skipped) and 10000 (kept)
    big = Frags.__new__(Frags)
    big.first, big.last, big.n_reads = first, last, 2
    assert oracle.find_reads_in_interval(5, 6, big).tolist() == [1]
    assert api.find_reads_in_interval(5, 6, big).tolist() == [1]
