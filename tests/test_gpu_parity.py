"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): integer outputs (assignments, ploidy, counters) and f64 scores bit-exact;
per-read log-likelihoods within 1e-5 relative.  Both a dyadic epsilon (0.03125: every sum exact in any order) and
the reference-style epsilon 0.04 (exercises the exact left-to-right f64 replay) are covered.
"""
import numpy as np
import pytest

import oracle
from floria_b200 import api, default_params, synth
from floria_b200.frags import Frags

pytestmark = pytest.mark.gpu

EPS = [0.03125, 0.04]


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_f64_identical(a, b, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    bad = np.nonzero(bits(a) != bits(b))[0] if a.ndim == 1 else np.argwhere(bits(a) != bits(b))
    assert len(bad) == 0, f"{what}: {len(bad)} f64 values differ, first at {bad[0]}: {a[tuple(np.atleast_1d(bad[0]))]!r} vs {b[tuple(np.atleast_1d(bad[0]))]!r}"


def edge_contig(seed, n_reads=90, n_snps=70, max_allele=3):
    """ragged reads with holes, q=0 and q=255 cells, up to 4 alleles, single-cell reads"""
    rng = np.random.default_rng(seed)
    reads = []
    for i in range(n_reads):
        span = int(rng.integers(1, 40))
        first = int(rng.integers(1, n_snps - span + 2))
        pos = [p for p in range(first, first + span) if rng.random() < 0.8 or p in (first, first + span - 1)]
        al = rng.integers(0, max_allele + 1, len(pos))
        q = rng.choice([0, 1, 2, 3, 10, 20, 30, 40, 60, 93, 255], len(pos))
        reads.append((pos, al, q))
    return Frags.from_reads(reads)


def random_assignment(rng, n, ploidy, frac_unassigned=0.1):
    hap = rng.integers(0, ploidy, n).astype(np.uint8)
    hap[rng.random(n) < frac_unassigned] = 255
    return hap


CASES = {
    "long": lambda: synth.make_contig(31, 300, 260, 3, span_mean=80).frags,
    "short": lambda: synth.make_contig(32, 600, 200, 3, paired_short=True, flip=0.01, qual_mode="short").frags,
    "edge": lambda: edge_contig(33),
    "wide": lambda: synth.make_contig(34, 40, 2500, 2, span_mean=1400).frags,  # > 1024 positions: several hist tiles
}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("eps", EPS)
@pytest.mark.parametrize("ploidy", [1, 2, 4])
def test_score_reads_matches_oracle(ctx, case, eps, ploidy):
    fr = CASES[case]()
    rng = np.random.default_rng(5)
    prm = default_params(epsilon=eps)
    sel = np.arange(fr.n_reads, dtype=np.uint32)
    if case != "wide":
        sel = sel[rng.random(fr.n_reads) < 0.8]  # a block is a subset of the contig's reads
    hap = random_assignment(rng, len(sel), ploidy)
    o_same, o_diff = oracle.score_reads(fr, sel, hap, ploidy, prm)
    same, diff, sq, dq, ne = ctx.score_reads(fr, sel, hap, ploidy, prm)
    assert_f64_identical(same, o_same, "same")
    assert_f64_identical(diff, o_diff, "diff")
    assert np.array_equal(sq.astype(np.float64) * 2.0 ** -26, same)
    # diff = exact weight part + n_empty epsilons (closed form only guaranteed for a dyadic epsilon)
    if eps == 0.03125:
        assert np.array_equal(dq.astype(np.float64) * 2.0 ** -26 + ne * eps, diff)


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("use_qual", [1, 0])
def test_hap_block_from_partition_matches_oracle(ctx, case, use_qual):
    fr = CASES[case]()
    rng = np.random.default_rng(6)
    prm = default_params(epsilon=0.04)
    sel = np.arange(fr.n_reads, dtype=np.uint32)
    hap = random_assignment(rng, len(sel), 3)
    lo = int(fr.first.min())
    n = int(fr.last.max()) - lo + 1
    oc, om = oracle.hap_block_from_partition(fr, sel, hap, 3, use_qual, prm, lo, n)
    gc, gm = ctx.hap_block_from_partition(fr, sel, hap, 3, use_qual, prm, lo, n)
    assert_f64_identical(gc.ravel(), oc.ravel(), "counts")
    assert np.array_equal(gm, om)


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("eps", EPS)
@pytest.mark.parametrize("use_phred", [1, 0])
def test_mec_stats_match_oracle(ctx, case, eps, use_phred):
    fr = CASES[case]()
    rng = np.random.default_rng(8)
    prm = default_params(epsilon=eps)
    sel = np.arange(fr.n_reads, dtype=np.uint32)
    hap = random_assignment(rng, len(sel), 3, 0.0)
    ob, oe = oracle.get_mec_stats_epsilon(fr, sel, hap, 3, use_phred, prm)
    gb, ge = ctx.get_mec_stats_epsilon(fr, sel, hap, 3, use_phred, prm)
    assert_f64_identical(gb, ob, "bases")
    assert_f64_identical(ge, oe, "errors")


@pytest.mark.parametrize("case", ["long", "short", "edge"])
@pytest.mark.parametrize("eps", EPS)
@pytest.mark.parametrize("ploidy", [1, 2, 3, 5])
def test_optimize_clustering_matches_oracle(ctx, case, eps, ploidy):
    fr = CASES[case]()
    rng = np.random.default_rng(9 + ploidy)
    prm = default_params(epsilon=eps)
    sel = np.arange(fr.n_reads, dtype=np.uint32)
    hap = random_assignment(rng, len(sel), ploidy, 0.0)
    oh, osc, onr = oracle.optimize_clustering(fr, sel, hap, ploidy, prm)
    gh, gsc, gnr = ctx.optimize_clustering(fr, sel, hap, ploidy, prm)
    assert np.array_equal(gh, oh), f"{int((gh != oh).sum())} assignments differ"
    assert_f64_identical([gsc], [osc], "score")
    assert gnr == onr


@pytest.mark.parametrize("team", [32, 8, 2])
@pytest.mark.parametrize("case", ["long", "short", "edge"])
def test_sweep_team_sizes_match_oracle(ctx, monkeypatch, team, case):
    """k_sweep gives a read a whole warp or a team of 8 / 2 lanes depending on how many 16-SNP groups the reads span
    (chosen per launch from the average); every team size must give the oracle's answer on every kind of input,
    including the exact epsilon replay (eps = 0.04) inside a team."""
    monkeypatch.setenv("FB_SWEEP_TEAM", str(team))
    fr = CASES[case]()
    rng = np.random.default_rng(21)
    sel = np.arange(fr.n_reads, dtype=np.uint32)
    for eps in EPS:
        prm = default_params(epsilon=eps)
        for ploidy in (2, 3, 5):
            hap = random_assignment(rng, len(sel), ploidy)
            o_same, o_diff = oracle.score_reads(fr, sel, hap, ploidy, prm)
            same, diff, _, _, _ = ctx.score_reads(fr, sel, hap, ploidy, prm)
            assert_f64_identical(same, o_same, f"same team={team} p={ploidy}")
            assert_f64_identical(diff, o_diff, f"diff team={team} p={ploidy}")
            hap0 = random_assignment(rng, len(sel), ploidy, 0.0)
            oh, osc, onr = oracle.optimize_clustering(fr, sel, hap0, ploidy, prm)
            gh, gsc, gnr = ctx.optimize_clustering(fr, sel, hap0, ploidy, prm)
            assert np.array_equal(gh, oh) and gnr == onr
            assert_f64_identical([gsc], [osc], "score")


def test_optimize_keeps_singleton_haplotypes(ctx):
    # `if new_part[i].len() == 1 { continue; }` (local_clustering.rs:346) and `partition[i].len() <= 1` (:301)
    fr = CASES["long"]()
    prm = default_params(epsilon=0.03125)
    sel = np.arange(fr.n_reads, dtype=np.uint32)
    hap = np.zeros(len(sel), np.uint8)
    hap[0] = 1
    hap[1:3] = 2
    oh, osc, onr = oracle.optimize_clustering(fr, sel, hap, 3, prm)
    gh, gsc, gnr = ctx.optimize_clustering(fr, sel, hap, 3, prm)
    assert np.array_equal(gh, oh) and gnr == onr
    assert_f64_identical([gsc], [osc])


@pytest.mark.parametrize("case", ["long", "short", "edge"])
@pytest.mark.parametrize("eps", EPS)
@pytest.mark.parametrize("ploidy", [2, 3, 4])
def test_beam_search_matches_oracle(ctx, case, eps, ploidy):
    fr = CASES[case]()
    prm = default_params(epsilon=eps)
    n = min(fr.n_reads, 160)
    sel = np.arange(n, dtype=np.uint32)
    cap = 200000
    oh, osc, (os_, od, ol, on) = oracle.beam_search_phasing(fr, sel, ploidy, prm, tap_cap=cap)
    gh, gsc, (gs, gd, gl, gn) = ctx.beam_search_phasing(fr, sel, ploidy, prm, tap_cap=cap)
    assert gn == on, "number of (node, haplotype) evaluations differs"
    assert_f64_identical(gs, os_, "tap same")
    assert_f64_identical(gd, od, "tap diff")
    # per-read log-likelihoods within 1e-5 relative (device libm vs host libm)
    assert np.allclose(gl, ol, rtol=1e-5, atol=1e-12)
    assert np.array_equal(gh, oh), f"{int((gh != oh).sum())} of {n} assignments differ"
    assert_f64_identical([gsc], [osc], "best score")


@pytest.mark.parametrize("ploidy", [6, 8])
def test_beam_search_high_ploidy_short_reads(ctx, ploidy):
    """many duplicate blocks (short windows) and the widest instantiations of the kernel"""
    fr = CASES["short"]()
    prm = default_params(epsilon=0.01, max_number_solns=4)
    sel = np.arange(90, dtype=np.uint32)
    oh, osc, (os_, od, ol, on) = oracle.beam_search_phasing(fr, sel, ploidy, prm, tap_cap=400000)
    gh, gsc, (gs, gd, gl, gn) = ctx.beam_search_phasing(fr, sel, ploidy, prm, tap_cap=400000)
    assert gn == on
    assert_f64_identical(gs, os_, "tap same")
    assert_f64_identical(gd, od, "tap diff")
    assert np.array_equal(gh, oh)
    assert_f64_identical([gsc], [osc], "best score")


@pytest.mark.parametrize("cta", [128, 256])
def test_beam_cta_sizes_match_oracle(ctx, monkeypatch, cta):
    """k_beam runs 256-thread CTAs (one per SM) for short work queues and 128-thread CTAs (two per SM) for long ones;
    both instantiations must reproduce the oracle: assignments, score bits, and the whole phase_blocks result."""
    monkeypatch.setenv("FB_BEAM_CTA", str(cta))
    for case, ploidy, eps in (("long", 3, 0.04), ("short", 2, 0.03125), ("edge", 4, 0.04)):
        fr = CASES[case]()
        prm = default_params(epsilon=eps)
        sel = np.arange(fr.n_reads, dtype=np.uint32)
        oh, osc, _ = oracle.beam_search_phasing(fr, sel, ploidy, prm)
        gh, gsc, _ = ctx.beam_search_phasing(fr, sel, ploidy, prm)
        assert np.array_equal(gh, oh), f"cta={cta} {case}: {int((gh != oh).sum())} assignments differ"
        assert_f64_identical([gsc], [osc], "beam score")
    c = synth.make_contig(52, 260, 240, 3, span_mean=70)
    prm = default_params(epsilon=0.04, max_ploidy=4)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 6000, 2000, 0.0005)
    g = ctx.phase_blocks(c.frags, lo, hi, prm)
    o = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=4)
    assert np.array_equal(g.best_ploidy, o.best_ploidy) and np.array_equal(g.hap, o.hap)
    assert np.array_equal(bits(g.mec_vector), bits(o.mec_vector)) and g.cells == o.cells


@pytest.mark.parametrize("cta", [128, 256])
def test_beam_search_reads_longer_than_the_staging_buffer(ctx, monkeypatch, cta):
    """k_beam stages a read's planes in shared memory when it spans at most FB_BEAM_RG = 128 groups (2048 SNPs) and
    reads longer ones from global memory: a contig whose reads span ~2300 of 2600 SNPs takes the second path on almost
    every step (and mixes both), for both CTA sizes."""
    monkeypatch.setenv("FB_BEAM_CTA", str(cta))
    c = synth.make_contig(61, 110, 2600, 2, span_mean=2300)
    fr = c.frags
    span = (fr.last.astype(np.int64) - fr.first.astype(np.int64))
    assert (span > 2100).sum() >= 20 and (span <= 2000).sum() >= 5
    sel = np.arange(fr.n_reads, dtype=np.uint32)
    for eps in EPS:
        prm = default_params(epsilon=eps)
        oh, osc, _ = oracle.beam_search_phasing(fr, sel, 2, prm)
        gh, gsc, _ = ctx.beam_search_phasing(fr, sel, 2, prm)
        assert np.array_equal(gh, oh), f"{int((gh != oh).sum())} assignments differ"
        assert_f64_identical([gsc], [osc], "beam score")


def test_beam_search_small_beam_and_single_read(ctx):
    fr = CASES["edge"]()
    for B in (1, 3):
        prm = default_params(epsilon=0.04, max_number_solns=B)
        sel = np.arange(60, dtype=np.uint32)
        oh, osc, _ = oracle.beam_search_phasing(fr, sel, 3, prm)
        gh, gsc, _ = ctx.beam_search_phasing(fr, sel, 3, prm)
        assert np.array_equal(gh, oh)
        assert_f64_identical([gsc], [osc])
    prm = default_params(epsilon=0.04)
    oh, osc, _ = oracle.beam_search_phasing(fr, [5], 2, prm)
    gh, gsc, _ = ctx.beam_search_phasing(fr, [5], 2, prm)
    assert np.array_equal(gh, oh)
    assert_f64_identical([gsc], [osc])


def _compare_block_results(g, o):
    assert np.array_equal(g.best_ploidy, o.best_ploidy)
    assert np.array_equal(g.ploidies_run, o.ploidies_run)
    assert_f64_identical(g.mec_vector.ravel(), o.mec_vector.ravel(), "mec_vector")
    assert_f64_identical(g.expected_errors.ravel(), o.expected_errors.ravel(), "expected_errors")
    assert np.array_equal(g.read_ptr, o.read_ptr)
    assert np.array_equal(g.read_ids, o.read_ids)
    assert np.array_equal(g.hap, o.hap), f"{int((g.hap != o.hap).sum())} assignments differ"
    assert (g.cells_sweep, g.cells_hist, g.cells_beam) == (o.cells_sweep, o.cells_hist, o.cells_beam)


@pytest.mark.parametrize("eps", EPS)
@pytest.mark.parametrize("truth_ploidy,max_ploidy", [(2, 3), (3, 5)])
def test_phase_blocks_matches_oracle(ctx, eps, truth_ploidy, max_ploidy):
    c = synth.make_contig(40 + truth_ploidy, 350, 320, truth_ploidy, span_mean=60)
    prm = default_params(epsilon=eps, max_ploidy=max_ploidy)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005)
    o = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=8)
    g = ctx.phase_blocks(c.frags, lo, hi, prm)
    _compare_block_results(g, o)
    # resident variant (inputs already in HBM) gives the same answer
    d = ctx.upload(c.frags)
    g2 = ctx.phase_blocks_resident(d, lo, hi, prm)
    d.free()
    _compare_block_results(g2, o)


def test_phase_blocks_short_reads_and_empty_blocks(ctx):
    c = synth.make_contig(44, 1500, 400, 3, paired_short=True, flip=0.01, qual_mode="short")
    prm = default_params(epsilon=0.01, max_ploidy=3)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 500, 500 // 3, 0.0005)
    # add an interval nobody covers (reference returns None for it, graph_processing.rs:129-131)
    lo = np.append(lo, np.uint32(100000)).astype(np.uint32)
    hi = np.append(hi, np.uint32(100010)).astype(np.uint32)
    o = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=8)
    g = ctx.phase_blocks(c.frags, lo, hi, prm)
    _compare_block_results(g, o)
    assert g.best_ploidy[-1] == 0


def test_no_stop_heuristic_and_sensitivities(ctx):
    c = synth.make_contig(45, 200, 150, 2, span_mean=50)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005)
    for kw in (dict(stopping_heuristic=0), dict(ploidy_sensitivity=1), dict(ploidy_sensitivity=3)):
        prm = default_params(epsilon=0.03125, max_ploidy=3, **kw)
        _compare_block_results(ctx.phase_blocks(c.frags, lo, hi, prm), oracle.phase_blocks(c.frags, lo, hi, prm, 8))


def test_rejects_bad_input(ctx):
    fr = CASES["long"]()
    prm = default_params(epsilon=0.04)
    bad = Frags(fr.row_ptr, fr.pos, fr.allele.copy(), fr.qual, fr.first, fr.last)
    bad.allele[3] = 4
    with pytest.raises(api.FloriaB200Error):
        ctx.upload(bad)
    unsorted = fr.subset(np.arange(fr.n_reads)[::-1])
    with pytest.raises(api.FloriaB200Error):
        ctx.upload(unsorted)
    with pytest.raises(api.FloriaB200Error):
        ctx.score_reads(fr, [0, 1], [0, 0], 1, default_params(epsilon=0.04, order_model=1))


def test_encode_linearity_property_at_scale(ctx):
    """Size-independent property at a larger size than the oracle is run on: the histogram is linear in the partition
    (counts of a union = sum of counts), and same + diff + n_empty*eps accounts for every stored cell's weight."""
    c = synth.make_contig(46, 4000, 3000, 4, span_mean=400)
    fr = c.frags
    prm = default_params(epsilon=0.03125)
    sel = np.arange(fr.n_reads, dtype=np.uint32)
    rng = np.random.default_rng(1)
    hap = rng.integers(0, 4, fr.n_reads).astype(np.uint8)
    lo, n = 1, 3000
    full, _ = ctx.hap_block_from_partition(fr, sel, hap, 4, 1, prm, lo, n)
    merged, _ = ctx.hap_block_from_partition(fr, sel, np.zeros_like(hap), 1, 1, prm, lo, n)
    assert np.array_equal(full.sum(axis=0), merged[0])
    same, diff, sq, dq, ne = ctx.score_reads(fr, sel, hap, 4, prm)
    lut = oracle.phred_lut().astype(np.float64)
    tot = np.add.reduceat(lut[fr.qual], fr.row_ptr[:-1].astype(np.int64))
    for h in range(4):
        assert np.array_equal(same[:, h] + (diff[:, h] - ne[:, h] * 0.03125) + 0.0 * tot, same[:, h] + dq[:, h] * 2.0 ** -26)
        # every present cell is exactly one of same / diff / empty
        w_empty = tot - (sq[:, h] + dq[:, h]) * 2.0 ** -26
        assert np.all(w_empty >= 0)
        assert np.all((ne[:, h] == 0) <= (w_empty == 0))


@pytest.mark.parametrize("wave", [0, 1, 2, 3])
def test_phase_blocks_ploidy_waves_match_oracle(ctx, monkeypatch, wave):
    """the ploidy loop advances in waves on the device (first wave: ploidies 1..FB_PLOIDY_WAVE, then one ploidy per wave for
    the blocks whose stopping rule has not fired); every schedule must give the oracle's result, counters included"""
    monkeypatch.setenv("FB_PLOIDY_WAVE", str(wave))
    for seed, truth, mp in ((71, 2, 5), (72, 4, 6), (73, 1, 4)):
        c = synth.make_contig(seed, 320, 280, truth, span_mean=70)
        prm = default_params(epsilon=0.04, max_ploidy=mp)
        lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 6000, 2000, 0.0005)
        g = ctx.phase_blocks(c.frags, lo, hi, prm)
        o = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=4)
        _compare_block_results(g, o)
        assert len(set(g.ploidies_run.tolist())) > 1 or mp <= 2, "the inputs should stop at different ploidies"
    prm = default_params(epsilon=0.04, max_ploidy=5, stopping_heuristic=0)
    c = synth.make_contig(74, 200, 200, 2, span_mean=60)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 6000, 2000, 0.0005)
    _compare_block_results(ctx.phase_blocks(c.frags, lo, hi, prm), oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=4))
