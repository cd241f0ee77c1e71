"""k_beam_wide (one beam-search instance spread over the whole GPU, fb_beam_wide.cuh) against the CPU oracle, against
k_beam (one CTA per instance) and against the committed configs[2]-shape golden fixture.

FB_BEAM_WIDE=1 forces every ploidy >= 2 instance through the grid-wide kernel, FB_BEAM_WIDE=0 through k_beam;
FB_BEAM_WIDE_GRID caps the cooperative grid (the SNP axis is dealt to the CTAs in 32-position chunks, so odd grid
sizes exercise the ownership arithmetic)."""
import os

import numpy as np
import pytest

import oracle
from floria_b200 import api, default_params, synth
from test_gpu_parity import CASES, EPS, _compare_block_results, assert_f64_identical, bits

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config3_fullspan_150x50k.npz")


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("case", ["long", "short", "edge", "wide"])
@pytest.mark.parametrize("eps", EPS)
@pytest.mark.parametrize("ploidy", [2, 3, 4])
def test_wide_beam_matches_oracle(ctx, monkeypatch, case, eps, ploidy):
    monkeypatch.setenv("FB_BEAM_WIDE", "1")
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
    assert np.allclose(gl, ol, rtol=1e-5, atol=1e-12)  # per-read log-likelihoods: device libm vs host libm
    assert np.array_equal(gh, oh), f"{int((gh != oh).sum())} of {n} assignments differ"
    assert_f64_identical([gsc], [osc], "best score")


@pytest.mark.parametrize("grid", [1, 2, 7, 33, 148])
def test_wide_beam_grid_sizes(ctx, monkeypatch, grid):
    """the result does not depend on how many CTAs share the SNP axis"""
    monkeypatch.setenv("FB_BEAM_WIDE", "1")
    monkeypatch.setenv("FB_BEAM_WIDE_GRID", str(grid))
    for case, ploidy, eps in (("long", 3, 0.04), ("short", 2, 0.03125), ("edge", 4, 0.04), ("wide", 2, 0.04)):
        fr = CASES[case]()
        prm = default_params(epsilon=eps)
        sel = np.arange(min(fr.n_reads, 200), dtype=np.uint32)
        oh, osc, _ = oracle.beam_search_phasing(fr, sel, ploidy, prm)
        gh, gsc, _ = ctx.beam_search_phasing(fr, sel, ploidy, prm)
        assert np.array_equal(gh, oh), f"grid={grid} {case}: {int((gh != oh).sum())} assignments differ"
        assert_f64_identical([gsc], [osc], "beam score")


@pytest.mark.parametrize("ploidy", [6, 8])
def test_wide_beam_high_ploidy_short_reads(ctx, monkeypatch, ploidy):
    """many equal states (short windows): exercises the exact comparison + the second grid barrier of a step"""
    monkeypatch.setenv("FB_BEAM_WIDE", "1")
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


def test_wide_beam_small_beam_and_single_read(ctx, monkeypatch):
    monkeypatch.setenv("FB_BEAM_WIDE", "1")
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


def test_wide_phase_blocks_matches_oracle(ctx, monkeypatch):
    """the batched call with every beam instance on the grid-wide kernel (instances run one after the other)"""
    monkeypatch.setenv("FB_BEAM_WIDE", "1")
    c = synth.make_contig(52, 260, 240, 3, span_mean=70)
    prm = default_params(epsilon=0.04, max_ploidy=4)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 6000, 2000, 0.0005)
    g = ctx.phase_blocks(c.frags, lo, hi, prm)
    o = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=4)
    _compare_block_results(g, o)


def test_wide_beam_long_reads_vs_oracle(ctx, monkeypatch):
    """configs[2] shape at reduced width: full-span reads over 6000 SNPs (375 groups per read: the automatic choice
    is the grid-wide kernel), ploidy 4, both epsilons, against the oracle"""
    monkeypatch.delenv("FB_BEAM_WIDE", raising=False)
    c = synth.make_contig(3, 64, 6000, 4, full_span=True)
    sel = np.arange(64, dtype=np.uint32)
    for eps in EPS:
        prm = default_params(epsilon=eps, max_ploidy=4)
        oh, osc, (os_, od, ol, on) = oracle.beam_search_phasing(c.frags, sel, 4, prm, tap_cap=20000)
        gh, gsc, (gs, gd, gl, gn) = ctx.beam_search_phasing(c.frags, sel, 4, prm, tap_cap=20000)
        assert gn == on
        assert_f64_identical(gs, os_, "tap same")
        assert_f64_identical(gd, od, "tap diff")
        assert np.allclose(gl, ol, rtol=1e-5, atol=1e-12)
        assert np.array_equal(gh, oh)
        assert_f64_identical([gsc], [osc], "best score")


def test_wide_equals_narrow_at_depth(ctx, monkeypatch):
    """2500 reads x 12 000 SNPs, ploidy 4 (banded, mean span 6000): far beyond what the oracle finishes in seconds; the
    two kernels share only the decision section, so bit-equal assignments and scores tie the grid-wide reductions,
    slices and barriers to the one-CTA kernel that the oracle tests pin at small sizes"""
    c = synth.make_contig(77, 2500, 12000, 4, span_mean=6000)
    sel = np.arange(c.frags.n_reads, dtype=np.uint32)
    prm = default_params(epsilon=0.04, max_ploidy=4)
    monkeypatch.setenv("FB_BEAM_WIDE", "0")
    nh, nsc, (ns, nd, nl, nn) = ctx.beam_search_phasing(c.frags, sel, 4, prm, tap_cap=100000)
    monkeypatch.setenv("FB_BEAM_WIDE", "1")
    wh, wsc, (ws, wd, wl, wn) = ctx.beam_search_phasing(c.frags, sel, 4, prm, tap_cap=100000)
    assert wn == nn
    assert_f64_identical(ws, ns, "tap same")
    assert_f64_identical(wd, nd, "tap diff")
    assert_f64_identical(wl, nl, "tap log-p")  # same device libm on both paths
    assert np.array_equal(wh, nh)
    assert_f64_identical([wsc], [nsc], "best score")
    # the beam recovers the truth partition up to a relabelling on this clean input
    conf = np.zeros((4, 4), np.int64)
    np.add.at(conf, (c.read_hap, wh), 1)
    assert (conf.max(axis=1).sum()) >= 0.99 * len(wh)


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden fixture not generated")
def test_config3_shape_golden(ctx, monkeypatch):
    """configs[2] shape (full-span reads x 50 000 SNPs, ploidy 4, beam 10) at 150 reads: beam -> optimize -> no-phred MEC
    through the C-ABI against the oracle outputs committed by tools/make_golden_c3.py"""
    monkeypatch.delenv("FB_BEAM_WIDE", raising=False)
    z = np.load(GOLDEN)
    n, S, p = int(z["n_reads"]), int(z["n_snps"]), int(z["ploidy"])
    c = synth.make_contig(3, n, S, p, full_span=True)
    fr = c.frags
    chk = np.array([int(fr.pos.astype(np.uint64).sum()), int(fr.allele.astype(np.uint64).sum()),
                    int(fr.qual.astype(np.uint64).sum()), int(fr.nnz)], dtype=np.uint64)
    assert np.array_equal(chk, z["checksum"]), "the synthetic generator drifted: regenerate the fixture"
    prm = default_params(epsilon=float(z["epsilon"]), max_ploidy=p)
    sel = np.arange(n, dtype=np.uint32)
    cap = len(z["tap_same"])
    gh, gsc, (gs, gd, gl, gn) = ctx.beam_search_phasing(fr, sel, p, prm, tap_cap=cap)
    assert gn == int(z["tap_n"])
    assert_f64_identical(gs, z["tap_same"], "tap same")
    assert_f64_identical(gd, z["tap_diff"], "tap diff")
    assert np.allclose(gl, z["tap_logp"], rtol=1e-5, atol=1e-12)
    assert np.array_equal(gh, z["beam_hap"])
    assert_f64_identical([gsc], [float(z["beam_score"])], "beam score")
    oh, osc, rounds = ctx.optimize_clustering(fr, sel, gh, p, prm)
    assert np.array_equal(oh, z["opt_hap"]) and rounds == int(z["opt_rounds"])
    assert_f64_identical([osc], [float(z["opt_score"])], "optimize score")
    gb, ge = ctx.get_mec_stats_epsilon(fr, sel, oh, p, 0, prm)
    assert_f64_identical(gb, z["mec_bases"], "no-phred bases")
    assert_f64_identical(ge, z["mec_errors"], "no-phred errors (the MEC of the ploidy loop)")


@pytest.mark.parametrize("wide", ["1", "0"])
def test_phase_block_with_pipelined_upload(ctx, monkeypatch, wide):
    """fb_phase_block from host buffers with the upload pipelined under the beam search (chunks of reads copied and packed on
    a second stream while k_beam_wide already consumes the leading reads; k_beam waits for the whole contig): same result as
    the plain upload and as the oracle"""
    monkeypatch.setenv("FB_BEAM_WIDE", wide)
    for seed, n, S, p, kw in ((81, 300, 400, 3, dict(span_mean=120)), (82, 90, 6000, 4, dict(full_span=True))):
        c = synth.make_contig(seed, n, S, p, **kw)
        prm = default_params(epsilon=0.04, max_ploidy=p)
        sel = np.arange(0, c.frags.n_reads, dtype=np.uint32)
        oh, ob, oe, oi = oracle.phase_block(c.frags, sel, p, prm)
        monkeypatch.setenv("FB_PIPELINE_UPLOAD", "1")
        gh, gb, ge, gi = ctx.phase_block(c.frags, sel, p, prm)
        monkeypatch.setenv("FB_PIPELINE_UPLOAD", "0")
        hh, hb, he, hi = ctx.phase_block(c.frags, sel, p, prm)
        for h, e, i in ((gh, ge, gi), (hh, he, hi)):
            assert np.array_equal(h, oh)
            assert_f64_identical(e, oe, "no-phred errors")
            assert (i["cells_sweep"], i["cells_hist"], i["cells_beam"]) == (oi["cells_sweep"], oi["cells_hist"], oi["cells_beam"])
            assert_f64_identical([i["beam_score"], i["opt_score"]], [oi["beam_score"], oi["opt_score"]], "scores")


def test_pipelined_upload_reports_bad_cells(ctx, monkeypatch):
    monkeypatch.setenv("FB_PIPELINE_UPLOAD", "1")
    c = synth.make_contig(83, 60, 300, 2, span_mean=100)
    bad = c.frags.allele.copy()
    bad[len(bad) // 2] = 7
    fr = type(c.frags)(c.frags.row_ptr, c.frags.pos, bad, c.frags.qual, c.frags.first, c.frags.last)
    with pytest.raises(api.FloriaB200Error, match="invalid cell"):
        ctx.phase_block(fr, None, 2, default_params())
    # the context stays usable
    h, b, e, i = ctx.phase_block(c.frags, None, 2, default_params())
    assert len(h) == c.frags.n_reads
