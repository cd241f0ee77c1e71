"""Parity at BASELINE.json sizes, through the C-ABI:
  configs[1]  the full 10k x 5k contig, all 73 blocks against the oracle (mec_vector bits included)
  configs[2]  the full 100k x 50k block: size-independent properties (truth recovery, score == recomputation from the
              histogram tables); the oracle-feasible depth of this shape is tests/test_gpu_beam_wide.py
  configs[3]  0.05 scale (100k paired short reads x 5k SNPs, ploidy 3) against the oracle
  configs[4]  20 contigs at real shape (2000 reads x 1000 SNPs, truth ploidy 2..6) through fb_phase_contigs against the
              committed oracle fixture (tools/make_golden_c5.py), 3 of them against the live oracle"""
import os

import numpy as np
import pytest

import oracle
from floria_b200 import api, default_params, synth
from test_gpu_parity import _compare_block_results, assert_f64_identical

pytestmark = pytest.mark.gpu
GOLDEN_C5 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config5_20contigs.npz")


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def test_config1_full_contig_matches_oracle(ctx):
    c = synth.config2()  # BASELINE.json configs[1]
    prm = default_params(epsilon=0.04, max_ploidy=2, block_length=10000)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005)
    assert len(lo) == 73
    g = ctx.phase_blocks(c.frags, lo, hi, prm)
    o = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=os.cpu_count() or 1)
    _compare_block_results(g, o)
    assert np.array_equal(g.block_cells, o.block_cells)


def test_config3_short_reads_at_5_percent_match_oracle(ctx):
    c = synth.config4(0.05)  # BASELINE.json configs[3] x 0.05
    prm = default_params(epsilon=0.01, max_ploidy=3, block_length=500)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 500, 500 // 3, 0.0005)
    assert c.frags.n_reads == 100000 and len(lo) > 900
    g = ctx.phase_blocks(c.frags, lo, hi, prm)
    o = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=os.cpu_count() or 1)
    _compare_block_results(g, o)


@pytest.mark.skipif(not os.path.exists(GOLDEN_C5), reason="golden fixture not generated")
def test_config4_20_contigs_through_phase_contigs():
    z = np.load(GOLDEN_C5)
    n = int(z["n_contigs"])
    prm = default_params(epsilon=float(z["epsilon"]), max_ploidy=int(z["max_ploidy"]))
    cs, blocks = [], []
    for k in range(n):
        c = synth.config5_contig(k)
        fr = c.frags
        chk = np.array([int(fr.pos.astype(np.uint64).sum()), int(fr.allele.astype(np.uint64).sum()),
                        int(fr.qual.astype(np.uint64).sum()), int(fr.nnz)], dtype=np.uint64)
        assert np.array_equal(chk, z[f"chk_{k}"]), "the synthetic generator drifted: regenerate the fixture"
        cs.append(fr)
        blocks.append(api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005))
    m = api.MultiContext([0, 0])  # two contexts / host threads / streams on the one device of the test box
    res, dev, ms = m.phase_contigs(cs, blocks, prm)
    assert set(dev.tolist()) == {0, 1}
    for k in range(n):
        r = res[k]
        assert np.array_equal(r.best_ploidy, z[f"best_{k}"]), f"contig {k}: best ploidy"
        assert np.array_equal(r.ploidies_run, z[f"run_{k}"])
        assert_f64_identical(r.mec_vector.ravel(), z[f"mec_{k}"].ravel(), f"contig {k} mec_vector")
        assert_f64_identical(r.expected_errors.ravel(), z[f"exp_{k}"].ravel(), f"contig {k} expected_errors")
        assert np.array_equal(r.read_ptr, z[f"ptr_{k}"])
        assert int(r.read_ids.astype(np.uint64).sum()) == int(z[f"ids_sum_{k}"])
        assert np.array_equal(r.hap, z[f"hap_{k}"]), f"contig {k}: {int((r.hap != z[f'hap_{k}']).sum())} assignments differ"
        assert np.array_equal(r.block_cells, z[f"cells_{k}"])
    for k in (0, 7, 19):  # and the live oracle on three of them
        o = oracle.phase_blocks(cs[k], blocks[k][0], blocks[k][1], prm, n_threads=os.cpu_count() or 1)
        assert np.array_equal(res[k].hap, o.hap) and np.array_equal(res[k].best_ploidy, o.best_ploidy)
        assert_f64_identical(res[k].mec_vector.ravel(), o.mec_vector.ravel(), "mec_vector")
    m.close()


def test_config2_full_size_properties(ctx):
    """the 100k x 50k block (4.9e9 cells) end to end at a dyadic epsilon: every read in exactly one haplotype, the truth
    partition recovered up to a relabelling, and optimize_clustering's score equal, bit for bit, to the MEC recomputed on
    the host from the phred-weighted histogram table of the returned partition (an independent kernel path: k_hist
    through fb_bench_block_tables, sums in numpy)"""
    P, eps = 4, 0.03125
    d = ctx.bench_synth_dense(100000, 50000, P, 3)
    prm = default_params(epsilon=eps, max_ploidy=P)
    hap, bases, errors, info = ctx.phase_block(d, None, P, prm)
    assert hap.max() < P and len(hap) == 100000
    conf = np.zeros((P, P), np.int64)
    np.add.at(conf, (d.src, hap), 1)
    assert conf.max(axis=1).sum() == 100000, "truth partition not recovered"
    assert info["cells_beam"] >= int(d.ctx.L.fb_dfrags_nnz(d.handle))
    counts, same_q, diff_q, n_empty = ctx.bench_block_tables(d, P, hap, prm)
    cnt = (counts & np.uint64((1 << 62) - 1)).astype(np.int64)  # [P, n_pos, 4] in units of 2^-26
    present = (counts >> np.uint64(62)).astype(bool).any(axis=2)
    mx = cnt.max(axis=2)
    others = cnt.sum(axis=2) - mx
    score = 0.0
    for h in range(P):
        e_h = float(others[h].sum()) / 2.0 ** 26 + eps * int((present[h] & (mx[h] <= (1 << 26))).sum())
        score += e_h  # binom_vec.iter().map(|x| x.1).sum(), local_clustering.rs:97-99
    assert_f64_identical([info["opt_score"]], [score * -1.0], "optimize score vs host recomputation")
    d.free()
