"""Golden fixtures (tests/golden/*.npz, written by tools/make_golden.py from the CPU oracle; the reference ships no
expected outputs for this path, see the script's header).  CPU: the oracle still reproduces them bit for bit.
GPU: the CUDA path, through the C-ABI, reproduces them bit for bit (f64 compared as bits, integers exactly)."""
import glob
import os

import numpy as np
import pytest

import oracle
from floria_b200 import api, default_params
from floria_b200.frags import Frags

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not any(t in os.path.basename(p) for t in ("config0", "config3_", "config5_")))  # those have their own tests
P = 3


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def load(path):
    g = np.load(path)
    fr = Frags(g["row_ptr"], g["pos"], g["allele"], g["qual"])
    prm = default_params(epsilon=float(g["epsilon"]), max_ploidy=int(g["max_ploidy"]), block_length=int(g["block_length"]))
    return g, fr, prm


def check_block_results(g, r):
    assert np.array_equal(r.best_ploidy, g["best_ploidy"])
    assert np.array_equal(r.ploidies_run, g["ploidies_run"])
    assert np.array_equal(r.read_ptr, g["read_ptr"]) and np.array_equal(r.read_ids, g["read_ids"])
    assert np.array_equal(r.hap, g["block_hap"]), "read -> haplotype assignments differ"
    assert np.array_equal(bits(r.mec_vector), bits(g["mec_vector"])), "MEC f64 bits differ"
    assert np.array_equal(bits(r.expected_errors), bits(g["expected_errors"]))
    assert [r.cells_sweep, r.cells_hist, r.cells_beam] == g["cells"].tolist()


def test_fixtures_exist():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g, fr, prm = load(path)
    lo, hi = oracle.get_range_with_lengths(g["snp_to_genome_pos"], int(g["block_length"]), int(g["block_length"]) // 3, 0.0005)
    assert np.array_equal(lo, g["blk_lo"]) and np.array_equal(hi, g["blk_hi"])
    check_block_results(g, oracle.phase_blocks(fr, lo, hi, prm, n_threads=2))
    sel = np.arange(fr.n_reads, dtype=np.uint32)
    same, diff = oracle.score_reads(fr, sel, g["part_hap"], P, prm)
    assert np.array_equal(bits(same), bits(g["same"])) and np.array_equal(bits(diff), bits(g["diff"]))
    counts, keymask = oracle.hap_block_from_partition(fr, sel, g["part_hap"], P, 1, prm, 1, g["counts"].shape[1])
    assert np.array_equal(bits(counts), bits(g["counts"])) and np.array_equal(keymask, g["keymask"])
    b, e = oracle.get_mec_stats_epsilon(fr, sel, g["part_hap"], P, 1, prm)
    assert np.array_equal(bits(b), bits(g["bases"])) and np.array_equal(bits(e), bits(g["errors"]))
    hap, score, _ = oracle.beam_search_phasing(fr, sel, P, prm)
    assert np.array_equal(hap, g["beam_hap"]) and bits(score) == bits(g["beam_score"])
    ohap, oscore, rounds = oracle.optimize_clustering(fr, sel, hap, P, prm)
    assert np.array_equal(ohap, g["opt_hap"]) and bits(oscore) == bits(g["opt_score"]) and rounds == int(g["opt_rounds"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_path_reproduces_golden(path):
    g, fr, prm = load(path)
    ctx = api.Context(0)
    try:
        lo, hi = api.get_range_with_lengths(g["snp_to_genome_pos"], int(g["block_length"]), int(g["block_length"]) // 3, 0.0005)
        assert np.array_equal(lo, g["blk_lo"]) and np.array_equal(hi, g["blk_hi"])
        check_block_results(g, ctx.phase_blocks(fr, lo, hi, prm))
        sel = np.arange(fr.n_reads, dtype=np.uint32)
        same, diff, _, _, _ = ctx.score_reads(fr, sel, g["part_hap"], P, prm)
        assert np.array_equal(bits(same), bits(g["same"])) and np.array_equal(bits(diff), bits(g["diff"]))
        counts, keymask = ctx.hap_block_from_partition(fr, sel, g["part_hap"], P, 1, prm, 1, g["counts"].shape[1])
        assert np.array_equal(bits(counts), bits(g["counts"])) and np.array_equal(keymask, g["keymask"])
        b, e = ctx.get_mec_stats_epsilon(fr, sel, g["part_hap"], P, 1, prm)
        assert np.array_equal(bits(b), bits(g["bases"])) and np.array_equal(bits(e), bits(g["errors"]))
        b, e = ctx.get_mec_stats_epsilon(fr, sel, g["part_hap"], P, 0, prm)
        assert np.array_equal(bits(b), bits(g["bases_nophred"])) and np.array_equal(bits(e), bits(g["errors_nophred"]))
        hap, score = ctx.beam_search_phasing(fr, sel, P, prm)[:2]
        assert np.array_equal(hap, g["beam_hap"]) and bits(score) == bits(g["beam_score"])
        ohap, oscore = ctx.optimize_clustering(fr, sel, hap, P, prm)[:2]
        assert np.array_equal(ohap, g["opt_hap"]) and bits(oscore) == bits(g["opt_score"])
    finally:
        ctx.close()
