#!/usr/bin/env python
"""Writes the golden fixtures under tests/golden/.

The reference (Rust) cannot be built or run in this image and ships no expected outputs for this path (SURVEY.md §4,
§8c), so these vectors are produced by the CPU ORACLE (oracle/floria_oracle.cpp, the C++ restatement of the reference
functions) on small seeded inputs.  They do not pin the oracle to the reference ("parity unpinned"); they
pin (a) the oracle against silent drift (compiler, libm, refactors) and (b) the CUDA path against a committed answer that
travels to the GPU box.  Every fixture stores its INPUTS (CSR arrays) next to the outputs, so it is self-contained.

    python tools/make_golden.py            # regenerate all fixtures
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from floria_b200 import api, default_params, synth  # noqa: E402
from floria_b200.frags import Frags  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def edge_frags(seed, n_reads=70, n_snps=60):
    """ragged reads with holes, q = 0 / 255, up to 4 alleles, single-cell reads"""
    rng = np.random.default_rng(seed)
    reads = []
    for _ in range(n_reads):
        span = int(rng.integers(1, 30))
        first = int(rng.integers(1, n_snps - span + 2))
        pos = [p for p in range(first, first + span) if rng.random() < 0.8 or p in (first, first + span - 1)]
        reads.append((pos, rng.integers(0, 4, len(pos)), rng.choice([0, 1, 2, 3, 10, 20, 30, 40, 60, 93, 255], len(pos))))
    return Frags.from_reads(reads)


def one_case(name, frags, g2p, eps, max_ploidy, block_length):
    prm = default_params(epsilon=eps, max_ploidy=max_ploidy, block_length=block_length)
    lo, hi = api.get_range_with_lengths(g2p, block_length, block_length // 3, 0.0005)
    ph = oracle.phase_blocks(frags, lo, hi, prm, n_threads=4)
    rng = np.random.default_rng(7)
    sel = np.arange(frags.n_reads, dtype=np.uint32)
    P = 3
    hap = rng.integers(0, P, frags.n_reads).astype(np.uint8)
    same, diff = oracle.score_reads(frags, sel, hap, P, prm)
    n_pos = int(frags.last.max())
    counts, keymask = oracle.hap_block_from_partition(frags, sel, hap, P, 1, prm, 1, n_pos)
    bases, errors = oracle.get_mec_stats_epsilon(frags, sel, hap, P, 1, prm)
    bases_np, errors_np = oracle.get_mec_stats_epsilon(frags, sel, hap, P, 0, prm)
    bs_hap, bs_score, _ = oracle.beam_search_phasing(frags, sel, P, prm)
    op_hap, op_score, op_rounds = oracle.optimize_clustering(frags, sel, bs_hap, P, prm)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        row_ptr=frags.row_ptr, pos=frags.pos, allele=frags.allele, qual=frags.qual, snp_to_genome_pos=g2p,
        epsilon=np.float64(eps), max_ploidy=np.uint32(max_ploidy), block_length=np.uint32(block_length),
        blk_lo=lo, blk_hi=hi, best_ploidy=ph.best_ploidy, ploidies_run=ph.ploidies_run, mec_vector=ph.mec_vector,
        expected_errors=ph.expected_errors, read_ptr=ph.read_ptr, read_ids=ph.read_ids, block_hap=ph.hap,
        cells=np.array([ph.cells_sweep, ph.cells_hist, ph.cells_beam], np.uint64),
        part_hap=hap, same=same, diff=diff, counts=counts, keymask=keymask, bases=bases, errors=errors,
        bases_nophred=bases_np, errors_nophred=errors_np, beam_hap=bs_hap, beam_score=np.float64(bs_score),
        opt_hap=op_hap, opt_score=np.float64(op_score), opt_rounds=np.uint32(op_rounds))
    print(f"{name}: {frags.n_reads} reads, {frags.nnz} cells, {len(lo)} blocks, best ploidy {ph.best_ploidy.tolist()}")


def main():
    os.makedirs(OUT, exist_ok=True)
    c = synth.make_contig(71, 160, 140, 3, span_mean=50)
    for eps, tag in ((0.04, "eps004"), (0.03125, "dyadic")):
        one_case(f"long_p3_{tag}", c.frags, c.snp_to_genome_pos, eps, 4, 5000)
    s = synth.make_contig(72, 400, 150, 2, paired_short=True, flip=0.01, qual_mode="short")
    one_case("short_p2_eps001", s.frags, s.snp_to_genome_pos, 0.01, 3, 3000)
    e = edge_frags(73)
    g2p = (100 * np.arange(int(e.last.max()) + 1, dtype=np.uint64) + 7)
    one_case("edge_eps004", e, g2p, 0.04, 3, 2000)


if __name__ == "__main__":
    main()
