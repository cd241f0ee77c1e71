"""BASELINE.json configs[0]: the reference's own quick-start input (tests/test_long.bam + tests/test.vcf on MN-03.fa,
README.md:85; 3 Klebsiella strains).  The fragments were extracted once in the build container by
tools/extract_frags.py (a restatement of get_vcf_profile / alignment_passed_check / frag_from_record; declared
differences: no supplementary merging, no local re-alignment) and committed as tests/golden/config0_long_frags.npz,
because the reference tree does not exist on the GPU box.  The hot path on REAL long reads (indels, uneven coverage,
real base qualities): CUDA path == CPU oracle, bit for bit, for the local phasing, the final read refinement, HAPQ
and the block-graph edge weights."""
import os

import numpy as np
import pytest

import oracle
from floria_b200 import api, default_params
from floria_b200.frags import Frags

PATH = os.path.join(os.path.dirname(__file__), "golden", "config0_long_frags.npz")


def load():
    g = np.load(PATH)
    fr = Frags(g["row_ptr"], g["pos"], g["allele"], g["qual"])
    L = int(g["block_length"])  # auto -l: max(p66 read length, 500), file_reader.rs:821
    lo, hi = oracle.get_range_with_lengths(g["snp_to_genome_pos"], L, L // 3, 0.0005)
    return g, fr, L, lo, hi


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_fixture_shape_and_oracle_phasing():
    g, fr, L, lo, hi = load()
    assert fr.n_reads == 1162 and fr.nnz == 75943 and len(g["snp_to_genome_pos"]) == 954
    assert fr.is_sorted() and int(fr.allele.max()) <= 1  # test.vcf is all biallelic SNPs
    assert len(lo) == 17
    prm = default_params(epsilon=0.04, max_ploidy=3, block_length=L)  # configs[0]: ploidy = 3
    r = oracle.phase_blocks(fr, lo, hi, prm, n_threads=4)
    # three strains: most blocks need all three haplotypes, none collapses to one
    assert (r.best_ploidy == 3).sum() >= 8 and r.best_ploidy.min() >= 2
    for j in range(r.n_blocks):
        h = r.hap[int(r.read_ptr[j]):int(r.read_ptr[j + 1])]
        assert len(h) > 0 and h.max() < r.best_ploidy[j]


@pytest.mark.gpu
@pytest.mark.parametrize("max_ploidy", [3, 5])
def test_config0_cuda_path_matches_oracle(max_ploidy):
    g, fr, L, lo, hi = load()
    prm = default_params(epsilon=0.04, max_ploidy=max_ploidy, block_length=L)
    ctx = api.Context(0)
    try:
        o = oracle.phase_blocks(fr, lo, hi, prm, n_threads=8)
        r = ctx.phase_blocks(fr, lo, hi, prm)
        assert np.array_equal(r.best_ploidy, o.best_ploidy) and np.array_equal(r.ploidies_run, o.ploidies_run)
        assert np.array_equal(r.read_ids, o.read_ids) and np.array_equal(r.hap, o.hap)
        assert np.array_equal(bits(r.mec_vector), bits(o.mec_vector))
        assert np.array_equal(bits(r.expected_errors), bits(o.expected_errors))
        assert r.cells == o.cells
        # haplosets of the local phasing -> final read refinement -> HAPQ (rows a14 / a15)
        ptr, reads, rlo, rhi = [0], [], [], []
        for j in range(o.n_blocks):
            ids = o.read_ids[int(o.read_ptr[j]):int(o.read_ptr[j + 1])]
            hp = o.hap[int(o.read_ptr[j]):int(o.read_ptr[j + 1])]
            for h in range(int(o.best_ploidy[j])):
                reads.extend(ids[hp == h].tolist())
                ptr.append(len(reads))
                rlo.append(int(lo[j]))
                rhi.append(int(hi[j]))
        ptr, reads = np.array(ptr, np.uint64), np.array(reads, np.uint32)
        rlo, rhi = np.array(rlo, np.uint32), np.array(rhi, np.uint32)
        op = oracle.process_reads_for_final_parts(fr, ptr, reads, rlo, rhi, prm)
        gp = ctx.process_reads_for_final_parts(fr, ptr, reads, rlo, rhi, prm)
        assert gp.n_parts == op.n_parts and np.array_equal(gp.part_ptr, op.part_ptr)
        assert np.array_equal(gp.read_ids, op.read_ids)
        assert np.array_equal(gp.range_lo, op.range_lo) and np.array_equal(gp.range_hi, op.range_hi)
        oh, orel, oavg = oracle.get_hapq(fr, op.part_ptr, op.read_ids, op.range_lo, op.range_hi, g["snp_to_genome_pos"], prm)
        gh, grel, gavg = ctx.get_hapq(fr, gp.part_ptr, gp.read_ids, gp.range_lo, gp.range_hi, g["snp_to_genome_pos"], prm)
        assert np.array_equal(gh, oh) and np.array_equal(bits(grel), bits(orel)) and bits(gavg) == bits(oavg)
    finally:
        ctx.close()
