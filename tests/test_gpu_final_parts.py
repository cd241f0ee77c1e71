"""GPU parity for rows a14/a15: process_reads_for_final_parts (part_block_manip.rs:174-288) and get_hapq (:517-620)."""
import numpy as np
import pytest

import oracle
from floria_b200 import api, default_params, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def parts_from_blocks(c, prm, block_length):
    """haplosets as the pipeline would hand them over: every (block, haplotype) of the local phasing, with the
    block's SNP range; blocks overlap by 1/3 so many reads sit in several haplosets."""
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, block_length, block_length // 3, 0.0005)
    r = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=8)
    ptr, reads, rlo, rhi = [0], [], [], []
    for j in range(r.n_blocks):
        ids = r.read_ids[r.read_ptr[j]:r.read_ptr[j + 1]]
        hp = r.hap[r.read_ptr[j]:r.read_ptr[j + 1]]
        for h in range(int(r.best_ploidy[j])):
            sel = ids[hp == h]
            reads.extend(sel.tolist())
            ptr.append(len(reads))
            rlo.append(int(lo[j]))
            rhi.append(int(hi[j]))
    return np.array(ptr, np.uint64), np.array(reads, np.uint32), np.array(rlo, np.uint32), np.array(rhi, np.uint32)


def same_parts(g, o):
    assert g.n_parts == o.n_parts
    assert np.array_equal(g.part_ptr, o.part_ptr)
    assert np.array_equal(g.read_ids, o.read_ids)
    assert np.array_equal(g.range_lo, o.range_lo) and np.array_equal(g.range_hi, o.range_hi)


@pytest.mark.parametrize("eps", [0.03125, 0.04])
@pytest.mark.parametrize("kind", ["long", "short"])
def test_final_parts_and_hapq_match_oracle(ctx, eps, kind):
    if kind == "long":
        c = synth.make_contig(51, 400, 360, 3, span_mean=60)
        bl = 10000
    else:
        c = synth.make_contig(52, 1500, 300, 2, paired_short=True, flip=0.01, qual_mode="short")
        bl = 800
    prm = default_params(epsilon=eps, max_ploidy=4, block_length=bl)
    ptr, reads, rlo, rhi = parts_from_blocks(c, prm, bl)
    assert len(ptr) > 3
    o = oracle.process_reads_for_final_parts(c.frags, ptr, reads, rlo, rhi, prm)
    g = ctx.process_reads_for_final_parts(c.frags, ptr, reads, rlo, rhi, prm)
    same_parts(g, o)
    # every read that was in some haploset ends in exactly one (or is dropped by the break-splitting quirk)
    assert len(np.unique(g.read_ids)) == len(g.read_ids)
    oh, orel, oavg = oracle.get_hapq(c.frags, o.part_ptr, o.read_ids, o.range_lo, o.range_hi, c.snp_to_genome_pos, prm)
    gh, grel, gavg = ctx.get_hapq(c.frags, g.part_ptr, g.read_ids, g.range_lo, g.range_hi, c.snp_to_genome_pos, prm)
    assert np.array_equal(gh, oh)
    assert np.array_equal(grel.view(np.uint64), orel.view(np.uint64))
    assert np.float64(gavg).view(np.uint64) == np.float64(oavg).view(np.uint64)
    assert gh.max() <= 60


def test_final_parts_random_overlapping_sets(ctx):
    """random haplosets, including empty ones, single-read ones and reads shared by many sets"""
    c = synth.make_contig(53, 300, 200, 3, span_mean=40)
    rng = np.random.default_rng(3)
    prm = default_params(epsilon=0.04)
    ptr, reads, rlo, rhi = [0], [], [], []
    for k in range(12):
        a = int(rng.integers(1, 150))
        b = int(min(200, a + rng.integers(5, 90)))
        cand = np.nonzero((c.frags.first <= b) & (c.frags.last >= a))[0]
        take = cand[rng.random(len(cand)) < (0.0 if k == 5 else 0.5)]
        if k == 7:
            take = take[:1]
        reads.extend(take.tolist())
        ptr.append(len(reads))
        rlo.append(a)
        rhi.append(b)
    ptr, reads = np.array(ptr, np.uint64), np.array(reads, np.uint32)
    rlo, rhi = np.array(rlo, np.uint32), np.array(rhi, np.uint32)
    o = oracle.process_reads_for_final_parts(c.frags, ptr, reads, rlo, rhi, prm)
    g = ctx.process_reads_for_final_parts(c.frags, ptr, reads, rlo, rhi, prm)
    same_parts(g, o)
    oh, orel, oavg = oracle.get_hapq(c.frags, o.part_ptr, o.read_ids, o.range_lo, o.range_hi, c.snp_to_genome_pos, prm)
    gh, grel, gavg = ctx.get_hapq(c.frags, g.part_ptr, g.read_ids, g.range_lo, g.range_hi, c.snp_to_genome_pos, prm)
    assert np.array_equal(gh, oh)
    assert np.array_equal(grel.view(np.uint64), orel.view(np.uint64))


def test_resident_contig_pipeline_packs_once(ctx):
    """fb_frags_upload once, then phase_blocks_resident -> update_hap_graph_resident -> process_reads_for_final_parts_resident
    -> get_hapq_resident on the same device-resident contig: identical to the host-buffer entry points"""
    c = synth.make_contig(54, 400, 360, 3, span_mean=60)
    prm = default_params(epsilon=0.04, max_ploidy=4, block_length=10000)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005)
    d = ctx.upload(c.frags)
    r = ctx.phase_blocks_resident(d, lo, hi, prm)
    h = ctx.phase_blocks(c.frags, lo, hi, prm)
    assert np.array_equal(r.hap, h.hap) and np.array_equal(r.best_ploidy, h.best_ploidy)
    ptr, reads, rlo, rhi, col_ptr, node_lo, node_hi = [0], [], [], [], [0], [], []
    for j in range(r.n_blocks):
        ids = r.read_ids[r.read_ptr[j]:r.read_ptr[j + 1]]
        hp = r.hap[r.read_ptr[j]:r.read_ptr[j + 1]]
        for k in range(int(r.best_ploidy[j])):
            reads.extend(ids[hp == k].tolist())
            ptr.append(len(reads))
            rlo.append(int(lo[j]))
            rhi.append(int(hi[j]))
        col_ptr.append(len(ptr) - 1)
    w_res = ctx.update_hap_graph(d, col_ptr, ptr, reads, rlo, rhi, prm)
    w_host = ctx.update_hap_graph(c.frags, col_ptr, ptr, reads, rlo, rhi, prm)
    assert np.array_equal(w_res, w_host)
    p_res = ctx.process_reads_for_final_parts(d, ptr, reads, rlo, rhi, prm)
    p_host = ctx.process_reads_for_final_parts(c.frags, ptr, reads, rlo, rhi, prm)
    same_parts(p_res, p_host)
    q_res = ctx.get_hapq(d, p_res.part_ptr, p_res.read_ids, p_res.range_lo, p_res.range_hi, c.snp_to_genome_pos, prm)
    q_host = ctx.get_hapq(c.frags, p_host.part_ptr, p_host.read_ids, p_host.range_lo, p_host.range_hi, c.snp_to_genome_pos, prm)
    assert np.array_equal(q_res[0], q_host[0]) and np.array_equal(q_res[1].view(np.uint64), q_host[1].view(np.uint64))
    assert q_res[2] == q_host[2]
    d.free()
