"""The device-side generator of the configs[2] roofline block must produce exactly the cells of the numpy generator, and
the timed sweep/hist loop must leave results identical to the fine-grained entry points."""
import numpy as np
import pytest

from floria_b200 import api, default_params, synth

pytestmark = pytest.mark.gpu


def test_device_synth_matches_numpy_generator():
    ctx = api.Context(0)
    R, S, P = 300, 1000, 4
    d = ctx.bench_synth_dense(R, S, P, 3)
    c = synth.make_contig(3, R, S, P, full_span=True, flip=0.04, qual_mode="long")
    h = ctx.upload(c.frags)
    q1, a1, p1 = ctx.download_planes(d)
    q2, a2, p2 = ctx.download_planes(h)
    assert np.array_equal(p1, p2) and np.array_equal(a1, a2) and np.array_equal(q1, q2)
    assert np.array_equal(d.src, c.read_hap)
    prm = default_params(epsilon=0.04)
    sw, hs, cells = ctx.bench_sweep_hist(d, P, d.src, prm, 2)
    assert cells == c.frags.nnz and (sw > 0).all() and (hs > 0).all()
    d.free()
    h.free()
    ctx.close()


@pytest.mark.parametrize("shape", [(2000, 3000, 3), (100000, 50000, 4)], ids=["small", "configs2_full_size"])
def test_hist_sweep_checksums_at_full_size(shape):
    """BASELINE.json configs[2] (100k full-span reads x 50k SNPs, ploidy 4: 4.9e9 stored cells) is far beyond what the
    CPU oracle can run, so the two bandwidth-bound kernels are tied together there by size-independent identities:
      * linearity: the ploidy-P count table summed over haplotypes equals the table of the merged partition;
      * checksum of checksums: for every haplotype h the total of its count table equals the sum over ITS reads of
        same(r, h) + diff(r, h) from the SCORE sweep (a read never sees an empty position in its own haplotype);
      * with one haplotype holding every read, same + diff of all reads equals the whole table.
    The small shape runs the same identities where the oracle-checked fine-grained entry points also agree."""
    R, S, P = shape
    ctx = api.Context(0)
    d = ctx.bench_synth_dense(R, S, P, 3)
    prm = default_params(epsilon=0.04)
    MASK = np.uint64((1 << 62) - 1)
    cP, sqP, dqP, neP = ctx.bench_block_tables(d, P, d.src, prm)
    c1, sq1, dq1, ne1 = ctx.bench_block_tables(d, 1, np.zeros(R, np.uint8), prm)
    vP = (cP & MASK).astype(np.int64)
    v1 = (c1 & MASK).astype(np.int64)
    assert np.array_equal(vP.sum(axis=0), v1[0]), "histogram is not linear in the partition"
    keyP = (cP >> np.uint64(62)) & np.uint64(1)
    key1 = (c1 >> np.uint64(62)) & np.uint64(1)
    assert np.array_equal(keyP.max(axis=0), key1[0]), "allele key sets differ"
    for h in range(P):
        mine = d.src == h
        assert int(vP[h].sum()) == int((sqP[mine, h] + dqP[mine, h]).sum()), f"haplotype {h}: table total != sweep total"
        assert not neP[mine, h].any(), "a read met an empty position in its own haplotype"
    assert int(v1.sum()) == int((sq1[:, 0] + dq1[:, 0]).sum())
    assert not ne1.any()
    # every read agrees at least as well with its source haplotype as with any other (4 % flips, abundant coverage)
    best = np.argmin(dqP, axis=1)
    assert (best == d.src).mean() > 0.99
    if R <= 5000:
        c = synth.make_contig(3, R, S, P, full_span=True, flip=0.04, qual_mode="long")
        sel = np.arange(R, dtype=np.uint32)
        _, _, sq, dq, ne = ctx.score_reads(c.frags, sel, d.src, P, prm)
        assert np.array_equal(sq, sqP) and np.array_equal(dq, dqP) and np.array_equal(ne, neP)
    d.free()
    ctx.close()
