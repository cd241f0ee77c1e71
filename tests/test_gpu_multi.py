"""Several devices behind the C-ABI (fb_init_multi / fb_phase_contigs): a contig list dealt to the devices by the
library's static LPT queue, every device's share batched into one call, per-contig results cut out again.  The results
must equal per-contig calls and the CPU oracle block for block.  With one GPU in the box the multi-context still runs
two contexts (two host threads, two streams, two allocator free lists) on device 0."""
import numpy as np
import pytest

import oracle
from floria_b200 import api, default_params, shard, synth
from test_gpu_parity import assert_f64_identical

pytestmark = pytest.mark.gpu


def _contigs(n, n_reads=260, n_snps=300):
    cs, blocks = [], []
    for k in range(n):
        c = synth.config5_contig(k, n_reads=n_reads + 17 * k, n_snps=n_snps + 11 * k, span_mean=60)
        cs.append(c.frags)
        blocks.append(api.get_range_with_lengths(c.snp_to_genome_pos, 8000, 8000 // 3, 0.0005))
    return cs, blocks


def _same(a, b, cells=True):
    assert np.array_equal(a.best_ploidy, b.best_ploidy)
    assert np.array_equal(a.ploidies_run, b.ploidies_run)
    assert_f64_identical(a.mec_vector.ravel(), b.mec_vector.ravel(), "mec_vector")
    assert_f64_identical(a.expected_errors.ravel(), b.expected_errors.ravel(), "expected_errors")
    assert np.array_equal(a.read_ptr, b.read_ptr) and np.array_equal(a.read_ids, b.read_ids)
    assert np.array_equal(a.hap, b.hap)
    if cells:
        assert a.cells == b.cells and np.array_equal(a.block_cells, b.block_cells)


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_phase_contigs_matches_per_contig_calls_and_oracle(devices):
    cs, blocks = _contigs(7)
    prm = default_params(epsilon=0.04, max_ploidy=4)
    m = api.MultiContext(devices)
    res, dev, ms = m.phase_contigs(cs, blocks, prm)
    assert len(res) == 7 and set(dev.tolist()) <= set(range(len(devices)))
    if len(devices) > 1:
        assert len(set(dev.tolist())) == len(devices), "every device gets a share of 7 contigs"
    owner = api.lpt_assign([api.contig_cost(f, len(b[0])) for f, b in zip(cs, blocks)], len(devices))
    assert np.array_equal(owner, dev)
    ctx = api.Context(0)
    for k in range(7):
        single = ctx.phase_blocks(cs[k], blocks[k][0], blocks[k][1], prm)
        _same(res[k], single)
        o = oracle.phase_blocks(cs[k], blocks[k][0], blocks[k][1], prm, n_threads=4)
        _same(res[k], o)
    # resident variant: same answer, no host->device copy of read data in the call
    d = m.upload(cs, blocks)
    res2, dev2, _ = m.phase_contigs_resident(d, prm)
    for k in range(7):
        _same(res2[k], res[k])
    d.free()
    ctx.close()
    m.close()


def test_two_contexts_on_one_device_do_not_share_freed_memory():
    """the caching allocator hands a freed block back to the same context only"""
    c = synth.make_contig(41, 300, 260, 3, span_mean=80)
    prm = default_params(epsilon=0.04, max_ploidy=3)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 8000, 2600, 0.0005)
    a, b = api.Context(0), api.Context(0)
    ra = a.phase_blocks(c.frags, lo, hi, prm)
    for _ in range(3):
        rb = b.phase_blocks(c.frags, lo, hi, prm)
        ra2 = a.phase_blocks(c.frags, lo, hi, prm)
        _same(ra, rb)
        _same(ra, ra2)
    a.close()
    b.close()


def test_lpt_assign_is_the_python_schedule():
    rng = np.random.default_rng(3)
    for n, bins in ((1, 1), (5, 8), (40, 3), (500, 8)):
        costs = rng.integers(1, 50, n).astype(np.float64)
        assert np.array_equal(api.lpt_assign(costs, bins), shard.lpt_assign(costs, bins))
