"""GPU parity for row f1: update_hap_graph edge weights (graph_processing.rs:22-76) incl. distance_read_haplo (a4)."""
import numpy as np
import pytest

import oracle
from floria_b200 import api, default_params, synth
from floria_b200.frags import Frags

pytestmark = pytest.mark.gpu


def graph_from_blocks(c, prm, block_length):
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, block_length, block_length // 3, 0.0005)
    r = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=8)
    col_ptr, node_ptr, reads, nlo, nhi = [0], [0], [], [], []
    for j in range(r.n_blocks):
        if r.best_ploidy[j] == 0:
            continue
        ids = r.read_ids[r.read_ptr[j]:r.read_ptr[j + 1]]
        hp = r.hap[r.read_ptr[j]:r.read_ptr[j + 1]]
        for h in range(int(r.best_ploidy[j])):
            reads.extend(ids[hp == h].tolist())
            node_ptr.append(len(reads))
            nlo.append(int(lo[j]))
            nhi.append(int(hi[j]))
        col_ptr.append(len(node_ptr) - 1)
    return (np.array(col_ptr, np.uint64), np.array(node_ptr, np.uint64), np.array(reads, np.uint32),
            np.array(nlo, np.uint32), np.array(nhi, np.uint32))


@pytest.mark.parametrize("kind", ["long", "short"])
def test_update_hap_graph_matches_oracle(kind):
    ctx = api.Context(0)
    if kind == "long":
        c = synth.make_contig(61, 400, 360, 3, span_mean=60)
        bl = 10000
    else:
        c = synth.make_contig(62, 1500, 300, 2, paired_short=True, flip=0.01, qual_mode="short")
        bl = 800
    prm = default_params(epsilon=0.04, max_ploidy=4, block_length=bl)
    g = graph_from_blocks(c, prm, bl)
    assert len(g[0]) > 3
    o = oracle.update_hap_graph(c.frags, *g, prm)
    d = ctx.update_hap_graph(c.frags, *g, prm)
    assert o.sum() > 0
    assert np.array_equal(d, o), f"{int((d != o).sum())} of {len(o)} edge weights differ"
    ctx.close()


def test_update_hap_graph_ties_and_zero_quality():
    """exact ties (-> neither), q = 0 keys and 4 alleles exercise the consensus / tie planes"""
    ctx = api.Context(0)
    rng = np.random.default_rng(4)
    reads = []
    for i in range(160):
        span = int(rng.integers(2, 30))
        first = int(rng.integers(1, 80 - span))
        pos = list(range(first, first + span))
        reads.append((pos, rng.integers(0, 4, span), rng.choice([0, 0, 3, 20, 20, 40], span)))
    fr = Frags.from_reads(reads)
    # three columns of 3 / 2 / 3 random nodes over different endpoints
    col_ptr, node_ptr, nreads, nlo, nhi = [0], [0], [], [], []
    for col, (a, b, k) in enumerate([(1, 40, 3), (25, 60, 2), (45, 80, 3)]):
        cand = np.nonzero((fr.first <= b) & (fr.last >= a))[0]
        lab = rng.integers(0, k, len(cand))
        for h in range(k):
            nreads.extend(cand[lab == h].tolist())
            node_ptr.append(len(nreads))
            nlo.append(a)
            nhi.append(b)
        col_ptr.append(len(node_ptr) - 1)
    g = (np.array(col_ptr, np.uint64), np.array(node_ptr, np.uint64), np.array(nreads, np.uint32),
         np.array(nlo, np.uint32), np.array(nhi, np.uint32))
    prm = default_params(epsilon=0.04)
    o = oracle.update_hap_graph(fr, *g, prm)
    d = ctx.update_hap_graph(fr, *g, prm)
    assert np.array_equal(d, o)
    ctx.close()
