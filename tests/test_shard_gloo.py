"""world_size-2 gloo test (CPU) of the multi-GPU host logic: static LPT work queue + variable-length gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from floria_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    n_units = 11
    costs = rng.integers(1, 100, n_units)
    owner = shard.lpt_assign(costs, world)
    mine = np.nonzero(owner == rank)[0]
    # fake partition records whose content is a function of the unit id
    read_ptr, read_ids, hap, bp = [0], [], [], []
    for u in mine:
        n = 3 + int(u) % 5
        read_ids.extend(range(100 * int(u), 100 * int(u) + n))
        hap.extend([(int(u) + k) % 3 for k in range(n)])
        read_ptr.append(len(read_ids))
        bp.append(1 + int(u) % 4)
    res = shard.gather_records(mine, read_ptr, read_ids, hap, bp, torch.device("cpu"), dst=0)
    if rank == 0:
        ok = sorted(res) == list(range(n_units))
        for u in range(n_units):
            n = 3 + u % 5
            b, rid, hp = res[u]
            ok &= b == 1 + u % 4 and list(rid) == list(range(100 * u, 100 * u + n))
            ok &= list(hp) == [(u + k) % 3 for k in range(n)]
        q.put(bool(ok))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_is_balanced_and_deterministic():
    costs = [9, 7, 6, 5, 4, 3, 2, 1, 1]
    o = shard.lpt_assign(costs, 3)
    loads = [sum(c for c, r in zip(costs, o) if r == k) for k in range(3)]
    assert max(loads) - min(loads) <= 2 and sorted(set(o.tolist())) == [0, 1, 2]
    assert np.array_equal(o, shard.lpt_assign(costs, 3))
    assert np.array_equal(shard.lpt_assign([5, 5, 5], 1), [0, 0, 0])


def test_gather_records_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_gather_records_single_process():
    res = shard.gather_records([4, 2], [0, 2, 3], [10, 11, 20], [0, 1, 0], [2, 1], torch.device("cpu"))
    assert res[4][0] == 2 and list(res[4][1]) == [10, 11] and list(res[2][2]) == [0]


def test_gather_records_lazy_view():
    """lazy=True hands back the received buffers (what bench.py times); per_rank() exposes them as array views and
    to_dict() parses them into the same mapping as the eager call"""
    g = shard.gather_records([4, 2], [0, 2, 3], [10, 11, 20], [0, 1, 0], [2, 1], torch.device("cpu"), lazy=True)
    assert isinstance(g, shard.GatheredRecords) and g.nbytes == 8 * (1 + 2 + 2 + 3) + 4 * 3 + 3
    (uids, bp, rp, rid, hp), = g.per_rank()
    assert uids.tolist() == [4, 2] and bp.tolist() == [2, 1] and rp.tolist() == [0, 2, 3]
    assert rid.tolist() == [10, 11, 20] and hp.tolist() == [0, 1, 0]
    d = g.to_dict()
    assert d[4][0] == 2 and list(d[4][1]) == [10, 11] and list(d[2][2]) == [0]


def _metagenome(n_contigs, n_reads, n_snps):
    from floria_b200 import api, synth

    contigs, blocks = [], []
    for k in range(n_contigs):
        c = synth.config5_contig(k, n_reads=n_reads, n_snps=n_snps, span_mean=40)
        contigs.append(c.frags)
        blocks.append(api.get_range_with_lengths(c.snp_to_genome_pos, 4000, 4000 // 3, 0.0005))
    return contigs, blocks


def test_concat_contigs_is_equivalent_to_per_contig_calls_oracle():
    """configs[4]-shaped input (many small contigs, mixed ploidy 2..6) at reduced size: one batched call over the
    concatenated contigs gives, block for block, the per-contig results (CPU oracle on both sides)."""
    import oracle
    from floria_b200 import default_params, shard

    contigs, blocks = _metagenome(4, 90, 80)
    prm = default_params(epsilon=0.04, max_ploidy=4)
    fr, lo, hi, owner, read_off, _ = shard.concat_contigs(contigs, blocks)
    assert fr.is_sorted()
    big = oracle.phase_blocks(fr, lo, hi, prm, n_threads=4)
    j = 0
    for k, (c, (clo, chi)) in enumerate(zip(contigs, blocks)):
        one = oracle.phase_blocks(c, clo, chi, prm, n_threads=2)
        for b in range(one.n_blocks):
            assert owner[j] == k
            assert big.best_ploidy[j] == one.best_ploidy[b]
            a0, a1 = int(big.read_ptr[j]), int(big.read_ptr[j + 1])
            b0, b1 = int(one.read_ptr[b]), int(one.read_ptr[b + 1])
            assert np.array_equal(big.read_ids[a0:a1] - read_off[k], one.read_ids[b0:b1])
            assert np.array_equal(big.hap[a0:a1], one.hap[b0:b1])
            assert np.array_equal(big.mec_vector[j].view(np.uint64), one.mec_vector[b].view(np.uint64))
            j += 1
    assert j == big.n_blocks


@pytest.mark.gpu
def test_batched_metagenome_matches_oracle_and_per_contig_calls_gpu():
    """The same equivalence through the CUDA path, plus bit-exact agreement of the batched call with the oracle."""
    import oracle
    from floria_b200 import api, default_params, shard

    contigs, blocks = _metagenome(6, 120, 100)
    prm = default_params(epsilon=0.04, max_ploidy=6)
    fr, lo, hi, owner, read_off, _ = shard.concat_contigs(contigs, blocks)
    ctx = api.Context(0)
    try:
        big = ctx.phase_blocks(fr, lo, hi, prm)
        ref = oracle.phase_blocks(fr, lo, hi, prm, n_threads=8)
        assert np.array_equal(big.best_ploidy, ref.best_ploidy) and np.array_equal(big.hap, ref.hap)
        assert np.array_equal(big.mec_vector.view(np.uint64), ref.mec_vector.view(np.uint64))
        assert big.cells == ref.cells
        j = 0
        for k, (c, (clo, chi)) in enumerate(zip(contigs, blocks)):
            one = ctx.phase_blocks(c, clo, chi, prm)
            for b in range(one.n_blocks):
                a0, a1 = int(big.read_ptr[j]), int(big.read_ptr[j + 1])
                b0, b1 = int(one.read_ptr[b]), int(one.read_ptr[b + 1])
                assert big.best_ploidy[j] == one.best_ploidy[b]
                assert np.array_equal(big.hap[a0:a1], one.hap[b0:b1])
                j += 1
        assert j == big.n_blocks
    finally:
        ctx.close()
