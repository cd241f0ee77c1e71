"""world_size-2 gloo test (CPU) of the multi-GPU host logic: static LPT work queue + variable-length gather."""
import os
import socket

import numpy as np
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
