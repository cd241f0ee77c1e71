"""Multi-GPU plumbing: contigs / SNP blocks are independent units (graph_processing.rs:345-362 pushes them under a
Mutex and re-sorts by index), so ranks take disjoint units of a static work queue and only the final partition
records are gathered (torch.distributed: NCCL on GPUs, gloo in the CPU tests).  No data-path collective."""
import numpy as np
import torch
import torch.distributed as dist


def lpt_assign(costs, n_ranks):
    """Static longest-processing-time-first assignment of units to ranks.  Returns rank index per unit; deterministic
    (ties broken by unit index) so every rank computes the same schedule without communication."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(len(costs)), -costs))
    load = np.zeros(n_ranks)
    owner = np.zeros(len(costs), dtype=np.int64)
    for u in order:
        r = int(np.argmin(load))  # first minimum
        owner[u] = r
        load[r] += costs[u]
    return owner


def block_costs(frags, blk_lo, blk_hi, max_ploidy):
    """cost estimate of a block: sum_p (beam * p * nnz) ~ nnz * p(p+1)/2"""
    first = frags.first.astype(np.int64)
    last = frags.last.astype(np.int64)
    nnz = (frags.row_ptr[1:] - frags.row_ptr[:-1]).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(nnz)])
    out = np.zeros(len(blk_lo))
    for j, (a, b) in enumerate(zip(blk_lo, blk_hi)):
        hi = np.searchsorted(first, int(b), side="right")
        sel = np.nonzero(last[:hi] >= int(a))[0]
        out[j] = float(nnz[sel].sum()) * max_ploidy * (max_ploidy + 1) / 2
    return out


def concat_contigs(contigs, blocks, gap=64):
    """Batch several contigs into ONE fb_phase_blocks call: contig k's SNP positions are shifted by the total length of
    the contigs before it (+ `gap`), so reads stay sorted by Frag::cmp, no read of one contig can overlap a block of
    another, and every block is phased exactly as in a per-contig call (blocks are independent units,
    graph_processing.rs:345-362).  This is how a rank pushes its whole share of a many-contig work queue through the
    GPU at once instead of one latency-bound call per contig.
    contigs: list of Frags; blocks: list of (blk_lo, blk_hi) arrays (1-based SNP ranges per contig).
    Returns (frags, blk_lo, blk_hi, block_contig, read_offset[k], pos_offset[k])."""
    from .frags import Frags

    row_ptr, pos, allele, qual, first, last = [np.zeros(1, np.uint64)], [], [], [], [], []
    lo_all, hi_all, owner = [], [], []
    read_off, pos_off = [], []
    p0, c0, r0 = 0, 0, 0
    for k, (fr, (lo, hi)) in enumerate(zip(contigs, blocks)):
        read_off.append(r0)
        pos_off.append(p0)
        row_ptr.append(fr.row_ptr[1:] + np.uint64(c0))
        pos.append(fr.pos + np.uint32(p0))
        allele.append(fr.allele)
        qual.append(fr.qual)
        first.append(fr.first + np.uint32(p0))
        last.append(fr.last + np.uint32(p0))
        lo_all.append(np.asarray(lo, np.uint32) + np.uint32(p0))
        hi_all.append(np.asarray(hi, np.uint32) + np.uint32(p0))
        owner.append(np.full(len(lo), k, np.int64))
        span = max(int(fr.last.max()) if fr.n_reads else 0, int(np.max(hi)) if len(hi) else 0)
        p0 += span + gap
        c0 += fr.nnz
        r0 += fr.n_reads
        if p0 >= 2 ** 32 - 2 ** 20:
            raise ValueError("batched contigs exceed the 32-bit SNP index space")
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    frags = Frags(cat(row_ptr, np.uint64), cat(pos, np.uint32), cat(allele, np.uint8), cat(qual, np.uint8),
                  cat(first, np.uint32), cat(last, np.uint32))
    return (frags, cat(lo_all, np.uint32), cat(hi_all, np.uint32), cat(owner, np.int64), np.array(read_off, np.int64),
            np.array(pos_off, np.int64))


class GatheredRecords:
    """The partition records of every rank as received on the destination rank: one byte buffer per rank, parsed on
    demand (`to_dict`).  Layout of a buffer: int64 n_units, unit_ids[n], best_ploidy[n], read_ptr[n+1], then uint32
    read_ids[total] and uint8 hap[total]."""

    def __init__(self, bufs):
        self.bufs = [np.ascontiguousarray(b) for b in bufs]

    @property
    def nbytes(self):
        return int(sum(len(b) for b in self.bufs))

    def per_rank(self):
        """[(unit_ids, best_ploidy, read_ptr, read_ids, hap)] as array views, one tuple per rank"""
        out = []
        for b in self.bufs:
            nu = int(b[:8].view(np.int64)[0])
            off = 8
            uids = b[off:off + 8 * nu].view(np.int64); off += 8 * nu
            bp = b[off:off + 8 * nu].view(np.int64); off += 8 * nu
            rp = b[off:off + 8 * (nu + 1)].view(np.int64); off += 8 * (nu + 1)
            tot = int(rp[-1]) if nu else 0
            rid = b[off:off + 4 * tot].view(np.uint32); off += 4 * tot
            out.append((uids, bp, rp, rid, b[off:off + tot]))
        return out

    def to_dict(self):
        """unit_id -> (best_ploidy, read_ids, hap)"""
        res = {}
        for uids, bp, rp, rid, hp in self.per_rank():
            for k in range(len(uids)):
                res[int(uids[k])] = (int(bp[k]), rid[rp[k]:rp[k + 1]].copy(), hp[rp[k]:rp[k + 1]].copy())
        return res


def gather_records(unit_ids, read_ptr, read_ids, hap, best_ploidy, device, dst=0, lazy=False):
    """Variable-length gather of per-unit partition records to rank `dst`.
    Every rank passes the records of the units it owns; returns on dst a dict unit_id -> (best_ploidy, read_ids, hap)
    (a GatheredRecords with lazy=True: the received buffers, parsed on demand), None elsewhere.
    One all_gather of sizes + one padded gather of the payload into a single [world, max] tensor + one copy to the host."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    unit_ids = np.asarray(unit_ids, dtype=np.int64)
    read_ptr = np.asarray(read_ptr, dtype=np.int64)
    head = np.concatenate([[len(unit_ids)], unit_ids, np.asarray(best_ploidy, dtype=np.int64), read_ptr]).astype(np.int64)
    payload = np.concatenate([head.view(np.uint8), np.asarray(read_ids, np.uint32).view(np.uint8),
                              np.asarray(hap, np.uint8)])
    if world == 1:
        bufs = [payload]
    else:
        n = torch.tensor([len(payload)], dtype=torch.int64, device=device)
        sizes = torch.zeros(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(sizes, n)
        sizes = sizes.cpu().tolist()
        mx = max(sizes)
        rec = torch.empty(mx, dtype=torch.uint8, device=device)
        rec[: len(payload)] = torch.from_numpy(payload).to(device, non_blocking=True)
        out = torch.empty((world, mx), dtype=torch.uint8, device=device) if rank == dst else None
        dist.gather(rec, list(out.unbind(0)) if rank == dst else None, dst=dst)
        if rank != dst:
            return None
        host = out.cpu().numpy()
        bufs = [host[r, :sizes[r]] for r in range(world)]
    g = GatheredRecords(bufs)
    return g if lazy else g.to_dict()
