#!/usr/bin/env python
"""bench.py — read x SNP cells scored per second through the B200 hot path (BASELINE.json metric).

A step = one pass of the hot path over one contig: fb_phase_blocks over every SNP block of the contig
(per block: read selection, ploidy loop of beam_search_phasing -> optimize_clustering -> no-phred MEC stats,
stopping rule), i.e. what floria's generate_hap_graph par_iter computes (graph_processing.rs:345-362).
Workload at N=1: BASELINE.json configs[1] (1 contig, 10k long-read frags x 5k SNPs, ploidy 2).  With N ranks every
rank phases its own contig of that shape (contigs shard independently: weak scaling, no data-path collective) and
rank 0 gathers the partition records over NCCL inside the timed region.

  value : cells/s with the contig already packed in HBM (fb_phase_blocks_resident)
  e2e   : cells/s through fb_phase_blocks with HOST (pinned) CSR buffers: H2D of the reads + packing + compute +
          D2H of the partition records, every step
  --impl reference : the CPU restatement of the reference path (oracle, all host threads) on a bounded sample of the
          same workload (the reference itself is Rust and cannot be built in this image).
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_CELL = 1.375  # 2-bit allele + 8-bit qual + 1-bit presence


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason samples of this rank's GPU (NVML, in process), taken BETWEEN the steps of the timed
    loop: right after a step's synchronize, while the GPU is still at its load clocks, and outside every step's CUDA-event
    window.  Two concurrent samplers were tried first (a polling thread, then a polling child process): on some hosts
    their NVML queries stalled the CUDA calls of the measured process for tens of milliseconds (one run: 33.7 ms per step
    with the sampler against 18.5 ms without, identical kernel times), so nothing polls while a step runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.sm, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        try:
            import pynvml

            vis = os.environ.get("CUDA_VISIBLE_DEVICES")  # remaps CUDA ordinals, not NVML indices
            if vis:
                index = int(vis.split(",")[index])
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def sample(self):
        if not self.ok:
            return
        try:
            self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:  # pragma: no cover
            pass

    def result(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "sampled": "between timed steps", "reasons": sorted(self.reasons)}


def make_workload(rank):
    from floria_b200 import api, default_params, synth

    # configs[1]; ranks > 0 get their own contig of the same shape (distinct seed)
    c = synth.make_contig(2 + 1000 * rank, 10000, 5000, 2, span_mean=500, flip=0.04, qual_mode="long")
    prm = default_params(epsilon=0.04, max_ploidy=2, block_length=10000)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005)
    desc = {"workload": "configs[1]: synthetic 1 contig, 10k long-read frags x 5k SNPs, ploidy 2 "
                        "(fb_phase_blocks over all SNP blocks, ploidy loop 1..2, beam 10, eps 0.04)",
            "n_reads": int(c.frags.n_reads), "n_snps": 5000, "stored_cells": int(c.frags.nnz),
            "n_blocks": int(len(lo)), "max_ploidy": 2, "beam": 10, "epsilon": 0.04}
    return c, prm, lo, hi, desc


def c3_roofline(ctx, peak, n_reads=100000, n_snps=50000, ploidy=4, iters=5):
    """BASELINE.json configs[2]: the 100k x 50k full-span block (5e9 cells, 6.9 GB packed, >> L2), ploidy 4.
    Times the two bandwidth-bound kernels of one optimize_clustering round: k_sweep (every read x every haplotype)
    and k_hist (hap_block_from_partition).  achieved = stored cells x 1.375 B / CUDA-event time per launch."""
    from floria_b200 import default_params

    d = ctx.bench_synth_dense(n_reads, n_snps, ploidy, 3)
    prm = default_params(epsilon=0.04)
    sw, hs, cells = ctx.bench_sweep_hist(d, ploidy, d.src, prm, iters)
    out = {"workload": f"configs[2]: {n_reads} full-span frags x {n_snps} SNPs, ploidy {ploidy}, eps 0.04",
           "stored_cells": cells, "packed_bytes": d.nbytes, "iters": iters, "bytes_per_cell": BYTES_PER_CELL}
    for name, t in (("k_sweep", sw), ("k_hist", hs)):
        ms = float(np.median(t))
        gbs = cells * BYTES_PER_CELL / 1e9 / (ms / 1e3)
        out[name] = {"ms_per_launch": ms, "cells_per_s": cells / (ms / 1e3), "achieved_GBps": gbs, "peak_GBps": peak,
                     "frac": gbs / peak}
    d.free()
    return out


def cpu_sample(c, prm, lo, hi, threads):
    """bounded sample of the same workload for the CPU arm: the first `threads` blocks (one per worker)."""
    n = int(min(len(lo), max(1, threads)))
    return lo[:n], hi[:n], f"first {n} of {len(lo)} SNP blocks of the same contig, {threads} threads over blocks"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from floria_b200 import synth  # noqa: F401

    c, prm, lo, hi, desc = make_workload(0)
    threads = os.cpu_count() or 1
    slo, shi, sample = cpu_sample(c, prm, lo, hi, threads)
    times, cells = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = oracle.phase_blocks(c.frags, slo, shi, prm, n_threads=threads)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            cells = r.cells
    ms = 1000.0 * sum(times) / max(len(times), 1)
    value = cells / (ms / 1000.0) if ms > 0 else 0.0
    line = {"impl": "reference", "metric": "read x SNP cells scored per second", "value": value, "unit": "cells/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+u64",
            "data": "synthetic", "config": desc,
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def pinned_frags(frags):
    """copy of the CSR arrays in pinned host memory (cudaHostAlloc via torch)"""
    import torch

    from floria_b200.frags import Frags

    keep = []

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()).pin_memory()
        keep.append(t)
        return t.numpy().view(a.dtype)

    f = Frags.__new__(Frags)
    f.row_ptr, f.pos, f.allele, f.qual = pin(frags.row_ptr), pin(frags.pos), pin(frags.allele), pin(frags.qual)
    f.first, f.last = pin(frags.first), pin(frags.last)
    f.n_reads, f.nnz = frags.n_reads, frags.nnz
    f._pinned = keep
    return f


def run_ours(args):
    import torch
    import torch.distributed as dist

    from floria_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.set_num_threads(1)  # the hot path is on the GPU; N ranks share the host cores with their planning threads
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = api.Context(local_rank)
    c, prm, lo, hi, desc = make_workload(rank)
    dfr = ctx.upload(c.frags)
    hfr = pinned_frags(c.frags)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def gather(res):
        """final gather of the partition records on rank 0 (NCCL over NVLink); units = this rank's blocks"""
        if world == 1:
            return None
        from floria_b200 import shard

        unit_ids = np.arange(res.n_blocks, dtype=np.int64) + rank * 1000000
        return shard.gather_records(unit_ids, res.read_ptr, res.read_ids, res.hap, res.best_ploidy, dev, dst=0, lazy=True)

    def one_pass(fn, steps, sampler=None):
        ev = []
        cells = 0
        for _ in range(steps):
            flush.zero_()  # L2 flush between iterations (inputs are smaller than L2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            res = fn()
            t1 = time.perf_counter()
            gather(res)
            t2 = time.perf_counter()
            if args.verbose and rank == 0:
                sys.stderr.write(f"[step] compute {1e3 * (t1 - t0):.2f} ms, gather {1e3 * (t2 - t1):.2f} ms\n")
            e1.record()
            torch.cuda.synchronize()
            if sampler:
                sampler.sample()
            ev.append(e0.elapsed_time(e1))
            cells = res.cells
        return ev, cells, res

    resident = lambda: ctx.phase_blocks_resident(dfr, lo, hi, prm)
    e2e_fn = lambda: ctx.phase_blocks(hfr, lo, hi, prm)

    one_pass(resident, args.warmup)
    one_pass(e2e_fn, min(args.warmup, 3))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t_before = ctx.timings()
    ev, cells, res = one_pass(resident, args.steps, sampler)
    t_after = ctx.timings()
    barrier()
    clocks = sampler.result() if sampler else None
    my_ms = sum(ev)
    barrier()
    ev2, cells2, res2 = one_pass(e2e_fn, args.steps, sampler)
    barrier()
    my_ms2 = sum(ev2)

    tot = torch.tensor([my_ms, my_ms2], device=dev, dtype=torch.float64)
    cl = torch.tensor([float(cells)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        dist.all_reduce(cl, op=dist.ReduceOp.SUM)
    ms_total, ms_total2 = float(tot[0].item()), float(tot[1].item())
    cells_all = float(cl[0].item())

    if rank == 0:
        K = args.steps
        value = cells_all * K / (ms_total / 1000.0)
        e2e_value = cells_all * K / (ms_total2 / 1000.0)
        d = {k: t_after[k] - t_before[k] for k in t_after}
        peak, peak_src = hbm_peak()
        # the bandwidth-bound kernels' algorithmic traffic; the beam kernel dominates the step
        kern = {
            "k_beam": {"ms": d["beam_ms"], "launches": d["n_beam_launches"], "cells": res.cells_beam * K},
            "k_sweep": {"ms": d["sweep_ms"], "launches": d["n_sweep_launches"], "cells": d["sweep_cells"]},
            "k_hist": {"ms": d["hist_ms"], "launches": d["n_hist_launches"], "cells": d["hist_cells"]},
        }
        for k in kern.values():
            k["GB/s"] = (k["cells"] * BYTES_PER_CELL / 1e9) / (k["ms"] / 1e3) if k["ms"] > 0 else 0.0
            k["frac_of_peak"] = k["GB/s"] / peak
            k["share_of_step"] = k["ms"] / max(d["total_ms"], 1e-9)
        dom = max(kern, key=lambda n: kern[n]["ms"])
        # DRAM traffic per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tj.get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["GB/s"], "peak": peak, "unit": "GB/s",
                    "frac": kern[dom]["frac_of_peak"], "traffic": traffic, "peak_source": peak_src,
                    "avg_launch_ms": kern[dom]["ms"] / max(kern[dom]["launches"], 1),
                    "algorithmic_bytes_per_cell": BYTES_PER_CELL, "kernels": kern,
                    "note": "k_beam is dependency-bound (one sequential step per read); the HBM-bound kernels are "
                            "k_sweep/k_hist, see kernels{} and profiles/"}
        h2d = int(sum(a.nbytes for a in (c.frags.row_ptr, c.frags.pos, c.frags.allele, c.frags.qual, c.frags.first,
                                         c.frags.last)) + lo.nbytes + hi.nbytes)
        d2h = int(res2.read_ids.nbytes + res2.hap.nbytes + res2.mec_vector.nbytes + res2.best_ploidy.nbytes)
        line = {"metric": "read x SNP cells scored per second", "value": value, "unit": "cells/s", "n_gpus": world,
                "steps": K, "warmup": args.warmup, "ms_per_step": ms_total / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64 (2^-26 fixed point) + f64", "data": "synthetic",
                "config": dict(desc, l2="flushed between timed iterations (256 MiB write)",
                               parallelism=f"contig-sharded x{world}, NCCL gather of partition records"),
                "clocks": clocks, "gpu_launches": int(d["n_launches"]),
                "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_total2 / K},
                "roofline": roofline, "cells_per_step": cells_all,
                "cells_breakdown": {"sweep": res.cells_sweep, "hist": res.cells_hist, "beam": res.cells_beam}}
        if world == 1 and not args.no_c3:
            line["roofline_c3"] = c3_roofline(ctx, peak)
        if world == 1 and not args.no_cpu:
            import oracle

            threads = os.cpu_count() or 1
            slo, shi, sample = cpu_sample(c, prm, lo, hi, threads)
            t0 = time.perf_counter()
            r = oracle.phase_blocks(c.frags, slo, shi, prm, n_threads=threads)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": r.cells / dt, "unit": "cells/s", "cores": threads, "kind": "port",
                                    "sample": sample, "seconds": dt}
            # the sample must agree with the GPU result on the same blocks
            n = len(slo)
            same = bool(np.array_equal(r.hap, res.hap[: int(res.read_ptr[n])]) and
                        np.array_equal(r.best_ploidy, res.best_ploidy[:n]))
            line["cpu_baseline"]["matches_gpu"] = same
        print(json.dumps(line), flush=True)
    dfr.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the configs[2] sweep/hist roofline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
