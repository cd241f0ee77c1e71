#!/usr/bin/env python
"""bench.py — read x SNP cells scored per second through the B200 hot path (BASELINE.json metric).

N = 1 (the headline): BASELINE.json configs[2], the largest single-GPU configuration: ONE block of 100k full-span
  fragments x 50k SNPs, ploidy 4, beam 10, eps 0.04.  A step = the body of floria's ploidy loop for that block
  (graph_processing.rs:140-162) through the C-ABI call fb_phase_block: beam_search_phasing -> optimize_clustering ->
  get_mec_stats_epsilon_no_phred.  (The block is phased at the fixed ploidy 4 on an explicit read list because
  find_reads_in_interval drops reads spanning more than 10000 SNPs, local_clustering.rs:44.)
    value : cells/s with the block already packed in HBM (fb_phase_block_resident)
    e2e   : cells/s through fb_phase_block with HOST (pinned) CSR buffers: H2D of 29 GB of cells + packing + compute +
            D2H of the assignment, every step
  Secondary objects of the same line: `configs1` (round-1 headline: fb_phase_blocks over the 73 blocks of the 10k x 5k
  contig) and `shard500` (BASELINE.json configs[4] on this many GPUs, see below).
N > 1 (torchrun, one rank per GPU): BASELINE.json configs[4], the 500-contig metagenome (1M frags, 500k SNPs, mixed
  ploidy 2..6, max ploidy 6) as ONE fixed workload sharded over the ranks (strong scaling): contigs are dealt by the
  library's static LPT queue (fb_lpt_assign), every rank phases its share in one batched call (fb_phase_contigs), rank 0
  gathers the partition records over NCCL inside the timed region.  No data-path collective.  The same workload on one
  GPU is the `shard500` object of the N = 1 line, so value(N) / shard500.value(1) is the strong-scaling speed-up.
  `shard_blocks` (every N): BASELINE.json configs[3] (2M short reads x 100k SNPs, ploidy 3, one contig) with its BLOCKS
  dealt to the ranks as contiguous ranges, the reference's other parallel axis; same strong-scaling reading.
--impl reference : the CPU restatement of the reference path (oracle/, all host threads) on a bounded sample of the
  same workload (the reference itself is Rust and cannot be built in this image; nothing of this repo's CUDA library is
  loaded by this arm).
"""
import argparse
import json
import os
import shutil
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_CELL = 1.375  # 2-bit allele + 8-bit qual + 1-bit presence
METRIC = "read x SNP cells scored per second"
C3 = dict(n_reads=100000, n_snps=50000, ploidy=4, beam=10, epsilon=0.04)
C3_DESC = ("configs[2]: synthetic 1 contig, 100k full-span frags x 50k SNPs, ploidy 4, beam 10 (one block; "
           "beam_search_phasing -> optimize_clustering -> no-phred MEC through fb_phase_block, eps 0.04)")
C5 = dict(n_contigs=500, n_reads=2000, n_snps=1000, max_ploidy=6, epsilon=0.04)
C5_DESC = ("configs[4]: synthetic metagenome, 500 contigs / 1M frags / 500k SNPs, mixed ploidy 2-6 (max ploidy 6), "
           "contig-sharded by a static LPT queue, one batched fb_phase_contigs call per GPU")
CPU_SAMPLE_READS = int(os.environ.get("FB_BENCH_CPU_READS", "8"))  # reads per CPU thread of the configs[2] sample (~10-15 s of oracle time per thread and step; the override is for the contract test)


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
    window (a concurrent sampler stalled CUDA calls on some hosts in round 1)."""

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


# ---- workloads --------------------------------------------------------------------------------------------------------------
def c3_config():
    return {"workload": C3_DESC, "n_reads": C3["n_reads"], "n_snps": C3["n_snps"], "ploidy": C3["ploidy"],
            "beam": C3["beam"], "epsilon": C3["epsilon"],
            "l2": "inputs (6.9 GB packed) are larger than the 126 MB L2; no flush needed between iterations"}


def c3_params():
    from floria_b200 import default_params

    return default_params(epsilon=C3["epsilon"], max_ploidy=C3["ploidy"], max_number_solns=C3["beam"])


def c3_cpu_sample(threads):
    """bounded sample of configs[2] for the CPU arm: the first CPU_SAMPLE_READS * threads reads of the block, cut into
    `threads` slices of consecutive reads; every slice is phased as its own block (beam -> optimize -> no-phred MEC at
    ploidy 4) by one thread: the reference's parallel axis is blocks (rayon par_iter, graph_processing.rs:345-362), a
    single block runs on one core."""
    from floria_b200 import synth

    n = CPU_SAMPLE_READS * threads
    c = synth.make_contig(3, n, C3["n_snps"], C3["ploidy"], full_span=True)  # == reads 0..n-1 of the 100k block
    slices = [np.arange(t * CPU_SAMPLE_READS, (t + 1) * CPU_SAMPLE_READS, dtype=np.uint32) for t in range(threads)]
    desc = (f"{threads} disjoint slices of {CPU_SAMPLE_READS} consecutive full-span reads x {C3['n_snps']} SNPs of the "
            f"same block, each phased as its own block at ploidy {C3['ploidy']} by one thread")
    return c.frags, slices, desc


def run_cpu_sample(frags, slices, prm, ploidy):
    """oracle.phase_block on every slice, one Python thread each (ctypes releases the GIL).  Returns (cells, seconds,
    results)."""
    import oracle

    out = [None] * len(slices)

    def work(t):
        out[t] = oracle.phase_block(frags, slices[t], ploidy, prm)

    th = [threading.Thread(target=work, args=(t,)) for t in range(len(slices))]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    cells = sum(o[3]["cells_sweep"] + o[3]["cells_hist"] + o[3]["cells_beam"] for o in out)
    return cells, dt, out


def c5_config(world):
    return {"workload": C5_DESC, "n_contigs": C5["n_contigs"], "reads_per_contig": C5["n_reads"],
            "snps_per_contig": C5["n_snps"], "max_ploidy": C5["max_ploidy"], "epsilon": C5["epsilon"],
            "l2": "flushed between timed iterations (256 MiB write)",
            "parallelism": f"contig-sharded x{world} (static LPT queue), one process per GPU, NCCL gather of the "
                           "partition records on rank 0"}


def c5_workload(world, rank):
    """configs[4]: this rank's share of the 500 contigs (generated locally: the generator is counter-based, so every rank
    computes the same schedule and its own contigs without communication)."""
    from floria_b200 import api, default_params, synth

    costs, meta = [], []
    for k in range(C5["n_contigs"]):
        c = synth.config5_contig(k, n_reads=C5["n_reads"], n_snps=C5["n_snps"])
        lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005)
        costs.append(api.contig_cost(c.frags, len(lo)))
        meta.append((c.frags, (lo, hi)))
    owner = api.lpt_assign(costs, world)
    mine = [k for k in range(C5["n_contigs"]) if owner[k] == rank]
    contigs = [meta[k][0] for k in mine]
    blocks = [meta[k][1] for k in mine]
    prm = default_params(epsilon=C5["epsilon"], max_ploidy=C5["max_ploidy"])
    stats = {"n_reads": int(sum(m[0].n_reads for m in meta)), "stored_cells": int(sum(m[0].nnz for m in meta)),
             "n_blocks": int(sum(len(m[1][0]) for m in meta))}
    return mine, contigs, blocks, prm, stats


# ---- the reference arm (CPU) ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = int(os.environ.get("FB_BENCH_CPU_THREADS", "0")) or os.cpu_count() or 1
    if args.gpus <= 1:
        prm = c3_params()
        frags, slices, sample = c3_cpu_sample(threads)
        config = c3_config()
        step = lambda: run_cpu_sample(frags, slices, prm, C3["ploidy"])[:2]
        scaling = "weak"
    else:
        # configs[4] sample: every k-th contig so that each thread gets >= 2 blocks, oracle threads over blocks
        import oracle
        from floria_b200 import default_params, synth

        prm = default_params(epsilon=C5["epsilon"], max_ploidy=C5["max_ploidy"])
        n_c = max(2, threads // 4)
        ks = list(range(0, C5["n_contigs"], C5["n_contigs"] // n_c))[:n_c]
        cs = [synth.config5_contig(k, n_reads=C5["n_reads"], n_snps=C5["n_snps"]) for k in ks]
        blocks = [oracle.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005) for c in cs]
        sample = (f"{len(ks)} of the {C5['n_contigs']} contigs (every {C5['n_contigs'] // n_c}-th), all their blocks, "
                  f"{threads} threads over blocks")
        config = c5_config(args.gpus)

        def step():
            t0 = time.perf_counter()
            cells = 0
            for c, (lo, hi) in zip(cs, blocks):
                cells += oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=threads).cells
            return cells, time.perf_counter() - t0

        scaling = "strong"
    times, cells = [], 0
    for i in range(args.warmup + args.steps):
        cells, dt = step()
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * sum(times) / max(len(times), 1)
    value = cells / (ms / 1000.0) if ms > 0 else 0.0
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64 (reference arithmetic)", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample,
                             # BASELINE.md "Baseline B": the real floria needs cargo + 236 crates (no network here); probed at run time
                             "rust_toolchain": {"cargo": shutil.which("cargo") is not None, "rustc": shutil.which("rustc") is not None},
                             "why_port": "the reference is Rust and cannot be built on this box (tools/pin_with_floria.sh builds and "
                                         "diffs it wherever a toolchain and the crates exist); the arm times the C++ restatement"},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- our arm --------------------------------------------------------------------------------------------------------------------
def pinned_alloc_factory():
    """alloc(nbytes) -> uint8 numpy array backed by pinned host memory (cudaHostAlloc via torch)"""
    import torch

    keep = []

    def alloc(n):
        t = torch.empty(max(int(n), 1), dtype=torch.uint8, pin_memory=True)
        keep.append(t)
        return t.numpy()

    alloc.keep = keep
    return alloc


def pinned_frags(frags):
    from floria_b200.frags import Frags

    alloc = pinned_alloc_factory()

    def pin(a):
        b = alloc(a.nbytes)[: a.nbytes].view(a.dtype)
        b[...] = a
        return b

    f = Frags.__new__(Frags)
    f.row_ptr, f.pos, f.allele, f.qual = pin(frags.row_ptr), pin(frags.pos), pin(frags.allele), pin(frags.qual)
    f.first, f.last = pin(frags.first), pin(frags.last)
    f.n_reads, f.nnz = frags.n_reads, frags.nnz
    f._pinned = alloc.keep
    return f


def timed_steps(fn, steps, torch, flush=None, sampler=None, after=None):
    """`steps` calls of fn(), each bracketed by CUDA events (the library's stream is a blocking stream, so events on the
    legacy default stream enclose its work) with a synchronize on both sides; returns (ms list, last result)."""
    ev, res = [], None
    for _ in range(steps):
        if flush is not None:
            flush.zero_()  # L2 flush between iterations (for inputs smaller than L2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = fn()
        if after is not None:
            after(res)
        e1.record()
        torch.cuda.synchronize()
        if sampler:
            sampler.sample()
        ev.append(e0.elapsed_time(e1))
    return ev, res


def kernel_table(d, cells_beam, peak):
    kern = {
        "k_beam_wide": {"ms": d["beam_ms"], "launches": d["n_beam_launches"], "cells": cells_beam},
        "k_sweep": {"ms": d["sweep_ms"], "launches": d["n_sweep_launches"], "cells": d["sweep_cells"]},
        "k_hist": {"ms": d["hist_ms"], "launches": d["n_hist_launches"], "cells": d["hist_cells"]},
    }
    for k in kern.values():
        k["GB/s"] = (k["cells"] * BYTES_PER_CELL / 1e9) / (k["ms"] / 1e3) if k["ms"] > 0 else 0.0
        k["frac_of_peak"] = k["GB/s"] / peak
        k["share_of_step"] = k["ms"] / max(d["total_ms"], 1e-9)
        k["avg_launch_ms"] = k["ms"] / max(k["launches"], 1)
    return kern


def traffic_of(kernel):
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return tj.get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def run_c3(args, ctx, torch, sampler):
    """N = 1 headline: configs[2] through fb_phase_block(_resident)."""
    prm = c3_params()
    P = C3["ploidy"]
    d = ctx.bench_synth_dense(C3["n_reads"], C3["n_snps"], P, 3)
    resident = lambda: ctx.phase_block(d, None, P, prm)
    timed_steps(resident, args.warmup, torch)
    t_before = ctx.timings()
    ev, res = timed_steps(resident, args.steps, torch, sampler=sampler)
    t_after = ctx.timings()
    hap, bases, errors, info = res
    cells = info["cells_sweep"] + info["cells_hist"] + info["cells_beam"]
    K = args.steps
    ms = sum(ev) / K
    dt = {k: t_after[k] - t_before[k] for k in t_after}
    peak, peak_src = hbm_peak()
    kern = kernel_table(dt, info["cells_beam"] * K, peak)
    dom = max(kern, key=lambda n: kern[n]["ms"])
    conf = np.zeros((P, P), np.int64)
    np.add.at(conf, (d.src, hap), 1)
    out = {
        "value": cells / (ms / 1e3), "ms_per_step": ms, "cells_per_step": cells,
        "cells_breakdown": {"sweep": info["cells_sweep"], "hist": info["cells_hist"], "beam": info["cells_beam"],
                            "note": "beam = reference-equivalent work: sum over reads of (#search nodes x cells of the "
                                    "read), global_clustering.rs:74-80; the device scores every distinct haplotype state "
                                    "once per read, so most of these cells are never touched by a load"},
        "gpu_launches": int(dt["n_launches"]),
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["GB/s"], "peak": peak, "unit": "GB/s",
                     "frac": kern[dom]["frac_of_peak"], "traffic": traffic_of(dom), "peak_source": peak_src,
                     "avg_launch_ms": kern[dom]["avg_launch_ms"], "algorithmic_bytes_per_cell": BYTES_PER_CELL,
                     "kernels": kern,
                     "note": "the beam search is one dependent step per read (100k sequential steps): it is bound by the "
                             "latency of a step (grid barrier + decision section), not by HBM; the HBM-bound kernels of the "
                             "step are k_sweep / k_hist, see kernels{}"},
        "result": {"beam_score": info["beam_score"], "opt_score": info["opt_score"], "opt_rounds": info["n_rounds"],
                   "mec": float(errors.sum()), "truth_recovery": float(conf.max(axis=1).sum() / len(hap))},
        "packed_bytes": d.nbytes,
    }
    # ---- end to end: HOST (pinned) CSR buffers through fb_phase_block --------------------------------------------------------
    if not args.no_e2e:
        t0 = time.perf_counter()
        hfr = ctx.bench_export_csr(d, alloc=pinned_alloc_factory())
        d.free()
        d = None
        export_s = time.perf_counter() - t0
        e2e_fn = lambda: ctx.phase_block(hfr, None, P, prm)
        timed_steps(e2e_fn, min(args.warmup, 1), torch)
        ev2, res2 = timed_steps(e2e_fn, args.steps, torch, sampler=sampler)
        ms2 = sum(ev2) / K
        h2d = int(sum(a.nbytes for a in (hfr.row_ptr, hfr.pos, hfr.allele, hfr.qual, hfr.first, hfr.last)))
        d2h = int(res2[0].nbytes + bases.nbytes + errors.nbytes)
        out["e2e"] = {"value": cells / (ms2 / 1e3), "unit": "cells/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "ms_per_step": ms2, "host_buffers": "pinned CSR (fb_frags), built once "
                      f"from the device-generated block in {export_s:.1f} s (untimed)",
                      "same_result": bool(np.array_equal(res2[0], hap))}
        del hfr
    if d is not None:
        d.free()
    return out


def run_c3_cpu_baseline(ctx):
    """the CPU restatement on the bounded sample, checked against the GPU on the same slices"""
    threads = os.cpu_count() or 1
    prm = c3_params()
    frags, slices, sample = c3_cpu_sample(threads)
    cells, dt, outs = run_cpu_sample(frags, slices, prm, C3["ploidy"])
    same = True
    for t in (0, len(slices) // 2, len(slices) - 1):  # GPU on three of the slices
        gh, gb, ge, gi = ctx.phase_block(frags, slices[t], C3["ploidy"], prm)
        oh, ob, oe, oi = outs[t]
        same &= bool(np.array_equal(gh, oh) and np.array_equal(ge.view(np.uint64), oe.view(np.uint64)) and
                     (gi["cells_sweep"], gi["cells_hist"], gi["cells_beam"]) ==
                     (oi["cells_sweep"], oi["cells_hist"], oi["cells_beam"]))
    return {"value": cells / dt, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample, "seconds": dt,
            "matches_gpu": same}


def run_configs1(args, ctx, torch):
    """round-1 headline kept as a secondary object: fb_phase_blocks over the 73 blocks of the 10k x 5k contig"""
    from floria_b200 import api, default_params, synth

    c = synth.make_contig(2, 10000, 5000, 2, span_mean=500, flip=0.04, qual_mode="long")
    prm = default_params(epsilon=0.04, max_ploidy=2, block_length=10000)
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005)
    dfr = ctx.upload(c.frags)
    hfr = pinned_frags(c.frags)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    steps = max(3, min(args.steps, 10))
    timed_steps(lambda: ctx.phase_blocks_resident(dfr, lo, hi, prm), 3, torch, flush)
    tb = ctx.timings()
    ev, res = timed_steps(lambda: ctx.phase_blocks_resident(dfr, lo, hi, prm), steps, torch, flush)
    ta = ctx.timings()
    timed_steps(lambda: ctx.phase_blocks(hfr, lo, hi, prm), 2, torch, flush)
    ev2, res2 = timed_steps(lambda: ctx.phase_blocks(hfr, lo, hi, prm), steps, torch, flush)
    dfr.free()
    ms, ms2 = sum(ev) / steps, sum(ev2) / steps
    return {"workload": "configs[1]: synthetic 1 contig, 10k long-read frags x 5k SNPs, ploidy 2 (fb_phase_blocks over "
                        "all 73 SNP blocks, ploidy loop 1..2, beam 10, eps 0.04)", "steps": steps,
            "value": res.cells / (ms / 1e3), "ms_per_step": ms, "e2e_value": res2.cells / (ms2 / 1e3),
            "e2e_ms_per_step": ms2, "beam_ms_per_step": (ta["beam_ms"] - tb["beam_ms"]) / steps,
            "cells_per_step": res.cells, "l2": "flushed between timed iterations (256 MiB write)"}


def run_shard500(args, torch, dist, rank, world, local_rank, sampler):
    """configs[4] on `world` GPUs, one rank per GPU: this rank's LPT share through fb_phase_contigs(_resident) + NCCL
    gather of the partition records on rank 0 (inside the timed region).  Returns on rank 0 the result object."""
    from floria_b200 import api, shard

    dev = torch.device("cuda", local_rank)
    mine, contigs, blocks, prm, stats = c5_workload(world, rank)
    m = api.MultiContext([local_rank])
    dcon = m.upload(contigs, blocks)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pin = [pinned_frags(f) for f in contigs]

    def gather(out):
        """NCCL gather of this rank's partition records on rank 0; unit id = contig * 2^20 + block"""
        if world == 1:
            return
        res = out[0]
        uids = [np.arange(r.n_blocks, dtype=np.int64) + (np.int64(k) << 20) for k, r in zip(mine, res)]
        sizes = [np.diff(r.read_ptr.astype(np.int64)) for r in res]
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
        rp = np.concatenate([[0], np.cumsum(cat(sizes, np.int64))]).astype(np.int64)
        shard.gather_records(cat(uids, np.int64), rp, cat([r.read_ids for r in res], np.uint32),
                             cat([r.hap for r in res], np.uint8), cat([r.best_ploidy for r in res], np.int64), dev,
                             dst=0, lazy=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = lambda: m.phase_contigs_resident(dcon, prm)
    e2e_fn = lambda: m.phase_contigs(pin, blocks, prm)
    timed_steps(resident, args.warmup, torch, flush, after=gather)
    barrier()
    l0 = int(m.timings(0)["n_launches"])
    ev, out = timed_steps(resident, args.steps, torch, flush, sampler=sampler, after=gather)
    launches = int(m.timings(0)["n_launches"]) - l0  # this rank's kernels inside the timed region
    barrier()
    timed_steps(e2e_fn, 1, torch, flush, after=gather)
    barrier()
    ev2, out2 = timed_steps(e2e_fn, args.steps, torch, flush, after=gather)
    barrier()
    my_cells = float(sum(r.cells for r in out[0]))
    tot = torch.tensor([sum(ev), sum(ev2)], device=dev, dtype=torch.float64)
    cl = torch.tensor([my_cells], device=dev, dtype=torch.float64)
    h2d_mine = float(sum(a.nbytes for f in pin for a in (f.row_ptr, f.pos, f.allele, f.qual, f.first, f.last)))
    d2h_mine = float(sum(r.read_ids.nbytes + r.hap.nbytes + r.mec_vector.nbytes + r.best_ploidy.nbytes for r in out2[0]))
    io = torch.tensor([h2d_mine, d2h_mine], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        dist.all_reduce(cl, op=dist.ReduceOp.SUM)
        dist.all_reduce(io, op=dist.ReduceOp.SUM)
    dcon.free()
    m.close()
    if rank != 0:
        return None
    K = args.steps
    ms, ms2 = float(tot[0].item()) / K, float(tot[1].item()) / K
    cells = float(cl[0].item())
    return {"config": c5_config(world), "workload_stats": stats,
            "value": cells / (ms / 1e3), "ms_per_step": ms, "cells_per_step": cells, "n_gpus": world,
            "e2e": {"value": cells / (ms2 / 1e3), "unit": "cells/s", "h2d_bytes_per_step": int(io[0].item()),
                    "d2h_bytes_per_step": int(io[1].item()), "ms_per_step": ms2},
            "contigs_on_rank0": len(mine), "gpu_launches": launches}


C4 = dict(n_reads=2000000, n_snps=100000, ploidy=3, epsilon=0.01, block_length=500)


def run_shard_blocks(args, torch, dist, rank, world, local_rank):
    """configs[3] (2M paired short reads x 100k SNPs, ploidy 3) as ONE contig whose BLOCKS are sharded over the ranks
    (the reference's other parallel axis, graph_processing.rs:345-362): every rank holds the contig resident, phases a
    contiguous range of blocks of equal estimated cost (reads x block incidence) through fb_phase_blocks_resident, and
    the partition records are gathered on rank 0 inside the timed region."""
    from floria_b200 import api, default_params, shard, synth

    dev = torch.device("cuda", local_rank)
    c = synth.config4(1.0)
    prm = default_params(epsilon=C4["epsilon"], max_ploidy=C4["ploidy"], block_length=C4["block_length"])
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, C4["block_length"], C4["block_length"] // 3, 0.0005)
    lo, hi = np.asarray(lo), np.asarray(hi)
    fr = c.frags
    # reads per block from the sorted end points (cost estimate), then `world` contiguous ranges of equal cost
    fs, ls = np.sort(fr.first), np.sort(fr.last)
    per_block = np.searchsorted(fs, hi, side="right") - np.searchsorted(ls, lo, side="left")
    cum = np.concatenate([[0], np.cumsum(per_block.astype(np.float64))])
    cuts = [int(np.searchsorted(cum, cum[-1] * r / world)) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, len(lo)
    a, b = cuts[rank], cuts[rank + 1]
    ctx = api.Context(local_rank)
    d = ctx.upload(fr)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def gather(res):
        if world == 1:
            return
        uids = np.arange(a, b, dtype=np.int64)
        shard.gather_records(uids, res.read_ptr.astype(np.int64), res.read_ids, res.hap, res.best_ploidy, dev, dst=0,
                             lazy=True)

    fn = lambda: ctx.phase_blocks_resident(d, lo[a:b], hi[a:b], prm)
    timed_steps(fn, args.warmup, torch, flush, after=gather)
    if world > 1:
        dist.barrier()
    ev, res = timed_steps(fn, args.steps, torch, flush, after=gather)
    tot = torch.tensor([sum(ev)], device=dev, dtype=torch.float64)
    cl = torch.tensor([float(res.cells)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        dist.all_reduce(cl, op=dist.ReduceOp.SUM)
    ctx.close()
    if rank != 0:
        return None
    ms = float(tot[0].item()) / args.steps
    cells = float(cl[0].item())
    return {"config": {"workload": "configs[3]: synthetic short-read, 2M paired frags x 100k SNPs, ploidy 3, one contig; "
                                   "its %d blocks (-l 500) dealt to the ranks as contiguous ranges of equal estimated cost"
                                   % len(lo), "n_reads": int(fr.n_reads), "n_snps": C4["n_snps"],
                       "max_ploidy": C4["ploidy"], "epsilon": C4["epsilon"], "n_blocks": int(len(lo)),
                       "parallelism": "block ranges x %d" % world},
            "value": cells / (ms / 1e3), "unit": "cells/s", "ms_per_step": ms, "cells_per_step": cells, "n_gpus": world,
            "scaling": "strong", "blocks_on_rank0": int(b - a)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from floria_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.set_num_threads(1)  # the hot path is on the GPU; N ranks share the host cores with their planning threads
    saved_stdout = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created: everything but the JSON line goes
        # to stderr (file descriptor 1 is pointed at 2 until the line is printed)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    base = {"metric": METRIC, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "vs_baseline": None, "dtype": "u64 (2^-26 fixed point) + f64", "data": "synthetic"}
    if world == 1:
        ctx = api.Context(local_rank)
        c3 = run_c3(args, ctx, torch, sampler)
        clocks = sampler.result() if sampler else None
        line = dict(base, value=c3.pop("value"), ms_per_step=c3.pop("ms_per_step"), scaling="weak", config=c3_config(),
                    clocks=clocks, gpu_launches=c3.pop("gpu_launches"), e2e=c3.pop("e2e", None),
                    roofline=c3.pop("roofline"), **c3)
        if not args.no_cpu:
            line["cpu_baseline"] = run_c3_cpu_baseline(ctx)
        if not args.no_secondary:
            line["configs1"] = run_configs1(args, ctx, torch)
        ctx.close()
        if not args.no_secondary:
            s5 = run_shard500(args, torch, dist, rank, world, local_rank, None)
            line["shard500"] = s5
            line["shard_blocks"] = run_shard_blocks(args, torch, dist, rank, world, local_rank)
            line["note"] = ("N=1 headline is configs[2] (one block: does not shard).  The multi-GPU workload is "
                            "configs[4] (`shard500`, strong scaling): compare value of the N>1 lines with "
                            "shard500.value of this line.")
        print(json.dumps(line), flush=True)
    else:
        s5 = run_shard500(args, torch, dist, rank, world, local_rank, sampler)
        sb = None if args.no_secondary else run_shard_blocks(args, torch, dist, rank, world, local_rank)
        if rank == 0:
            clocks = sampler.result() if sampler else None
            line = dict(base, value=s5["value"], ms_per_step=s5["ms_per_step"], scaling="strong", config=s5["config"],
                        clocks=clocks, e2e=s5["e2e"], cells_per_step=s5["cells_per_step"], shard_blocks=sb,
                        gpu_launches=s5["gpu_launches"],
                        note="ONE fixed workload (configs[4], 500 contigs) sharded over the ranks; the 1-GPU value of "
                             "the same workload is shard500.value of the N=1 line")
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            print(json.dumps(line), flush=True)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg of configs[2]")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[1] / configs[4] secondary objects")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
