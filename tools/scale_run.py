"""Throughput of the batched hot path on the many-instance configs (not the bench line; evidence for profiles/):
   python tools/scale_run.py c5 [n_contigs]     BASELINE.json configs[4]: metagenome, mixed ploidy 2..6, one batched call
   python tools/scale_run.py c4 [scale]         BASELINE.json configs[3]: paired short reads, ploidy 3
Prints cells/s of fb_phase_blocks_resident and the per-kernel CUDA-event times; checks run-to-run determinism and the
partition invariants (every read of a block in exactly one haplotype < best ploidy)."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from floria_b200 import api, default_params, shard, synth  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "c5"
ctx = api.Context(0)
t0 = time.perf_counter()
if kind == "c5":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    contigs, blocks = [], []
    for k in range(n):
        c = synth.config5_contig(k)
        contigs.append(c.frags)
        blocks.append(api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005))
    fr, lo, hi, owner, _, _ = shard.concat_contigs(contigs, blocks)
    prm = default_params(epsilon=0.04, max_ploidy=6)
    name = f"configs[4]: {n} contigs batched"
else:
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
    c = synth.config4(scale)
    fr = c.frags
    lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 500, 500 // 3, 0.0005)
    prm = default_params(epsilon=0.01, max_ploidy=3, block_length=500)
    name = f"configs[3] x {scale}"
print(f"{name}: {fr.n_reads} reads, {fr.nnz} stored cells, {len(lo)} blocks (generated in {time.perf_counter() - t0:.1f} s)", flush=True)
d = ctx.upload(fr)
res = None
for it in range(3):
    b = ctx.timings()
    t = time.perf_counter()
    r = ctx.phase_blocks_resident(d, lo, hi, prm)
    w = time.perf_counter() - t
    a = ctx.timings()
    print(f"pass {it}: {1e3 * w:.1f} ms wall, {r.cells / w:.3e} cells/s; " +
          ", ".join(f"{k} {a[k] - b[k]:.1f}" for k in ("beam_ms", "sweep_ms", "hist_ms", "mec_ms", "select_ms", "total_ms")), flush=True)
    if res is not None:
        assert np.array_equal(res.hap, r.hap) and np.array_equal(res.best_ploidy, r.best_ploidy), "not deterministic"
    res = r
for j in range(res.n_blocks):
    h = res.hap[int(res.read_ptr[j]):int(res.read_ptr[j + 1])]
    assert len(h) == 0 or h.max() < res.best_ploidy[j]
print("ploidy histogram:", np.bincount(res.best_ploidy, minlength=8).tolist(), "cells", res.cells_sweep, res.cells_hist, res.cells_beam)
