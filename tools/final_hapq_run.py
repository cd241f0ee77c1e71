"""Timing of rows a14 / a15 (process_reads_for_final_parts, get_hapq) at the configs[3] shape (paired short reads, ploidy 3):
   python tools/final_hapq_run.py [scale=0.25]
floria feeds these calls with the haplosets of its path extraction (out of scope here); this script uses every block's
best partition as haplosets (ranges = the block ranges), which gives each read 1-2 candidate haplosets like the real input."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from floria_b200 import api, default_params, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
ctx = api.Context(0)
c = synth.config4(scale)
fr = c.frags
lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 500, 500 // 3, 0.0005)
prm = default_params(epsilon=0.01, max_ploidy=3, block_length=500)
t = time.perf_counter()
r = ctx.phase_blocks(fr, lo, hi, prm)
print(f"configs[3] x {scale}: {fr.n_reads} reads, {fr.nnz} cells, {len(lo)} blocks; phase_blocks {1e3 * (time.perf_counter() - t):.1f} ms, "
      f"{r.cells / (time.perf_counter() - t):.3e} cells/s", flush=True)
pp, pr, rl, rh = [0], [], [], []
for j in range(r.n_blocks):
    a, b = int(r.read_ptr[j]), int(r.read_ptr[j + 1])
    ids, hap = r.read_ids[a:b], r.hap[a:b]
    for h in range(int(r.best_ploidy[j])):
        sel = ids[hap == h]
        pr.append(sel)
        pp.append(pp[-1] + len(sel))
        rl.append(lo[j])
        rh.append(hi[j])
pr = np.concatenate(pr) if pr else np.zeros(0, np.uint32)
print(f"{len(pp) - 1} haplosets, {len(pr)} (read, haploset) memberships", flush=True)
for rep in range(2):
    t = time.perf_counter()
    parts = ctx.process_reads_for_final_parts(fr, pp, pr, rl, rh, prm)
    t1 = time.perf_counter()
    hapq, rel, avg = ctx.get_hapq(fr, parts.part_ptr, parts.read_ids, parts.range_lo, parts.range_hi, c.snp_to_genome_pos, prm)
    t2 = time.perf_counter()
    print(f"rep {rep}: process_reads_for_final_parts {1e3 * (t1 - t):.1f} ms ({parts.n_parts} parts out) | get_hapq {1e3 * (t2 - t1):.1f} ms "
          f"(avg_err {avg:.4f}, median HAPQ {int(np.median(hapq))})", flush=True)
