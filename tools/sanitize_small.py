"""small end-to-end run for compute-sanitizer (memcheck / racecheck)"""
import sys; sys.path.insert(0, '.')
import numpy as np
from floria_b200 import api, default_params, synth
ctx = api.Context(0)
c = synth.make_contig(77, 160, 200, 3, span_mean=50)
prm = default_params(epsilon=0.04, max_ploidy=4)
lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 3333, 0.0005)
r = ctx.phase_blocks(c.frags, lo, hi, prm)
print('phase ok', r.best_ploidy)
ptr, reads, rlo, rhi = [0], [], [], []
for j in range(r.n_blocks):
    ids = r.read_ids[r.read_ptr[j]:r.read_ptr[j+1]]; hp = r.hap[r.read_ptr[j]:r.read_ptr[j+1]]
    for h in range(int(r.best_ploidy[j])):
        reads.extend(ids[hp == h].tolist()); ptr.append(len(reads)); rlo.append(int(lo[j])); rhi.append(int(hi[j]))
p = ctx.process_reads_for_final_parts(c.frags, ptr, reads, rlo, rhi, prm)
hq = ctx.get_hapq(c.frags, p.part_ptr, p.read_ids, p.range_lo, p.range_hi, c.snp_to_genome_pos, prm)
print('final ok', p.n_parts, hq[0][:5])
d = ctx.bench_synth_dense(2000, 3000, 3, 3)
print(ctx.bench_sweep_hist(d, 3, d.src, prm, 1)[2])
