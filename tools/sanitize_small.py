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

# short paired reads (2-lane sweep teams) and a small batched metagenome (mixed ploidy)
from floria_b200 import shard
s2 = synth.make_contig(78, 300, 120, 2, paired_short=True, flip=0.01, qual_mode="short")
lo2, hi2 = api.get_range_with_lengths(s2.snp_to_genome_pos, 3000, 1000, 0.0005)
print('short ok', ctx.phase_blocks(s2.frags, lo2, hi2, default_params(epsilon=0.01, max_ploidy=3)).best_ploidy[:4])
contigs, blocks = [], []
for k in range(3):
    cc = synth.config5_contig(k, n_reads=80, n_snps=70, span_mean=30)
    contigs.append(cc.frags); blocks.append(api.get_range_with_lengths(cc.snp_to_genome_pos, 4000, 1333, 0.0005))
fr, lo3, hi3, _, _, _ = shard.concat_contigs(contigs, blocks)
print('batched ok', ctx.phase_blocks(fr, lo3, hi3, default_params(epsilon=0.04, max_ploidy=5)).best_ploidy[:6])
