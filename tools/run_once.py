"""resident passes of configs[1] (10k x 5k, ploidy 2, fb_phase_blocks over 73 blocks): for ncu captures and FB_BEAM_PROF"""
import sys; sys.path.insert(0, '.')
from floria_b200 import api, default_params, synth
ctx = api.Context(0)
c = synth.config2()
prm = default_params(epsilon=0.04, max_ploidy=2, block_length=10000)
lo, hi = api.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005)
d = ctx.upload(c.frags)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for i in range(n):
    b = ctx.timings()['beam_ms']
    r = ctx.phase_blocks_resident(d, lo, hi, prm)
    print(r.cells, "beam ms", ctx.timings()['beam_ms'] - b)
