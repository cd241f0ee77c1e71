"""one resident pass of the bench workload (for ncu captures)"""
import sys; sys.path.insert(0, '.')
import bench
from floria_b200 import api
ctx = api.Context(0)
c, prm, lo, hi, desc = bench.make_workload(0)
d = ctx.upload(c.frags)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for i in range(n):
    r = ctx.phase_blocks_resident(d, lo, hi, prm)
print(r.cells, ctx.timings()['beam_ms'])
