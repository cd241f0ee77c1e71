import sys; sys.path.insert(0,'.')
import bench, time
from floria_b200 import api
ctx = api.Context(0)
c, prm, lo, hi, desc = bench.make_workload(0)
d = ctx.upload(c.frags)
for i in range(3):
    t=time.perf_counter(); r = ctx.phase_blocks_resident(d, lo, hi, prm); print('wall ms', (time.perf_counter()-t)*1e3)
print(ctx.timings())
