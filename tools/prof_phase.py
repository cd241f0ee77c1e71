import sys, os; sys.path.insert(0,'.')
import bench, time
from floria_b200 import api
ctx = api.Context(0)
c, prm, lo, hi, desc = bench.make_workload(0)
if len(sys.argv) > 1: prm.epsilon = float(sys.argv[1])
d = ctx.upload(c.frags)
for i in range(3):
    t0 = ctx.timings()
    t=time.perf_counter(); r = ctx.phase_blocks_resident(d, lo, hi, prm); w=(time.perf_counter()-t)*1e3
    t1 = ctx.timings()
    print('wall ms %.2f' % w, {k: round(t1[k]-t0[k],3) for k in ('beam_ms','sweep_ms','hist_ms','mec_ms','select_ms','total_ms','n_launches')})
t=time.perf_counter(); r = ctx.phase_blocks(c.frags, lo, hi, prm); print('e2e wall ms %.2f' % ((time.perf_counter()-t)*1e3))
t=time.perf_counter(); r = ctx.phase_blocks(c.frags, lo, hi, prm); print('e2e wall ms %.2f' % ((time.perf_counter()-t)*1e3))
