"""per-rank time of the configs[4] LPT shares when each share runs alone on one GPU: python tools/share_balance.py [world=8]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import bench
from floria_b200 import api
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ts = []
for rank in range(world):
    mine, contigs, blocks, prm, stats = bench.c5_workload(world, rank)
    m = api.MultiContext([0])
    d = m.upload(contigs, blocks)
    for _ in range(2):
        m.phase_contigs_resident(d, prm)
    t = time.perf_counter(); res, dev, ms = m.phase_contigs_resident(d, prm); dt = time.perf_counter() - t
    cells = sum(r.cells for r in res)
    pl = np.concatenate([r.best_ploidy for r in res])
    print(f"rank {rank}: {len(mine)} contigs, nnz {sum(c.nnz for c in contigs)}, {1e3*dt:.1f} ms wall, device {ms[0]:.1f} ms, cells {cells:.3e}, mean best ploidy {pl.mean():.2f}", flush=True)
    ts.append(dt)
    d.free(); m.close()
print(f"max {1e3*max(ts):.1f} ms, mean {1e3*np.mean(ts):.1f} ms, imbalance {max(ts)/np.mean(ts):.3f}")
