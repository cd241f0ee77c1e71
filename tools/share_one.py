"""one LPT share of configs[4] alone on one GPU: python tools/share_one.py [world=8] [rank=3] [reps=4]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import bench
from floria_b200 import api
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
mine, contigs, blocks, prm, stats = bench.c5_workload(world, rank)
m = api.MultiContext([0])
d = m.upload(contigs, blocks)
ts = []
for i in range(reps + 2):
    t = time.perf_counter(); res, dev, ms = m.phase_contigs_resident(d, prm); dt = time.perf_counter() - t
    if i >= 2: ts.append(dt)
pl = np.concatenate([r.best_ploidy for r in res])
print(f"world {world} rank {rank}: {len(mine)} contigs, {sum(r.n_blocks for r in res)} blocks, wall ms {[round(1e3*x,1) for x in ts]}, device {ms[0]:.1f}, best ploidy sum {int(pl.sum())}, cells {sum(r.cells for r in res):.4e}", flush=True)
