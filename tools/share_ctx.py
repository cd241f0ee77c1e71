"""one rank's configs[4] share through 1, 2, 3 or 4 contexts on the SAME device (the library deals the contigs to the
contexts; their kernels interleave on the GPU, so one context's serial phases overlap another's work):
python tools/share_ctx.py [world=8] [rank=0]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import bench
from floria_b200 import api
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mine, contigs, blocks, prm, stats = bench.c5_workload(world, rank)
for nctx in (1, 2, 3, 4):
    m = api.MultiContext([0] * nctx)
    d = m.upload(contigs, blocks)
    ts = []
    for _ in range(4):
        t = time.perf_counter(); res, dev, ms = m.phase_contigs_resident(d, prm); ts.append(time.perf_counter() - t)
    print(f"world {world} rank {rank}: {len(mine)} contigs on {nctx} context(s): {1e3*min(ts[1:]):.1f} ms wall (device ms per context {np.round(ms, 1).tolist()})", flush=True)
    d.free(); m.close()
