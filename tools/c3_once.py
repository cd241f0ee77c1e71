"""configs[2]-shaped sweep/hist loop (for ncu captures): python tools/c3_once.py [n_reads] [n_snps] [ploidy] [iters]"""
import sys; sys.path.insert(0, '.')
import numpy as np
from floria_b200 import api, default_params
R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
P = int(sys.argv[3]) if len(sys.argv) > 3 else 4
it = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = api.Context(0)
d = ctx.bench_synth_dense(R, S, P, 3)
sw, hs, cells = ctx.bench_sweep_hist(d, P, d.src, default_params(epsilon=0.04), it)
bpc = 1.375
print(f"cells {cells}  sweep ms {np.median(sw):.3f} ({cells*bpc/1e9/(np.median(sw)/1e3):.0f} GB/s)  hist ms {np.median(hs):.3f} ({cells*bpc/1e9/(np.median(hs)/1e3):.0f} GB/s)")
