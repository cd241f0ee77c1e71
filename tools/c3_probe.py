"""configs[2] probe: beam -> optimize -> no-phred MEC on a resident full-span block.
python tools/c3_probe.py [n_reads] [n_snps] [ploidy] [eps]"""
import sys, time; sys.path.insert(0, '.')
import numpy as np
from floria_b200 import api, default_params
R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
P = int(sys.argv[3]) if len(sys.argv) > 3 else 4
eps = float(sys.argv[4]) if len(sys.argv) > 4 else 0.04
ctx = api.Context(0)
t = time.time()
d = ctx.bench_synth_dense(R, S, P, 3)
print(f"synth {time.time()-t:.2f} s, {d.nbytes/1e9:.2f} GB packed", flush=True)
prm = default_params(epsilon=eps, max_ploidy=P)
for rep in range(2):
    t = time.time()
    hap, bases, errors, info = ctx.phase_block(d, None, P, prm)
    wall = time.time() - t
    tm = ctx.timings()
    cells = info["cells_sweep"] + info["cells_hist"] + info["cells_beam"]
    conf = np.zeros((P, P), np.int64); np.add.at(conf, (d.src, hap), 1)
    print(f"rep {rep}: wall {wall*1e3:.1f} ms | {info} | cells {cells:.3e} -> {cells/wall:.3e} cells/s | "
          f"truth recovery {conf.max(axis=1).sum()/R:.4f} | mec {errors.sum():.1f}", flush=True)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()}, flush=True)
