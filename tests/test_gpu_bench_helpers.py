"""The device-side generator of the configs[2] roofline block must produce exactly the cells of the numpy generator, and
the timed sweep/hist loop must leave results identical to the fine-grained entry points."""
import numpy as np
import pytest

from floria_b200 import api, default_params, synth

pytestmark = pytest.mark.gpu


def test_device_synth_matches_numpy_generator():
    ctx = api.Context(0)
    R, S, P = 300, 1000, 4
    d = ctx.bench_synth_dense(R, S, P, 3)
    c = synth.make_contig(3, R, S, P, full_span=True, flip=0.04, qual_mode="long")
    h = ctx.upload(c.frags)
    q1, a1, p1 = ctx.download_planes(d)
    q2, a2, p2 = ctx.download_planes(h)
    assert np.array_equal(p1, p2) and np.array_equal(a1, a2) and np.array_equal(q1, q2)
    assert np.array_equal(d.src, c.read_hap)
    prm = default_params(epsilon=0.04)
    sw, hs, cells = ctx.bench_sweep_hist(d, P, d.src, prm, 2)
    assert cells == c.frags.nnz and (sw > 0).all() and (hs > 0).all()
    d.free()
    h.free()
    ctx.close()
