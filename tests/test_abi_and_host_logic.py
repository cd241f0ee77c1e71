"""CPU-only checks: the C-ABI library loads and exports every declared symbol, fails loudly without a GPU, and the
host-compilable logic shared with the kernels (fb_seq.h) matches independent implementations."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from floria_b200 import api, build, synth
from floria_b200._cdefs import f64p, ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "floria_b200.h")).read()
    declared = set(re.findall(r"\b(fb_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(api.EXPORTS)
    L = api.load_library()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} is declared in include/floria_b200.h but not exported"
    # every other header under include/ (measurement helpers) must resolve too
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if h.endswith(".h") and h != "floria_b200.h":
            for name in set(re.findall(r"\b(fb_[a-z_0-9]+)\s*\(", open(os.path.join(ROOT, "include", h)).read())):
                assert hasattr(L, name), f"{name} is declared in include/{h} but not exported"


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.FloriaB200Error) as ei:
        api.Context(0)
    assert "no CUDA device" in str(ei.value) or "fb_init failed" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "floria_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "floria_oracle" not in txt, f


def test_host_helpers_match_oracle():
    c = synth.make_contig(21, 500, 700, 2, span_mean=60)
    for bl, dens in ((10000, 0.0005), (3000, 0.0005), (500, 0.02)):
        lo1, hi1 = oracle.get_range_with_lengths(c.snp_to_genome_pos, bl, bl // 3, dens)
        lo2, hi2 = api.get_range_with_lengths(c.snp_to_genome_pos, bl, bl // 3, dens)
        assert np.array_equal(lo1, lo2) and np.array_equal(hi1, hi2)
        for a, b in list(zip(lo1, hi1))[:5]:
            assert np.array_equal(oracle.find_reads_in_interval(int(a), int(b), c.frags),
                                  api.find_reads_in_interval(int(a), int(b), c.frags))


@pytest.fixture(scope="module")
def ht():
    L = C.CDLL(build.build_hosttest())
    L.ht_seqsum.restype = C.c_double
    L.ht_seqsum.argtypes = [C.POINTER(C.c_longlong), C.c_ulonglong, C.c_double, C.c_int]
    L.ht_binom.restype = C.c_double
    L.ht_binom.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_double, C.c_double]
    L.ht_lse.restype = C.c_double
    L.ht_lse.argtypes = [f64p, C.c_int]
    L.ht_mec_threshold.restype = C.c_double
    L.ht_mec_threshold.argtypes = [C.c_uint, C.c_double, C.c_uint]
    return L


def test_seqsum_equals_plain_left_to_right_f64_sum(ht):
    """The run-lumping accumulator used by the kernels must equal `for x in items: s += x` bit for bit."""
    rng = np.random.default_rng(7)
    lut = oracle.phred_lut().astype(np.float64)
    for trial in range(300):
        n = int(rng.integers(1, 400))
        eps = [0.04, 0.01, 0.03125, 0.1, 1.0 / 3.0][trial % 5]
        p_eps = [0.0, 0.02, 0.3, 0.9][trial % 4]
        q = rng.integers(0, 60, n)
        w = (lut[q] * 2.0 ** 26).astype(np.int64)
        if trial % 7 == 0:
            w *= int(rng.integers(1, 2000))  # large counts, as in the MEC reduction
        if trial % 11 == 0:
            w *= 1 << 12  # push past 2^27 where even dyadic sums round
        is_eps = rng.random(n) < p_eps
        items = np.where(is_eps, -1, w).astype(np.int64)
        s = 0.0
        for k in range(n):
            s = s + (eps if items[k] < 0 else float(items[k]) * 2.0 ** -26)
        for chunk in (1, 16, 32):
            got = ht.ht_seqsum(items.ctypes.data_as(C.POINTER(C.c_longlong)), n, eps, chunk)
            assert got == s, (trial, chunk, got.hex(), s.hex())


def test_heap_emulation_matches_oracle_heap(ht):
    rng = np.random.default_rng(11)
    OL = oracle.lib()
    for trial in range(200):
        n = int(rng.integers(1, 120))
        width = int(rng.integers(1, 40))
        scores = np.round(rng.random(n) * 8) / 4.0 if trial % 2 else rng.random(n)  # many ties / no ties
        scores = np.ascontiguousarray(scores, np.float64)
        a_d, a_s = np.zeros(n, np.int32), np.zeros(n, np.int32)
        b_d, b_s = np.zeros(n, np.int32), np.zeros(n, np.int32)
        la = ht.ht_heap(ptr(scores, f64p), n, width, a_d.ctypes.data_as(C.POINTER(C.c_int)),
                        a_s.ctypes.data_as(C.POINTER(C.c_int)))
        lb = OL.orc_heap_trace(ptr(scores, f64p), n, width, b_d.ctypes.data_as(C.POINTER(C.c_int)),
                               b_s.ctypes.data_as(C.POINTER(C.c_int)))
        assert la == lb == min(n, width)
        assert np.array_equal(a_d[:la], b_d[:lb]) and np.array_equal(a_s[:la], b_s[:lb])
        assert np.all(np.diff(scores[a_s[:la]]) >= 0)


def test_scalar_formulas_match_oracle(ht):
    for n, k, e in [(0, 0, 0.04), (10, 0, 0.04), (10, 2, 0.04), (10, 10, 0.04), (100, 4, 0.04), (57, 9, 0.03125)]:
        assert ht.ht_binom(n, k, e, 0.25) == oracle.stable_binom_cdf_p_rev(n, k, e, 0.25)
    ps = np.array([oracle.stable_binom_cdf_p_rev(10, k, 0.04, 0.25) for k in (0, 2, 5)])
    assert ht.ht_lse(ptr(ps, f64p), 3) == oracle.log_sum_exp(ps)
    for s in (1, 2, 3):
        for p in (2, 3, 4, 5, 6):
            assert ht.ht_mec_threshold(p, 0.04, s) == oracle.mec_threshold(p, 0.04, s)


def test_add_eps_n_equals_the_plain_loop(ht):
    """fb_add_eps_n (m consecutive `S += eps` in closed form per binade) must be bit-identical to the loop."""
    ht.ht_add_eps_n.restype = C.c_double
    ht.ht_add_eps_n.argtypes = [C.c_double, C.c_double, C.c_ulonglong]
    rng = np.random.default_rng(5)
    cases = []
    for eps in [0.04, 0.01, 0.03125, 0.1, 1e-3, 0.3, 2.0 ** -30 * 3, 0.04 * (1 + 2.0 ** -52), 0.75]:
        for S in [0.0, 2.0 ** -26, 0.5, 1.0 - 2.0 ** -53, 1.0, 3.96875, 17.25, 1023.999, 2.0 ** 20 + 0.5, 7.0 * 2.0 ** -26]:
            for m in [0, 1, 2, 3, 17, 100, 1000, 4097]:
                cases.append((S, eps, m))
    for _ in range(400):
        S = float(rng.integers(0, 2 ** 40)) * 2.0 ** -26 if rng.random() < 0.5 else float(rng.random() * 10.0 ** float(rng.integers(-3, 6)))
        cases.append((S, float(rng.random() * 0.2 + 1e-4), int(rng.integers(0, 3000))))
    for S, eps, m in cases:
        want = np.float64(S)
        e = np.float64(eps)
        for _ in range(m):
            want = want + e
        got = ht.ht_add_eps_n(S, eps, m)
        assert np.float64(got).view(np.uint64) == np.float64(want).view(np.uint64), (S, eps, m, got, float(want))
