"""A compiled C client of include/floria_b200.h (gcc, tests/c_abi_smoke.c): the struct layouts the C compiler sees must be
the ones the ctypes mirror (floria_b200/_cdefs.py) uses, and the calls a cgo / Rust binding would make
(init -> upload -> phase -> final parts -> hapq) must give what the Python binding gives on the same contig."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from floria_b200 import _cdefs, api, default_params
from floria_b200.frags import Frags

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    api.load_library()  # the library must exist (no CPU fallback); built by __graft_entry__.build()
    out = str(tmp_path_factory.mktemp("cabi") / "c_abi_smoke")
    lib = os.path.join(ROOT, "floria_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-L" + lib, "-lfloria_b200",
                           "-Wl,-rpath," + lib, "-lm", "-o", out])
    return out


def test_struct_layouts_match_the_ctypes_mirror(exe):
    got = {}
    for line in subprocess.check_output([exe, "layout"], text=True).split("\n"):
        if line:
            k, v = line.split()
            got[k] = int(v)
    mirror = {"fb_params": _cdefs.FbParams, "fb_frags": _cdefs.FbFrags, "fb_block_results": _cdefs.FbBlockResults,
              "fb_block_phase": _cdefs.FbBlockPhase, "fb_parts": _cdefs.FbParts, "fb_timings": _cdefs.FbTimings}
    n_fields = 0
    for name, cls in mirror.items():
        assert got[name] == C.sizeof(cls), f"sizeof({name})"
        for f, _ in cls._fields_:
            key = f"{name}.{f}"
            if key in got:
                assert got[key] == getattr(cls, f).offset, key
                n_fields += 1
    assert n_fields >= 45


def _contig():
    """tests/c_abi_smoke.c run(): the same splitmix64 stream"""
    state = [12345]
    M = (1 << 64) - 1

    def sm64():
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & M
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    R, S, SPAN = 240, 200, 40
    truth0 = [sm64() & 1 for _ in range(S)]
    truth = [truth0, [1 - x for x in truth0]]
    g = np.array([100 * s + 7 for s in range(S)], np.uint64)
    reads = []
    for r in range(R):
        f = 1 + (r * (S - SPAN)) // R
        h = sm64() & 1
        pos, al, q = [], [], []
        for k in range(SPAN):
            a = truth[h][f + k - 1]
            if sm64() % 25 == 0:
                a = 1 - a
            pos.append(f + k)
            al.append(a)
            q.append(10 + sm64() % 30)
        reads.append((pos, al, q))
    return Frags.from_reads(reads, sort=False), g


@pytest.mark.gpu
def test_c_client_matches_python_binding(exe):
    out = subprocess.check_output([exe, "run"], text=True).strip().split("\n")
    fr, g = _contig()
    ctx = api.Context(0)
    prm = default_params(max_ploidy=3)
    lo, hi = api.get_range_with_lengths(g, 5000, 5000 // 3, 0.0005)
    r = ctx.phase_blocks(fr, lo, hi, prm)
    assert out[0] == f"blocks {len(lo)}"
    for j in range(len(lo)):
        m = r.mec_vector[j]
        assert out[1 + j] == (f"block {j} ploidy {r.best_ploidy[j]} mec {m[0]:.17g} {m[1]:.17g} {m[2]:.17g} "
                              f"reads {int(r.read_ptr[j + 1] - r.read_ptr[j])}")
    pp, pr, rl, rh = [0], [], [], []
    for j in range(len(lo)):
        a, b = int(r.read_ptr[j]), int(r.read_ptr[j + 1])
        for h in range(int(r.best_ploidy[j])):
            pr.extend(r.read_ids[a:b][r.hap[a:b] == h].tolist())
            pp.append(len(pr))
            rl.append(lo[j])
            rh.append(hi[j])
    parts = ctx.process_reads_for_final_parts(fr, pp, pr, rl, rh, prm)
    hapq, rel, avg = ctx.get_hapq(fr, parts.part_ptr, parts.read_ids, parts.range_lo, parts.range_hi, g, prm)
    k = 1 + len(lo)
    assert out[k] == f"parts {parts.n_parts}"
    for i in range(parts.n_parts):
        assert out[k + 1 + i] == (f"part {i} range {parts.range_lo[i]}-{parts.range_hi[i]} reads "
                                  f"{int(parts.part_ptr[i + 1] - parts.part_ptr[i])} hapq {hapq[i]} rel {rel[i]:.17g}")
    assert out[-1] == f"avg_err {avg:.17g}"
    ctx.close()
