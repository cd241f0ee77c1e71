"""Builds floria_b200/libfloria_b200.so (CUDA, sm_100a) in-tree with nvcc.  `python -m floria_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfloria_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall", "--fmad=false", "-Xptxas", "-v",
]


def sources():
    """the translation units: compiled to objects in parallel, then linked into one shared library"""
    return [os.path.join(CSRC, f) for f in ("fb_lib.cu", "fb_beam_tu.cu", "fb_beam_wide_tu.cu", "fb_reader.cpp")]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    inc = os.path.join(HERE, "..", "include")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(inc, f) for f in os.listdir(inc)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(HERE, os.path.splitext(os.path.basename(src))[0] + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen([NVCC] + FLAGS + ["-c", "-o", obj, src], stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    log, failed = "", False
    for p in procs:
        out, err = p.communicate()
        log += out + err
        failed |= p.returncode != 0
    if not failed:
        res = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-pthread", "-o", OUT] + objs + ["-lz"],
                             capture_output=True, text=True)
        log += res.stdout + res.stderr
        failed = res.returncode != 0
    if verbose or failed:
        sys.stderr.write(log)
    if failed:
        raise RuntimeError("nvcc failed")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as fh:
        fh.write(log)
    return OUT


def build_hosttest(force=False):
    """CPU build of fb_seq.h's host-compilable logic (unit tests only)."""
    out = os.path.join(HERE, "libfb_hosttest.so")
    src = os.path.join(CSRC, "fb_hosttest.cpp")
    if force or not os.path.exists(out) or os.path.getmtime(out) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(CSRC, "fb_seq.h"))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", out, src])
    return out


if __name__ == "__main__":
    build_hosttest(force="--force" in sys.argv)
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
