"""bench.py --impl reference: the line the driver parses carries every key of the contract, and the arm never touches the
CUDA library (FB_LIB points at a file that does not exist: loading it would raise).  A two-read sample on two threads keeps
the run to seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(extra_env, args):
    env = dict(os.environ, FB_BENCH_CPU_READS="2", FB_BENCH_CPU_THREADS="2", FB_LIB="/nonexistent/libfloria_b200.so", **extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"] + args,
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_line_at_n1():
    lines = run_reference({}, [])
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["metric"] == "read x SNP cells scored per second" and d["unit"] == "cells/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("configs[2]") and d["config"]["n_reads"] == 100000 and d["config"]["ploidy"] == 4
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] == d["value"] and "slices of 2 consecutive" in cb["sample"]
    assert set(cb["rust_toolchain"]) == {"cargo", "rustc"} and "restatement" in cb["why_port"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_prints_on_rank_0_only():
    assert run_reference({"RANK": "1", "WORLD_SIZE": "2"}, ["--gpus", "2"]) == []


def test_reference_arm_line_at_n2_is_the_sharded_workload():
    lines = run_reference({"RANK": "0", "WORLD_SIZE": "2"}, ["--gpus", "2"])
    d = json.loads(lines[-1])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "strong"
    assert d["config"]["workload"].startswith("configs[4]") and d["config"]["n_contigs"] == 500
    assert d["cpu_baseline"]["cores"] == 2 and "contigs" in d["cpu_baseline"]["sample"] and d["value"] > 0
