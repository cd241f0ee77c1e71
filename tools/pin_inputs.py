"""Inputs and our side of tools/pin_with_floria.sh: for every pinning case the H-PoP fragment file + SNP coordinate file that
the Rust harness reads, and the local_parts tree produced by this repository (the CPU oracle here, so that the script runs
on a machine without a GPU; the CUDA path is bit-identical to the oracle by the -m gpu tests).
python tools/pin_inputs.py <out_dir>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from floria_b200 import default_params, synth, writers  # noqa: E402
from floria_b200.frags import Frags  # noqa: E402


def cases():
    z = np.load(os.path.join(ROOT, "tests", "golden", "config0_long_frags.npz"), allow_pickle=True)
    fr = Frags(z["row_ptr"], z["pos"], z["allele"], z["qual"])
    yield "config0", fr, z["snp_to_genome_pos"], None, dict(epsilon=0.04, max_ploidy=5, block_length=int(z["block_length"]))
    c = synth.config2(0.1)
    yield "synth", c.frags, c.snp_to_genome_pos, None, dict(epsilon=0.04, max_ploidy=3, block_length=10000)


def main(out):
    os.makedirs(out, exist_ok=True)
    order_model = int(os.environ.get("FB_ORDER_MODEL", "0"))
    for name, fr, g, names, kw in cases():
        names = names or [f"read{i}" for i in range(fr.n_reads)]
        fr.write_hpop(os.path.join(out, f"{name}.hpop"), ids=names)
        with open(os.path.join(out, f"{name}.snps"), "w") as fh:
            fh.write("\n".join(str(int(x)) for x in g) + "\n")
        with open(os.path.join(out, f"{name}.params"), "w") as fh:
            fh.write(f"{kw['epsilon']} {kw['max_ploidy']} {kw['block_length']}\n")
        prm = default_params(order_model=order_model, **kw)
        lo, hi = oracle.get_range_with_lengths(g, kw["block_length"], kw["block_length"] // 3, 0.0005)
        r = oracle.phase_blocks(fr, lo, hi, prm, n_threads=os.cpu_count() or 1)
        files = writers.write_local_parts(os.path.join(out, "ours", name), fr, r, lo, names)
        print(f"{name}: {fr.n_reads} reads, {len(lo)} blocks, {len(files)} local_parts files")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "pin"))
