"""Output formats (floria_b200/writers.py) against hand-written expectations derived from file_writer.rs, and the
local_parts dump on oracle output (the artefact tools/pin_with_floria.sh diffs against a real floria --debug run)."""
import os

import numpy as np

import oracle
from floria_b200 import default_params, synth, writers
from floria_b200.frags import Frags


def tiny():
    # read: (positions, alleles, quals)
    reads = [([1, 2, 3], [0, 1, 0], [30, 30, 30]), ([1, 2, 3], [0, 1, 1], [20, 20, 20]), ([2, 3, 4], [1, 0, 2], [10, 10, 10]),
             ([5], [1], [40])]
    fr = Frags.from_reads(reads)
    fr.names = [f"r{i}" for i in fr.order]
    return fr


def test_all_parts_file_debug_form(tmp_path):
    fr = tiny()
    p = tmp_path / "local_parts" / "0-0-1-2.haplosets"
    writers.write_all_parts_file(str(p), fr, [np.array([2, 0]), np.array([], np.int64), np.array([1])])
    # empty sets are skipped but keep their index; reads sorted by Frag::cmp; `id\tfirst\tlast`
    assert p.read_text() == f"#0\n{fr.names[0]}\t1\t3\n{fr.names[2]}\t2\t4\n#2\n{fr.names[1]}\t1\t3\n"


def test_errors_cov_running_sum_quirk():
    # counts {0:3, 1:5, 2:4} at one position: ascending order compares 3 > 0, 5 > 3, 4 > 8 (running SUM): max_count = 5
    reads = [([1], [0], [30])] * 3 + [([1], [1], [30])] * 5 + [([1], [2], [30])] * 4
    fr = Frags.from_reads(reads)
    cov, err, errors, support = writers.get_errors_cov_from_frags(fr, np.arange(12), 1, 1)
    assert (cov, errors, support) == (12.0, 7.0, 12.0) and err == 7.0 / 12.0
    # {0:5, 1:3, 2:4}: 5 > 0 yes, 3 > 5 no, 4 > 8 no -> 5; {0:2, 1:2, 2:5}: 2, then 2 > 2 no, 5 > 4 yes -> 5
    reads = [([1], [0], [30])] * 2 + [([1], [1], [30])] * 2 + [([1], [2], [30])] * 5
    fr = Frags.from_reads(reads)
    assert writers.get_errors_cov_from_frags(fr, np.arange(9), 1, 1)[2] == 4.0


def test_final_haplosets_and_vartigs(tmp_path):
    fr = tiny()
    g = np.array([100, 250, 300, 420, 900], np.uint64)
    parts = [np.array([0, 1, 2]), np.array([3])]
    ranges = [(1, 4), (5, 5)]
    out = tmp_path / "ctg"
    writers.write_all_parts_file(str(out / "ctg.haplosets"), fr, parts, contig="ctg", ranges=ranges, out_dir_label="D",
                                 snp_to_genome_pos=g, hapqs=[33, 0], rel_err=[0.5, 2.0])
    txt = (out / "ctg.haplosets").read_text().split("\n")
    # part 0: supports per position 1..4 = 2, 3, 3, 1; majority 2, 3, 2 (0:2 vs 1:1), 1 -> errors 1; cov = 9 / 4
    assert txt[0] == ">HAP0.D\tCONTIG:ctg\tSNPRANGE:1-4\tBASERANGE:101-421\tCOV:2.250\tERR:0.1111\tHAPQ:33\tREL_ERR:0.500"
    assert txt[4] == ">HAP1.D\tCONTIG:ctg\tSNPRANGE:5-5\tBASERANGE:901-901\tCOV:1.000\tERR:0.0000\tHAPQ:0\tREL_ERR:2.000"
    row = writers.write_vartigs(str(out), fr, parts, ranges, "ctg", g, [33, 0], [0.5, 2.0], 0.0625, 1000,
                                top_dir=str(tmp_path))
    v = (out / "ctg.vartigs").read_text().split("\n")
    assert v[1] == "0102" and v[3] == "1"  # consensus alleles of SNPs 1..4 and of SNP 5
    info = (out / "vartig_info.txt").read_text().split("\n")
    assert info[0] == f">HAP0.{out}\tSNPRANGE:1-4" and info[1] == "1:100\t0\t0:2\t" and info[3] == "3:300\t0\t0:2|1:1\t"
    # straincount = 5 covered SNP slots / 5 SNPs; multiplicity = (320 + 0) / 1000; coverage = (4*2.25 + 1) / 5
    assert row == "ctg\t1.000\t0.320\t2.000\t320\t0.800\t0.800\t0.000\t0.0625\n"
    assert (tmp_path / "contig_ploidy_info.tsv").read_text() == writers.CONTIG_PLOIDY_HEADER + row


def test_local_parts_dump_of_oracle_output(tmp_path):
    c = synth.make_contig(99, 300, 260, 2, span_mean=60)
    prm = default_params(epsilon=0.04, max_ploidy=3)
    lo, hi = oracle.get_range_with_lengths(c.snp_to_genome_pos, 10000, 10000 // 3, 0.0005)
    r = oracle.phase_blocks(c.frags, lo, hi, prm, n_threads=2)
    files = writers.write_local_parts(str(tmp_path), c.frags, r, lo)
    assert len(files) == int((r.best_ploidy > 0).sum())
    j = 0
    name = os.path.basename(files[0])
    assert name == f"{j}-0-{int(lo[j])}-{int(r.best_ploidy[j])}.haplosets"
    lines = open(files[0]).read().split("\n")
    n_reads = int(r.read_ptr[1] - r.read_ptr[0])
    assert sum(1 for x in lines if x.startswith("#")) <= int(r.best_ploidy[0])
    assert sum(1 for x in lines if x and not x.startswith("#")) == n_reads
