"""fb_read_frags (include/floria_b200_reader.h, row f2 in C++ behind the C ABI) on hand-made BAM / VCF files: the cells of
frag_from_record worked out by hand (every CIGAR operation), the record filter of alignment_passed_check, the mate merge and
the supplementary-alignment merge of combine_frags (file_reader.rs:491-659), Frag::cmp order; against the Python extractor
(tools/extract_frags.py, an independent restatement) record by record; and, where the reference tree is present (this
container, not the GPU box), on floria's own tests/test_long.bam against the committed fixture."""
import os

import numpy as np
import pytest

from floria_b200 import api
from test_extract_frags import bam_record, ef, write_bam

VCF = ("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
       "c1\t11\t.\tA\tG\t.\t.\t.\n"      # SNP 1 at 0-based 10
       "c1\t14\t.\tC\tT,G\t.\t.\t.\n"    # SNP 2 at 13, triallelic
       "c1\t16\t.\tAT\tA\t.\t.\t.\n"     # indel: skipped, does not consume a SNP index
       "c1\t21\t.\tT\tC\t.\t.\t.\n"      # SNP 3 at 20
       "c1\t31\t.\tG\tA\t.\t.\t.\n"      # SNP 4 at 30
       "c1\t50041\t.\tA\tT\t.\t.\t.\n"   # SNP 5 at 50040: farther than supp_aln_dist_cutoff from SNP 4
       "nobam\t5\t.\tA\tC\t.\t.\t.\n"    # a contig the BAM does not have
       "other\t5\t.\tA\tC\t.\t.\t.\n")

SEQ_A = "NN" + "ACGT" + "C" + "AGA" + "TTTCAA" + "CCACCCCC"
CIG_A = [(2, "S"), (4, "M"), (1, "I"), (3, "M"), (2, "D"), (6, "M"), (5, "N"), (8, "M")]


def reads_of(fr):
    return [tuple(map(lambda a: a.tolist(), fr.read(i))) for i in range(fr.n_reads)]


def test_cells_filter_and_order(tmp_path):
    vcf = tmp_path / "t.vcf"
    vcf.write_text(VCF)
    qual_a = list(range(10, 10 + len(SEQ_A)))
    bam = tmp_path / "t.bam"
    write_bam(str(bam), [("c1", 100000), ("other", 50)], [
        bam_record(0, 8, 60, 0, "readA", CIG_A, SEQ_A, qual_a),                       # the hand-traced read of test_extract_frags
        bam_record(0, 10, 3, 0, "lowq", [(5, "M")], "GAAAA", [30] * 5),               # MAPQ < 15
        bam_record(0, 10, 60, 256, "secondary", [(5, "M")], "GAAAA", [30] * 5),
        bam_record(0, 10, 60, 1024, "duplicate", [(5, "M")], "GAAAA", [30] * 5),
        bam_record(0, 10, 60, 16, "readB", [(1, "X"), (2, "D"), (4, "=")], "TAAAA", [30] * 5),   # no cell: kept apart
        bam_record(0, 9, 60, 0, "readC", [(6, "M")], "AGAACT", [40, 41, 42, 43, 44, 45]),        # SNP1 'G' (1), SNP2 'T' (1)
        bam_record(0, 10, 60, 0, "readD", [(25, "M")], "A" * 3 + "C" + "A" * 6 + "T" + "A" * 14, list(range(25))),
        bam_record(1, 0, 60, 0, "elsewhere", [(6, "M")], "ACACAC", [30] * 6)])
    fr, g2p, info = api.read_frags(str(bam), str(vcf))
    assert info["contig"] == "c1" and info["n_records"] == 7 and info["n_passed"] == 4 and info["n_without_snps"] == 1
    assert g2p.tolist() == [10, 13, 20, 30, 50040]
    # readA: {1: (1, 14), 2: (2, 18), 3: (1, 23), 4: (1, 28)} (worked out in test_extract_frags); readC: ref 10 is read index 1
    # ('G' = allele 1, q 41), ref 13 is index 4 ('C' = REF, allele 0, q 44); readD starts at 10: SNP1 'A' = REF (0, q 0),
    # SNP2 index 3 'C' = REF (0, q 3), SNP3 index 10 'T' = REF (0, q 10), SNP4 (ref 30) index 20 'A' = ALT (1, q 20)
    # Frag::cmp: first position ascending, then LAST position descending, then record order: readA (1..4), readD (1..4), readC (1..2)
    assert reads_of(fr) == [([1, 2, 3, 4], [1, 2, 1, 1], [14, 18, 23, 28]),
                            ([1, 2, 3, 4], [0, 0, 0, 1], [0, 3, 10, 20]),
                            ([1, 2], [1, 0], [41, 44])]
    assert fr.first.tolist() == [1, 1, 1] and fr.last.tolist() == [4, 4, 2]
    assert info["read_len_p66"] == sorted([len(SEQ_A), 5, 6, 25])[int(4 * 0.66)]
    # an explicit contig, and one without SNP-covering reads
    fr2, g2, info2 = api.read_frags(str(bam), str(vcf), contig="other")
    assert info2["contig"] == "other" and fr2.n_reads == 1 and g2.tolist() == [4] and reads_of(fr2) == [([1], [0], [30])]
    with pytest.raises(api.FloriaB200Error, match="not in the BAM header"):
        api.read_frags(str(bam), str(vcf), contig="nobam")
    with pytest.raises(api.FloriaB200Error, match="cannot read BAM"):
        api.read_frags(str(tmp_path / "missing.bam"), str(vcf))


def test_mates_and_supplementary_alignments(tmp_path):
    vcf = tmp_path / "t.vcf"
    vcf.write_text(VCF)
    bam = tmp_path / "t.bam"
    recs = [
        # a proper pair, second mate first in the file: both cover SNP 2 (ref 13), the SECOND mate's call wins (extend)
        bam_record(0, 12, 60, 1 | 2 | 128, "pair", [(10, "M")], "ATAAAAAACA", list(range(50, 60))),   # SNP2 'T'(1) q51, SNP3 'C'(1) q58
        bam_record(0, 9, 60, 1 | 2 | 64, "pair", [(6, "M")], "AGAAGA", list(range(20, 26))),           # SNP1 'G'(1) q21, SNP2 'G'(2) q24
        # a long read split in two: primary covers SNP 1, the supplementary part (MAPQ 60) covers SNPs 3-4 -> merged
        bam_record(0, 28, 60, 2048, "split", [(5, "H"), (6, "M")], "AAAAAA", [33] * 6),                # SNP4 'A'(1)
        bam_record(0, 8, 60, 0, "split", [(6, "M"), (6, "H")], "AAGAAA", [34] * 6),                    # SNP1 'G'(1), SNP2 'A' no allele
        # supplementary part too far away (SNP 5, 50 kb): the primary alone is kept
        bam_record(0, 50038, 60, 2048, "far", [(5, "M")], "AATAA", [35] * 5),                          # SNP5 'T'(1)
        bam_record(0, 18, 60, 0, "far", [(5, "M")], "AACAA", [36] * 5),                                # SNP3 'C'(1)
        # supplementary with MAPQ < 60 is filtered before the merge
        bam_record(0, 28, 59, 2048, "weak", [(6, "M")], "AAAAAA", [37] * 6),
        bam_record(0, 10, 60, 0, "weak", [(2, "M")], "GA", [38] * 2),                                  # SNP1 'G'(1)
        # only the supplementary alignment survives (primary MAPQ 3): nothing is emitted
        bam_record(0, 28, 60, 2048, "orphan", [(6, "M")], "AAAAAA", [39] * 6),
        bam_record(0, 10, 3, 0, "orphan", [(2, "M")], "GA", [39] * 2),
    ]
    write_bam(str(bam), [("c1", 100000)], recs)
    fr, g2p, info = api.read_frags(str(bam), str(vcf))
    got = reads_of(fr)
    # pair: mate 1 {1: (1, 21), 2: (2, 24)} extended by mate 2 {2: (1, 51), 3: (1, 58)} -> SNP 2 takes mate 2's call
    assert ([1, 2, 3], [1, 1, 1], [21, 51, 58]) in got
    assert ([1, 4], [1, 1], [34, 33]) in got          # split: primary + supplementary
    assert ([3], [1], [36]) in got                    # far: primary only
    assert ([1], [1], [38]) in got                    # weak
    assert len(got) == 4 and info["n_passed"] == 8
    # Frag::cmp order: (1..4) split, (1..3) pair, (1..1) weak, (3..3) far
    assert got == [([1, 4], [1, 1], [34, 33]), ([1, 2, 3], [1, 1, 1], [21, 51, 58]), ([1], [1], [38]), ([3], [1], [36])]
    # --dont-use-supp-aln: supplementary records are dropped by the filter
    fr2, _, info2 = api.read_frags(str(bam), str(vcf), use_supp_aln=False)
    assert ([1], [1], [34]) in reads_of(fr2) and ([1, 4], [1, 1], [34, 33]) not in reads_of(fr2) and info2["n_passed"] == 5
    # a larger cutoff merges the far part too
    fr3, _, _ = api.read_frags(str(bam), str(vcf), supp_aln_dist_cutoff=100000)
    assert ([3, 5], [1, 1], [36, 35]) in reads_of(fr3)


def test_equals_the_python_extractor_on_random_reads(tmp_path):
    """an independent restatement (tools/extract_frags.py) on 300 random reads with random CIGARs: same fragments"""
    rng = np.random.default_rng(5)
    L = 4000
    snp_pos = np.sort(rng.choice(np.arange(5, L - 5), 300, replace=False))
    lines = ["##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"]
    for p in snp_pos:
        ref, alt = rng.choice(list("ACGT"), 2, replace=False)
        lines.append(f"c1\t{p + 1}\t.\t{ref}\t{alt}\t.\t.\t.\n")
    vcf = tmp_path / "r.vcf"
    vcf.write_text("".join(lines))
    recs = []
    for i in range(300):
        start = int(rng.integers(0, L - 600))
        cig, qlen = [], 0
        for _ in range(int(rng.integers(1, 6))):
            op = str(rng.choice(list("MMMMIDNSX=")))
            n = int(rng.integers(1, 120))
            cig.append((n, op))
            if op in "MIS=X":
                qlen += n
        if not any(o in "M=X" for _, o in cig):
            cig.append((50, "M"))
            qlen += 50
        seq = "".join(rng.choice(list("ACGT"), qlen))
        qual = rng.integers(0, 60, qlen).tolist()
        flag = int(rng.choice([0, 16, 256, 1024, 4]))
        recs.append(bam_record(0, start, int(rng.choice([3, 20, 60])), flag, f"r{i}", cig, seq, qual))
    bam = tmp_path / "r.bam"
    write_bam(str(bam), [("c1", L)], recs)
    fr, g2p, info = api.read_frags(str(bam), str(vcf), use_supp_aln=False)
    prof = ef.read_vcf(str(vcf))
    pos_map, al_map, order = prof["c1"]
    want = []
    for ref, pos, mapq, flag, name, cigar, sq, qual, l_seq in ef.bam_records(str(bam)):
        if not ef.passed(flag, mapq):
            continue
        cells = ef.frag_from_record(pos, cigar, sq, qual, pos_map, al_map)
        if cells:
            ks = sorted(cells)
            want.append((ks, [cells[k][0] for k in ks], [cells[k][1] for k in ks]))
    from floria_b200.frags import Frags

    w = Frags.from_reads(want)
    assert g2p.tolist() == order
    assert fr.n_reads == w.n_reads > 50
    for a, b in ((fr.row_ptr, w.row_ptr), (fr.pos, w.pos), (fr.allele, w.allele), (fr.qual, w.qual), (fr.first, w.first),
                 (fr.last, w.last)):
        assert np.array_equal(a, b)


REF_BAM = "/root/reference/tests/test_long.bam"
REF_VCF = "/root/reference/tests/test.vcf"
FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config0_long_frags.npz")


@pytest.mark.skipif(not (os.path.exists(REF_BAM) and os.path.exists(REF_VCF) and os.path.exists(FIXTURE)),
                    reason="the reference tree exists in the build container only")
def test_floria_quick_start_data_equals_the_committed_fixture():
    """BASELINE.json configs[0]: floria's own tests/test_long.bam + tests/test.vcf; with supplementary alignments dropped the
    C++ reader gives the committed fixture bit for bit; with them merged (floria's default) 8 more records take part"""
    z = np.load(FIXTURE)
    fr, g2p, info = api.read_frags(REF_BAM, REF_VCF, use_supp_aln=False)
    assert info["contig"] == str(z["contig"])
    for k, a in (("row_ptr", fr.row_ptr), ("pos", fr.pos), ("allele", fr.allele), ("qual", fr.qual), ("snp_to_genome_pos", g2p)):
        assert np.array_equal(a, z[k]), k
    assert max(info["read_len_p66"], 500) == int(z["block_length"])
    fr2, _, info2 = api.read_frags(REF_BAM, REF_VCF)
    assert info2["n_passed"] >= info["n_passed"] and fr2.nnz >= fr.nnz and fr2.n_reads <= fr.n_reads + 8
