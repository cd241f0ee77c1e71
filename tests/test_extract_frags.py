"""tools/extract_frags.py (row f2: BAM + VCF -> fragments) on a hand-made BAM: every CIGAR operation, the flag / MAPQ
filter of alignment_passed_check (file_reader.rs:185-237), the single-base-allele filter of get_vcf_profile (:239-314)
and the allele / quality extraction of frag_from_record (:661-736), against cells worked out by hand."""
import gzip
import importlib.util
import os
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("extract_frags", os.path.join(ROOT, "tools", "extract_frags.py"))
ef = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ef)

CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
OPS = {c: i for i, c in enumerate("MIDNSHP=X")}


def bam_record(ref_id, pos, mapq, flag, name, cigar, seq, qual):
    cig = [(n << 4) | OPS[o] for n, o in cigar]
    packed = bytearray((len(seq) + 1) // 2)
    for i, b in enumerate(seq):
        packed[i >> 1] |= CODE[b] << (4 if i % 2 == 0 else 0)
    body = struct.pack("<iiBBHHHIiii", ref_id, pos, len(name) + 1, mapq, 0, len(cig), flag, len(seq), -1, -1, 0)
    body += name.encode() + b"\0" + struct.pack("<%dI" % len(cig), *cig) + bytes(packed) + bytes(qual)
    return struct.pack("<i", len(body)) + body


def write_bam(path, refs, records):
    data = b"BAM\1" + struct.pack("<i", 0) + struct.pack("<i", len(refs))
    for name, ln in refs:
        data += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln)
    data += b"".join(records)
    with gzip.open(path, "wb") as fh:  # BGZF is a series of gzip members; one member is a valid special case
        fh.write(data)


def test_extractor_on_a_hand_made_bam(tmp_path):
    vcf = tmp_path / "t.vcf"
    vcf.write_text(
        "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
        "c1\t11\t.\tA\tG\t.\t.\t.\n"      # SNP 1 at 0-based 10
        "c1\t14\t.\tC\tT,G\t.\t.\t.\n"    # SNP 2 at 13, triallelic
        "c1\t16\t.\tAT\tA\t.\t.\t.\n"     # indel: skipped, does not consume a SNP index
        "c1\t21\t.\tT\tC\t.\t.\t.\n"      # SNP 3 at 20
        "c1\t31\t.\tG\tA\t.\t.\t.\n"      # SNP 4 at 30
        "other\t5\t.\tA\tC\t.\t.\t.\n")
    prof = ef.read_vcf(str(vcf))
    assert prof["c1"][2] == [10, 13, 20, 30] and prof["c1"][1][13] == ["C", "T", "G"]
    pos_map, al_map, _ = prof["c1"]

    # read A: starts at 8, 2S 4M 1I 3M 2D 6M 5N 8M: reference walk 8..11 | ins | 12..14 | del 15,16 | 17..22 | skip 23..27 | 28..35
    seqA = "NN" + "ACGT" + "C" + "AGA" + "TTTCAA" + "CCACCCCC"
    qualA = list(range(10, 10 + len(seqA)))
    cigA = [(2, "S"), (4, "M"), (1, "I"), (3, "M"), (2, "D"), (6, "M"), (5, "N"), (8, "M")]
    cells = ef.frag_from_record(8, [(n << 4) | OPS[o] for n, o in cigA], _pack(seqA), bytes(qualA), pos_map, al_map)
    # SNP1 (ref 10): third M base = read index 2+2 = 4 -> 'G' = ALT (allele 1), quality 14
    # SNP2 (ref 13): second base of the 3M after the insertion: read index 2+4+1+1 = 8 -> 'G' = allele 2, quality 18
    # SNP3 (ref 20): 6M covers 17..22, read index 2+4+1+3 = 10 -> offset 3 -> index 13 -> 'C' = ALT (allele 1), quality 23
    # SNP4 (ref 30): 8M covers 28..35, read index 16 -> offset 2 -> index 18 -> 'A' = ALT (allele 1), quality 28
    assert cells == {1: (1, 14), 2: (2, 18), 3: (1, 23), 4: (1, 28)}

    # a base that matches no allele is dropped; a deletion over a SNP gives no cell; =/X behave like M
    seqB = "TAAAA"
    cells = ef.frag_from_record(10, [(1 << 4) | OPS["X"], (2 << 4) | OPS["D"], (4 << 4) | OPS["="]], _pack(seqB),
                                bytes([30] * 5), pos_map, al_map)
    assert cells == {}  # 'T' at SNP1 is neither A nor G; SNP2 (ref 13) is behind the read's aligned bases (13..16 = AAAA: 'A' no allele)

    # the record filter
    assert ef.passed(0, 60) and ef.passed(16, 15)
    assert not ef.passed(0, 14)          # MAPQ below 15
    assert not ef.passed(256, 60)        # secondary
    assert not ef.passed(4, 60) and not ef.passed(512, 60) and not ef.passed(1024, 60)  # unmapped / QC fail / duplicate
    assert not ef.passed(2048, 60)       # supplementary: dropped by this extractor (declared)

    # the BAM container
    bam = tmp_path / "t.bam"
    write_bam(str(bam), [("c1", 1000), ("other", 50)], [
        bam_record(0, 8, 60, 0, "readA", cigA, seqA, qualA),
        bam_record(0, 10, 3, 0, "lowq", [(5, "M")], "AAAAA", [30] * 5),
        bam_record(1, 0, 60, 0, "elsewhere", [(6, "M")], "ACACAC", [30] * 6)])
    recs = list(ef.bam_records(str(bam)))
    assert [r[4] for r in recs] == ["readA", "lowq", "elsewhere"] and recs[0][0] == "c1" and recs[2][0] == "other"
    assert recs[0][1] == 8 and recs[0][2] == 60 and recs[0][8] == len(seqA)
    assert ef.frag_from_record(recs[0][1], recs[0][5], recs[0][6], recs[0][7], pos_map, al_map) == {
        1: (1, 14), 2: (2, 18), 3: (1, 23), 4: (1, 28)}


def _pack(seq):
    packed = bytearray((len(seq) + 1) // 2)
    for i, b in enumerate(seq):
        packed[i >> 1] |= CODE[b] << (4 if i % 2 == 0 else 0)
    return bytes(packed)
