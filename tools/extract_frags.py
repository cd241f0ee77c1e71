#!/usr/bin/env python
"""BASELINE.json configs[0]: turn the reference's own quick-start input (tests/test_long.bam + tests/test.vcf,
README.md:85) into the fragment arrays that cross the C-ABI, and commit them as a fixture (the reference tree does not
exist on the GPU box).  Run in the build container only:

    python tools/extract_frags.py /root/reference/tests/test_long.bam /root/reference/tests/test.vcf tests/golden/config0_long_frags.npz

Restates, in plain Python over zlib (no htslib / pysam in this image):
  * get_vcf_profile            src/file_reader.rs:239-314   single-base alleles only, 1-based SNP counter per contig
  * alignment_passed_check     src/file_reader.rs:185-237   MAPQ >= 15, flags & 1796 == 0, no secondary
  * frag_from_record           src/file_reader.rs:661-736   aligned pairs (M/=/X) at SNP positions; allele = index of the
                                                            read base among [REF, ALT...]; quality = BAM base quality
  * sort + counter_id          src/bin/floria.rs:289-293    Frag::cmp order
Declared differences from a real floria run (this is an INPUT fixture for the oracle-vs-CUDA parity of configs[0], not a
claim of byte equality with floria's own fragments): supplementary alignments are dropped (= `--dont-use-supp-aln`; 8 of
1240 records) instead of merged by combine_frags (file_reader.rs:491-659), and the local re-alignment around SNPs
(alignment.rs:7-64, block-aligner, active because the quick-start passes -r) is not restated."""
import gzip
import struct
import sys

import numpy as np

SEQ = "=ACMGRSVTWYHKDBN"


def read_vcf(path):
    """-> {contig: (gn_pos0 -> snp index 1.., gn_pos0 -> [allele bytes], [gn_pos0 per snp])}"""
    prof = {}
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as fh:
        for ln in fh:
            if ln.startswith("#"):
                continue
            f = ln.rstrip("\n").split("\t")
            alleles = [f[3]] + f[4].split(",")
            if any(len(a) != 1 or a.upper() not in "ACGT" for a in alleles):
                continue
            pos_map, al_map, order = prof.setdefault(f[0], ({}, {}, []))
            gn = int(f[1]) - 1  # rust-htslib Record::pos() is 0-based
            pos_map[gn] = len(order) + 1
            al_map[gn] = [a for a in alleles]
            order.append(gn)
    return prof


def bam_records(path):
    with gzip.open(path, "rb") as fh:
        data = fh.read()
    assert data[:4] == b"BAM\1"
    (l_text,) = struct.unpack_from("<i", data, 4)
    o = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", data, o)
    o += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, o)
        name = data[o + 4:o + 4 + l_name - 1].decode()
        o += 4 + l_name + 4
        refs.append(name)
    while o < len(data):
        (bs,) = struct.unpack_from("<i", data, o)
        ref_id, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, _nr, _np, _tl = struct.unpack_from("<iiBBHHHIiii", data, o + 4)
        p = o + 36
        name = data[p:p + l_rn - 1].decode()
        p += l_rn
        cigar = struct.unpack_from("<%dI" % n_cig, data, p)
        p += 4 * n_cig
        sq = data[p:p + (l_seq + 1) // 2]
        p += (l_seq + 1) // 2
        qual = data[p:p + l_seq]
        o += 4 + bs
        yield refs[ref_id] if ref_id >= 0 else None, pos, mapq, flag, name, cigar, sq, qual, l_seq


def passed(flag, mapq, mapq_cutoff=15):
    if flag & 2048:
        return False  # supplementary: dropped here (see the header)
    if mapq < mapq_cutoff or flag & 1796 or flag & 256:
        return False
    return True


def frag_from_record(pos, cigar, sq, qual, pos_map, al_map):
    cells = {}
    r, g = 0, pos
    for c in cigar:
        op, ln = c & 0xF, c >> 4
        if op in (0, 7, 8):  # M = X: aligned pairs with both coordinates
            for k in range(ln):
                gp = g + k
                if gp in pos_map:
                    b = sq[(r + k) >> 1]
                    base = SEQ[(b >> 4) if ((r + k) & 1) == 0 else (b & 0xF)]
                    for i, a in enumerate(al_map[gp]):
                        if base == a:
                            cells[pos_map[gp]] = (i, qual[r + k])
                            break
            r += ln
            g += ln
        elif op in (1, 4):  # I S
            r += ln
        elif op in (2, 3):  # D N
            g += ln
    return cells


def main():
    bam, vcf, out = sys.argv[1:4]
    prof = read_vcf(vcf)
    reads, lens = [], []
    contig = None
    n_rec = 0
    for ref, pos, mapq, flag, name, cigar, sq, qual, l_seq in bam_records(bam):
        n_rec += 1
        if ref is None or ref not in prof:
            continue
        contig = contig or ref
        if ref != contig or not passed(flag, mapq):
            continue
        lens.append(l_seq)
        pos_map, al_map, _ = prof[ref]
        cells = frag_from_record(pos, cigar, sq, qual, pos_map, al_map)
        if cells:
            ks = sorted(cells)
            reads.append((ks, [cells[k][0] for k in ks], [cells[k][1] for k in ks]))
    sys.path.insert(0, ".")
    from floria_b200.frags import Frags

    fr = Frags.from_reads(reads)  # Frag::cmp order
    g2p = np.array(prof[contig][2], dtype=np.uint64)
    lens = np.sort(np.array(lens))
    p66 = int(lens[int(len(lens) * 0.66)])
    np.savez_compressed(out, row_ptr=fr.row_ptr, pos=fr.pos, allele=fr.allele, qual=fr.qual, snp_to_genome_pos=g2p,
                        block_length=np.uint32(max(p66, 500)), contig=np.array(contig))
    print(f"{n_rec} BAM records, {len(lens)} passed, {fr.n_reads} fragments with SNPs, {fr.nnz} cells, {len(g2p)} SNPs, "
          f"contig {contig}, p66 read length {p66}")


if __name__ == "__main__":
    main()
