"""Synthetic read x SNP blocks of the shapes BASELINE.json names.

Counter-based PRNG: value(stream, idx) = splitmix64_mix(seed + stream*K + (idx+1)*GAMMA), so every cell is a
pure function of (seed, stream, index) and the same input can be regenerated in any order or language.
Seeds follow BASELINE.md: 0xF10A1A00 + config number.
"""
import numpy as np

from .frags import Frags

GAMMA = np.uint64(0x9E3779B97F4A7C15)
STREAM_K = np.uint64(0xD1B54A32D192ED03)
SEED_BASE = 0xF10A1A00


def _mix(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rng_u64(seed, stream, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + np.uint64(stream) * STREAM_K
        return _mix(base + (idx + np.uint64(1)) * GAMMA)


def rng_u01(seed, stream, idx):
    return (rng_u64(seed, stream, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def rng_int(seed, stream, idx, lo, hi):
    """uniform integer in [lo, hi]"""
    n = np.uint64(hi - lo + 1)
    return (rng_u64(seed, stream, idx) % n).astype(np.int64) + lo


class Contig:
    def __init__(self, frags, snp_to_genome_pos, truth_hap, truth_alleles, read_hap):
        self.frags = frags
        self.snp_to_genome_pos = snp_to_genome_pos
        self.truth_alleles = truth_alleles  # [ploidy, n_snps]
        self.read_hap = read_hap  # true source haplotype per (sorted) read
        self.truth_hap = truth_hap


def make_truth(seed, ploidy, n_snps):
    """p haplotypes over S SNPs; columns from {0,1} (5 % triallelic {0,1,2}), resampled until polymorphic."""
    s = np.arange(n_snps, dtype=np.uint64)
    tri = rng_u01(seed, 1, s) < 0.05
    nall = np.where(tri, 3, 2).astype(np.int64)
    truth = np.zeros((ploidy, n_snps), dtype=np.uint8)
    done = np.zeros(n_snps, dtype=bool)
    for attempt in range(64):
        cand = np.zeros((ploidy, n_snps), dtype=np.uint8)
        for h in range(ploidy):
            r = rng_u64(seed, 2 + attempt, s * np.uint64(ploidy) + np.uint64(h))
            cand[h] = (r % nall.astype(np.uint64)).astype(np.uint8)
        poly = (cand != cand[0:1]).any(axis=0) if ploidy > 1 else np.ones(n_snps, dtype=bool)
        take = ~done & poly
        truth[:, take] = cand[:, take]
        done |= take
        if done.all():
            break
    return truth, nall


def make_contig(
    config_seed,
    n_reads,
    n_snps,
    ploidy,
    span_mean=500,
    span_sigma=0.5,
    full_span=False,
    present=0.98,
    flip=0.04,
    qual_mode="long",
    paired_short=False,
    snp_spacing=100,
):
    seed = SEED_BASE + config_seed
    truth, nall = make_truth(seed, ploidy, n_snps)
    rid = np.arange(n_reads, dtype=np.uint64)
    # source haplotype ~ abundances 1/k
    w = 1.0 / np.arange(1, ploidy + 1)
    cdf = np.cumsum(w / w.sum())
    src = np.searchsorted(cdf, rng_u01(seed, 100, rid), side="right").clip(0, ploidy - 1)

    if paired_short:
        # mates of 1-3 SNPs separated by a gap of 2-5 SNPs
        l1 = rng_int(seed, 101, rid, 1, 3)
        gap = rng_int(seed, 102, rid, 2, 5)
        l2 = rng_int(seed, 103, rid, 1, 3)
        span = l1 + gap + l2
        span = np.minimum(span, n_snps)
    elif full_span:
        span = np.full(n_reads, n_snps, dtype=np.int64)
    else:
        u1 = np.maximum(rng_u01(seed, 101, rid), 1e-300)
        u2 = rng_u01(seed, 102, rid)
        z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
        span = np.rint(span_mean * np.exp(span_sigma * z - 0.5 * span_sigma * span_sigma)).astype(np.int64)
        span = span.clip(1, n_snps)
    first = (rng_u64(seed, 104, rid) % (np.uint64(n_snps) - span.astype(np.uint64) + np.uint64(1))).astype(
        np.int64
    ) + 1
    last = first + span - 1

    order = np.lexsort((np.arange(n_reads), -last, first))  # Frag::cmp
    first, last, span, src, orig = first[order], last[order], span[order], src[order], rid[order]
    if paired_short:
        l1, gap, l2 = l1[order], gap[order], l2[order]

    # one candidate cell per position of the span
    tot = int(span.sum())
    row_of = np.repeat(np.arange(n_reads), span)
    starts = np.zeros(n_reads + 1, dtype=np.int64)
    starts[1:] = np.cumsum(span)
    off = np.arange(tot, dtype=np.int64) - starts[row_of]
    pos = first[row_of] + off  # 1-based SNP index
    key = orig[row_of] * np.uint64(1 << 20) + off.astype(np.uint64)
    keep = rng_u01(seed, 200, key) < present
    keep |= (off == 0) | (off == span[row_of] - 1)  # first/last position are always covered
    if paired_short:
        in_gap = (off >= l1[row_of]) & (off < (l1 + gap)[row_of])
        keep = ~in_gap
        keep &= off < span[row_of]
    t = truth[src[row_of], pos - 1]
    na = nall[pos - 1]
    do_flip = rng_u01(seed, 201, key) < flip
    shift = 1 + (rng_u64(seed, 202, key) % (na.astype(np.uint64) - np.uint64(1))).astype(np.int64)
    allele = np.where(do_flip, (t.astype(np.int64) + shift) % na, t).astype(np.uint8)
    if qual_mode == "long":
        qual = rng_int(seed, 203, key, 5, 40).astype(np.uint8)
    elif qual_mode == "short":
        u = rng_u01(seed, 203, key)
        qual = np.select([u < 0.05, u < 0.15, u < 0.40], [2, 12, 23], 37).astype(np.uint8)
    else:
        raise ValueError(qual_mode)

    pos, allele, qual, row_keep = pos[keep], allele[keep], qual[keep], row_of[keep]
    counts = np.bincount(row_keep, minlength=n_reads)
    row_ptr = np.zeros(n_reads + 1, dtype=np.uint64)
    row_ptr[1:] = np.cumsum(counts)
    frags = Frags(row_ptr, pos.astype(np.uint32), allele, qual, first.astype(np.uint32), last.astype(np.uint32))
    s = np.arange(n_snps, dtype=np.uint64)
    snp_to_genome_pos = (snp_spacing * s + rng_int(seed, 300, s, 0, snp_spacing // 2 - 1).astype(np.uint64)).astype(
        np.uint64
    )
    return Contig(frags, snp_to_genome_pos, src, truth, src)


# ---- the named configs (BASELINE.json "configs") ---------------------------------------------------------------
def config2(scale=1.0):
    """1 contig, 10k long-read frags x 5k SNPs, ploidy 2."""
    return make_contig(2, int(10000 * scale), int(5000 * scale), 2, span_mean=500, flip=0.04, qual_mode="long")


def config3_banded(n_reads=100000, n_snps=50000, ploidy=4, span_mean=5000):
    return make_contig(3, n_reads, n_snps, ploidy, span_mean=span_mean, flip=0.04, qual_mode="long")


def config4(scale=1.0):
    """short-read, 2M frags x 100k SNPs, ploidy 3."""
    return make_contig(
        4, int(2000000 * scale), int(100000 * scale), 3, paired_short=True, flip=0.01, qual_mode="short"
    )


def config5_contig(k, n_reads=2000, n_snps=1000, span_mean=100):
    """contig k of the 500-contig metagenome: S~1000, R~2000, ploidy ~ U{2..6}."""
    ploidy = int(rng_int(SEED_BASE + 5, 7, np.array([k]), 2, 6)[0])
    c = make_contig(5 * 100000 + k, n_reads, n_snps, ploidy, span_mean=span_mean, flip=0.04, qual_mode="long")
    c.ploidy = ploidy
    return c
