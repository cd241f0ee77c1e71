"""Output formats of the path, byte for byte as floria writes them, so that results can be DIFFED
against a real floria run (tools/pin_with_floria.sh):

  local_parts/<j>-<l>-<snp_lo>-<ploidy>.haplosets   the --debug dump of every block's best partition
                                                      (graph_processing.rs:289-300 -> file_writer.rs:919-993, no ranges)
  <contig>.haplosets                                  final haplosets with COV / ERR / HAPQ / REL_ERR headers
                                                      (file_writer.rs:919-993 with ranges)
  <contig>.vartigs, vartig_info.txt                   consensus alleles of every haploset (file_writer.rs:699-815, 306-372)
  contig_ploidy_info.tsv                              one row per contig (file_writer.rs:818-917, header constants.rs:24)

Host-side text formatting only: the numbers come from the C-ABI calls (fb_phase_blocks, fb_process_reads_for_final_parts,
fb_get_hapq); the unweighted allele counts needed for COV / ERR and the consensus strings are recomputed here with numpy
from the CSR reads (a host restatement of utils_frags.rs:596-657 get_errors_cov_from_frags and of the max_by_key rule,
canonical order: ascending allele, so the LAST maximal allele wins ties as in Iterator::max_by_key)."""
import os

import numpy as np

CONTIG_PLOIDY_HEADER = ("contig\taverage_straincount\twhole_contig_multiplicity\tapproximate_coverage_ignoring_indels\t"
                        "total_vartig_bases_covered\taverage_straincount_min15hapq\taverage_straincount_min30hapq\t"
                        "average_straincount_min45hapq\tavg_err\n")  # constants.rs:24


def _names(frags, names):
    if names is not None:
        return names
    n = getattr(frags, "names", None)
    return n if n is not None else [f"read{i}" for i in range(frags.n_reads)]


def _sorted_part(frags, ids):
    """Frag::cmp order (types_structs.rs:87-93) == ascending counter_id for a contig sorted by Frag::cmp"""
    return np.sort(np.asarray(ids, dtype=np.int64))


def allele_counts(frags, ids, lo, hi):
    """set_to_seq_dict(frags, false) restricted to SNP positions lo..hi (1-based, inclusive): counts [hi-lo+1, 4]"""
    n = int(hi) - int(lo) + 1
    cnt = np.zeros((max(n, 0), 4), np.int64)
    for r in ids:
        p, a, _ = frags.read(int(r))
        m = (p >= lo) & (p <= hi)
        np.add.at(cnt, (p[m].astype(np.int64) - int(lo), a[m].astype(np.int64)), 1)
    return cnt


def get_errors_cov_from_frags(frags, ids, lo, hi):
    """utils_frags.rs:596-657 (use_phred = false, mean = true).  Canonical allele order (ascending): the running-sum
    quirk of lines 616-624 (`if *count > snp_support { max_count_pos = *count }`) is reproduced as written."""
    cnt = allele_counts(frags, ids, lo, hi)
    errors = total_support = 0.0
    nonzero = 0
    supports = []
    for row in cnt:
        snp_support = 0.0
        max_count_pos = 0.0
        if row.any():
            nonzero += 1
            for c in row:
                if c == 0:
                    continue  # absent key
                if float(c) > snp_support:
                    max_count_pos = float(c)
                snp_support += float(c)
        total_support += snp_support
        errors += snp_support - max_count_pos
        supports.append(snp_support)
    cov = (sum(sorted(supports)) / nonzero) if nonzero > 0 else 0.0
    err = errors / total_support if total_support != 0 else float("nan")
    return cov, err, errors, total_support


def write_all_parts_file(path, frags, parts, names=None, contig="", ranges=None, out_dir_label="",
                         snp_to_genome_pos=None, hapqs=None, rel_err=None):
    """file_writer.rs:919-993.  parts: list of counter_id arrays; ranges empty/None -> the `#i` debug form."""
    names = _names(frags, names)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as fh:
        for i, ids in enumerate(parts):
            if len(ids) == 0:
                continue
            ids = _sorted_part(frags, ids)
            if not ranges:
                fh.write(f"#{i}\n")
            else:
                lo, hi = int(ranges[i][0]), int(ranges[i][1])
                cov, err, _, _ = get_errors_cov_from_frags(frags, ids, lo, hi)
                fh.write(f">HAP{i}.{out_dir_label}\tCONTIG:{contig}\tSNPRANGE:{lo}-{hi}\t"
                         f"BASERANGE:{int(snp_to_genome_pos[lo - 1]) + 1}-{int(snp_to_genome_pos[hi - 1]) + 1}\t"
                         f"COV:{cov:.3f}\tERR:{err:.4f}\tHAPQ:{int(hapqs[i])}\tREL_ERR:{rel_err[i]:.3f}\n")
            for r in ids:
                fh.write(f"{names[int(r)]}\t{int(frags.first[r])}\t{int(frags.last[r])}\n")


def write_local_parts(out_dir, frags, block_results, blk_lo, names=None):
    """graph_processing.rs:289-300: one `<j>-<l>-<snp_lo>-<best_ploidy>.haplosets` per phased block (l is always 0:
    WEIRD_SPLIT = false leaves one partition per ploidy, graph_processing.rs:166-185)"""
    d = os.path.join(out_dir, "local_parts")
    os.makedirs(d, exist_ok=True)
    written = []
    for j in range(block_results.n_blocks):
        p = int(block_results.best_ploidy[j])
        if p == 0:
            continue  # get_local_hap_blocks returned None (no reads in the interval)
        a, b = int(block_results.read_ptr[j]), int(block_results.read_ptr[j + 1])
        ids, hap = block_results.read_ids[a:b], block_results.hap[a:b]
        parts = [ids[hap == h] for h in range(p)]
        path = os.path.join(d, f"{j}-0-{int(blk_lo[j])}-{p}.haplosets")
        write_all_parts_file(path, frags, parts, names)
        written.append(path)
    return written


def consensus_alleles(frags, ids, lo, hi):
    """file_writer.rs:306-372 write_fragset_haplotypes: per position lo..hi the allele with the largest unweighted count
    (max_by_key: the last maximum in iteration order; canonical order = ascending allele), 15 ('?') without coverage.
    Returns (alleles uint8 [hi-lo+1], counts [hi-lo+1, 4])."""
    cnt = allele_counts(frags, ids, lo, hi)
    best = np.full(len(cnt), 15, np.uint8)
    for k, row in enumerate(cnt):
        if row.any():
            m = row.max()
            best[k] = np.nonzero(row == m)[0][-1]
    return best, cnt


def write_vartigs(out_dir, frags, parts, ranges, contig, snp_to_genome_pos, hapqs, rel_err, avg_err, contig_len,
                  top_dir=None):
    """file_writer.rs:699-917: <contig>.vartigs, vartig_info.txt and the contig's row of contig_ploidy_info.tsv"""
    os.makedirs(out_dir, exist_ok=True)
    top_dir = top_dir or out_dir
    n_snps = len(snp_to_genome_pos)
    covered = np.zeros((4, n_snps))  # all, hapq>=15, >=30, >=45
    coverage = np.zeros(n_snps)
    total_bases = 0
    with open(os.path.join(out_dir, f"{contig}.vartigs"), "w") as vf, \
            open(os.path.join(out_dir, "vartig_info.txt"), "w") as info:
        for i, ids in enumerate(parts):
            if len(ids) == 0:
                continue
            lo, hi = int(ranges[i][0]), int(ranges[i][1])
            left_gn, right_gn = int(snp_to_genome_pos[lo - 1]), int(snp_to_genome_pos[hi - 1])
            total_bases += right_gn - left_gn
            cov, err, _, _ = get_errors_cov_from_frags(frags, ids, lo, hi)
            q = int(hapqs[i])
            covered[0, lo - 1:hi] += 1.0
            coverage[lo - 1:hi] += cov
            for t, thr in ((1, 15), (2, 30), (3, 45)):
                if q >= thr:
                    covered[t, lo - 1:hi] += 1.0
            vf.write(f">HAP{i}.{out_dir}\tCONTIG:{contig}\tSNPRANGE:{lo}-{hi}\tBASERANGE:{left_gn + 1}-{right_gn + 1}\t"
                     f"COV:{cov:.3f}\tERR:{err:.4f}\tHAPQ:{q}\tREL_ERR:{rel_err[i]:.3f}\n")
            best, cnt = consensus_alleles(frags, ids, lo, hi)
            info.write(f">HAP{i}.{out_dir}\tSNPRANGE:{lo}-{hi}\n")
            for k in range(hi - lo + 1):  # (a haploset always holds cells: hap_map is never empty, file_writer.rs:320-323)
                pos = lo + k
                info.write(f"{pos}:{int(snp_to_genome_pos[pos - 1])}\t")
                row = cnt[k]
                if not row.any():
                    info.write("?\tNA\t\n")
                else:
                    info.write(f"{int(best[k])}\t" + "|".join(f"{a}:{int(c)}" for a, c in enumerate(row) if c) + "\t\n")
            vf.write("".join(chr(int(x) + 48) for x in best) + "\n")
    nonzero = int((covered[0] > 0).sum())
    row = (f"{contig}\t{covered[0].sum() / n_snps:.3f}\t{total_bases / contig_len:.3f}\t"
           f"{(coverage.sum() / nonzero) if nonzero else float('nan'):.3f}\t{total_bases}\t{covered[1].sum() / n_snps:.3f}\t"
           f"{covered[2].sum() / n_snps:.3f}\t{covered[3].sum() / n_snps:.3f}\t{avg_err:.4f}\n")
    path = os.path.join(top_dir, "contig_ploidy_info.tsv")
    new = not os.path.exists(path)
    with open(path, "a") as fh:
        if new:
            fh.write(CONTIG_PLOIDY_HEADER)
        fh.write(row)
    return row
