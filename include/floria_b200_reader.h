/* floria_b200_reader.h — SURVEY.md §8 row f2: BAM + VCF -> the fragment arrays that cross floria_b200.h.
 *
 * Host-only C++ (zlib for BGZF / gzip), no device work.  Replaces, for one contig, what
 *   src/file_reader.rs:239-314  get_vcf_profile            (single-base alleles, 1-based SNP counter per contig)
 *   src/file_reader.rs:185-237  alignment_passed_check     (MAPQ, flag masks, supplementary rules)
 *   src/file_reader.rs:661-736  frag_from_record           (aligned pairs at SNP positions -> allele index + base quality)
 *   src/file_reader.rs:491-659  combine_frags              (mates of a pair, supplementary alignments of a long read)
 *   src/bin/floria.rs:289-293   sort by Frag::cmp (types_structs.rs:87-93), counter_id = index
 * compute.  NOT restated: alignment::realign (src/alignment.rs:7-64, the third-party block-aligner crate, active with -r),
 * the hybrid short+long path (one BAM here), BAM index queries (the whole file is scanned; records are filtered by contig).
 * Where the reference's result depends on hash-map or thread order (which primary alignment wins when a read has several,
 * the order of equal fragments before the sort) the BAM record order is used. */
#ifndef FLORIA_B200_READER_H
#define FLORIA_B200_READER_H
#include "floria_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t mapq_cutoff;         /* -m, default 15 (parse_cmd_line.rs:149) */
    uint32_t use_supp_aln;        /* !--dont-use-supp-aln (file_reader.rs:353), default 1 */
    int64_t supp_aln_dist_cutoff; /* default 40000 (parse_cmd_line.rs:35) */
} fb_reader_options;

/* library-owned result of fb_read_frags */
typedef struct {
    fb_frags frags;              /* reads in Frag::cmp order, counter_id == index; positions ascending within a read */
    uint64_t n_snps;
    uint64_t *snp_to_genome_pos; /* [n_snps] 0-based genome position of SNP i + 1 (vcf_snp_pos_to_gn_pos_map) */
    uint64_t n_records;          /* BAM records of the contig */
    uint64_t n_passed;           /* ... that passed alignment_passed_check */
    uint64_t n_without_snps;     /* fragments that cover no SNP (file_reader.rs:449-458: kept apart, not returned) */
    uint32_t read_len_p66;       /* 66th percentile of the passed records' lengths (block-length heuristic input) */
    uint32_t _pad;
    char contig[256];
} fb_frag_set;

void fb_reader_options_default(fb_reader_options *);
/* contig == NULL or "": the first BAM reference that has SNPs in the VCF and at least one record.  Returns FB_OK or an
 * FB_ERR_* code; fb_reader_last_error() (thread-local) holds the message. */
int fb_read_frags(const char *bam_path, const char *vcf_path, const char *contig, const fb_reader_options *,
                  fb_frag_set **out);
void fb_free_frag_set(fb_frag_set *);
const char *fb_reader_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
