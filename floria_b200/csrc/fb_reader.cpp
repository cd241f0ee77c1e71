// fb_reader.cpp — BAM + VCF -> fragments (include/floria_b200_reader.h).  Host-only C++17 + zlib.
// Every step cites the reference lines it restates (paths relative to /root/reference); nothing here touches the device.
#include <zlib.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/floria_b200_reader.h"

namespace {

thread_local std::string g_reader_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_reader_err = buf;
    return code;
}

// gzip / BGZF (a series of gzip members) / plain file -> bytes
bool slurp_gz(const char *path, std::vector<uint8_t> &out) {
    gzFile f = gzopen(path, "rb");
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    out.clear();
    std::vector<uint8_t> buf(1 << 22);
    for (;;) {
        const int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n < 0) {
            gzclose(f);
            return false;
        }
        if (n == 0) break;
        out.insert(out.end(), buf.begin(), buf.begin() + n);
    }
    gzclose(f);
    return true;
}

// ---- get_vcf_profile (file_reader.rs:239-314) for the contigs of the BAM header -------------------------------------------
struct ContigSnps {
    std::unordered_map<int64_t, uint32_t> pos_to_snp;          // vcf_pos_to_snp_counter_map[contig]
    std::unordered_map<int64_t, std::vector<uint8_t>> alleles;  // vcf_pos_allele_map[contig]
    std::map<uint32_t, int64_t> snp_to_pos;                     // vcf_snp_pos_to_gn_pos_map[contig]
};

bool read_vcf(const char *path, const std::unordered_map<std::string, int> &ref_index, std::vector<ContigSnps> &prof) {
    std::vector<uint8_t> data;
    if (!slurp_gz(path, data)) return false;
    uint32_t snp_counter = 1;
    int last_ref = -1;
    size_t o = 0;
    while (o < data.size()) {
        size_t e = o;
        while (e < data.size() && data[e] != '\n') ++e;
        std::string ln(reinterpret_cast<const char *>(data.data()) + o, e - o);
        o = e + 1;
        if (!ln.empty() && ln.back() == '\r') ln.pop_back();
        if (ln.empty() || ln[0] == '#') continue;
        // CHROM POS ID REF ALT ...
        std::vector<std::string> f;
        size_t s = 0;
        while (f.size() < 5) {
            const size_t t = ln.find('\t', s);
            f.push_back(ln.substr(s, t == std::string::npos ? std::string::npos : t - s));
            if (t == std::string::npos) break;
            s = t + 1;
        }
        if (f.size() < 5) continue;
        auto it = ref_index.find(f[0]);
        if (it == ref_index.end()) continue;  // :267-270: contigs the BAM does not know are skipped
        if (it->second != last_ref) {         // :273-276: the counter restarts whenever the contig changes
            snp_counter = 1;
            last_ref = it->second;
        }
        // alleles = REF, ALT1, ALT2, ...; a SNP iff every allele is ONE base out of ACGT (case-insensitive) (:287-301)
        std::vector<std::string> al{f[3]};
        for (size_t a = 0;;) {
            const size_t t = f[4].find(',', a);
            al.push_back(f[4].substr(a, t == std::string::npos ? std::string::npos : t - a));
            if (t == std::string::npos) break;
            a = t + 1;
        }
        bool is_snp = true;
        std::vector<uint8_t> al_vec;
        for (const std::string &x : al) {
            if (x.size() != 1) {  // htslib hands out no empty alleles; a longer one is an indel / MNP
                is_snp = false;
                break;
            }
            const char u = (char)toupper((unsigned char)x[0]);
            if (u != 'A' && u != 'C' && u != 'G' && u != 'T') {
                is_snp = false;
                break;
            }
            al_vec.push_back((uint8_t)x[0]);
        }
        if (!is_snp) continue;
        const int64_t pos0 = strtoll(f[1].c_str(), nullptr, 10) - 1;  // rust-htslib Record::pos() is 0-based
        ContigSnps &c = prof[it->second];
        c.snp_to_pos[snp_counter] = pos0;  // :303-306 (plain inserts: a repeated position keeps the later counter)
        c.pos_to_snp[pos0] = snp_counter;
        c.alleles[pos0] = al_vec;
        ++snp_counter;
    }
    return true;
}

// ---- one alignment's fragment (frag_from_record, file_reader.rs:661-736) -------------------------------------------------
struct RFrag {
    std::map<uint32_t, std::pair<uint8_t, uint8_t>> cells;  // snp position -> (allele index, base quality): seq_dict + qual_dict
    uint32_t first = 0xFFFFFFFFu, last = 0;                   // build_frag defaults: MAX / MIN
    uint16_t flags = 0;
    bool paired = false;
    uint64_t counter = 0;  // index of the record among the contig's records (frag_from_record's `count`)
};

const char SEQ_CODE[] = "=ACMGRSVTWYHKDBN";

// alignment_passed_check (file_reader.rs:185-237) with filter_supplementary = true (:352)
bool passed_check(uint16_t flags, uint8_t mapq, bool use_supp, uint32_t mapq_cutoff) {
    const bool is_paired = (flags & 64) || (flags & 128);
    if (flags & 2048) {
        if (is_paired) return false;  // no supplementary alignments for short reads
        if (!use_supp) return false;
        if (mapq < 60) return false;  // mapq_supp_cutoff
    }
    if (mapq < mapq_cutoff) return false;
    if (flags & 1796) return false;  // erroneous alignment
    if (flags & 256) return false;   // secondary
    return true;
}

// b's cells into a (HashMap::extend: a key present in both takes b's value), file_reader.rs:539-546 / :631-640
void extend(RFrag &a, const RFrag &b) {
    for (const auto &kv : b.cells) a.cells[kv.first] = kv.second;
    a.first = std::min(a.first, b.first);
    a.last = std::max(a.last, b.last);
}

// Frag::cmp (types_structs.rs:87-93): (first asc, last desc, counter asc)
bool frag_less(const RFrag &x, const RFrag &y) {
    if (x.first != y.first) return x.first < y.first;
    if (x.last != y.last) return x.last > y.last;
    return x.counter < y.counter;
}

template <class T>
T rd(const uint8_t *p) {
    T v;
    memcpy(&v, p, sizeof(T));
    return v;
}

}  // namespace

extern "C" {

void fb_reader_options_default(fb_reader_options *o) {
    o->mapq_cutoff = 15;
    o->use_supp_aln = 1;
    o->supp_aln_dist_cutoff = 40000;
}

const char *fb_reader_last_error(void) { return g_reader_err.c_str(); }

void fb_free_frag_set(fb_frag_set *s) {
    if (!s) return;
    free((void *)s->frags.row_ptr);
    free((void *)s->frags.first);
    free((void *)s->frags.last);
    free((void *)s->frags.pos);
    free((void *)s->frags.allele);
    free((void *)s->frags.qual);
    free(s->snp_to_genome_pos);
    free(s);
}

int fb_read_frags(const char *bam_path, const char *vcf_path, const char *contig, const fb_reader_options *opt_in,
                  fb_frag_set **out) {
    if (!bam_path || !vcf_path || !out) return fail(FB_ERR_ARG, "fb_read_frags: null argument");
    fb_reader_options opt;
    if (opt_in)
        opt = *opt_in;
    else
        fb_reader_options_default(&opt);
    *out = nullptr;

    // ---- BAM header ---------------------------------------------------------------------------------------------------
    std::vector<uint8_t> bam;
    if (!slurp_gz(bam_path, bam)) return fail(FB_ERR_ARG, "cannot read BAM file %s", bam_path);
    if (bam.size() < 12 || memcmp(bam.data(), "BAM\1", 4) != 0) return fail(FB_ERR_ARG, "%s is not a BAM file", bam_path);
    size_t o = 8 + (size_t)rd<int32_t>(bam.data() + 4);
    if (o + 4 > bam.size()) return fail(FB_ERR_ARG, "truncated BAM header");
    const int32_t n_ref = rd<int32_t>(bam.data() + o);
    o += 4;
    std::vector<std::string> refs;
    std::unordered_map<std::string, int> ref_index;
    for (int32_t i = 0; i < n_ref; ++i) {
        if (o + 4 > bam.size()) return fail(FB_ERR_ARG, "truncated BAM header");
        const int32_t l_name = rd<int32_t>(bam.data() + o);
        if (l_name < 1 || o + 4 + (size_t)l_name + 4 > bam.size()) return fail(FB_ERR_ARG, "truncated BAM header");
        refs.emplace_back(reinterpret_cast<const char *>(bam.data()) + o + 4, (size_t)l_name - 1);
        ref_index.emplace(refs.back(), i);
        o += 4 + (size_t)l_name + 4;
    }
    const size_t rec0 = o;

    std::vector<ContigSnps> prof(refs.size());
    if (!read_vcf(vcf_path, ref_index, prof)) return fail(FB_ERR_ARG, "cannot read VCF file %s", vcf_path);

    // ---- which contig -------------------------------------------------------------------------------------------------
    int tid = -1;
    if (contig && contig[0]) {
        auto it = ref_index.find(contig);
        if (it == ref_index.end()) return fail(FB_ERR_ARG, "contig %s is not in the BAM header", contig);
        tid = it->second;
    } else {
        for (size_t p = rec0; p + 36 <= bam.size();) {  // first reference with SNPs that a record maps to
            const int32_t bs = rd<int32_t>(bam.data() + p);
            const int32_t r = rd<int32_t>(bam.data() + p + 4);
            if (r >= 0 && r < n_ref && !prof[r].pos_to_snp.empty()) {
                tid = r;
                break;
            }
            p += 4 + (size_t)bs;
        }
        if (tid < 0) return fail(FB_ERR_ARG, "no BAM record maps to a contig with SNPs in the VCF");
    }
    const ContigSnps &cs = prof[tid];

    // ---- records -> fragments, bucketed by read name in order of first appearance (file_reader.rs:389-441) ---------------------
    std::vector<std::vector<RFrag>> buckets;
    std::unordered_map<std::string, size_t> bucket_of;
    std::vector<uint32_t> lens;
    uint64_t n_records = 0, n_passed = 0;
    for (size_t p = rec0; p + 36 <= bam.size();) {
        const int32_t bs = rd<int32_t>(bam.data() + p);
        if (bs < 32 || p + 4 + (size_t)bs > bam.size()) return fail(FB_ERR_ARG, "truncated BAM record at byte %zu", p);
        const uint8_t *r = bam.data() + p + 4;
        p += 4 + (size_t)bs;
        const int32_t ref_id = rd<int32_t>(r);
        if (ref_id != tid) continue;
        const uint64_t count = n_records++;
        const int32_t pos = rd<int32_t>(r + 4);
        const uint8_t l_rn = r[8], mapq = r[9];
        const uint16_t n_cig = rd<uint16_t>(r + 12), flags = rd<uint16_t>(r + 14);
        const uint32_t l_seq = rd<uint32_t>(r + 16);
        const size_t need = 32 + (size_t)l_rn + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq;
        if (need > (size_t)bs) return fail(FB_ERR_ARG, "malformed BAM record %llu", (unsigned long long)count);
        if (!passed_check(flags, mapq, opt.use_supp_aln != 0, opt.mapq_cutoff)) continue;
        ++n_passed;
        lens.push_back(l_seq);
        const char *name = reinterpret_cast<const char *>(r + 32);
        const uint8_t *cig = r + 32 + l_rn;
        const uint8_t *sq = cig + 4 * (size_t)n_cig;
        const uint8_t *ql = sq + (l_seq + 1) / 2;
        RFrag fr;
        fr.flags = flags;
        fr.paired = (flags & 64) || (flags & 128);
        fr.counter = count;
        // aligned_pairs_full: only M / = / X give a pair with both coordinates; I and S advance the read, D and N the reference
        int64_t g = pos;
        uint32_t q = 0;
        for (uint16_t c = 0; c < n_cig; ++c) {
            const uint32_t v = rd<uint32_t>(cig + 4 * (size_t)c);
            const uint32_t op = v & 0xF, ln = v >> 4;
            if (op == 0 || op == 7 || op == 8) {
                for (uint32_t k = 0; k < ln; ++k) {
                    auto it = cs.pos_to_snp.find(g + k);
                    if (it == cs.pos_to_snp.end()) continue;
                    const uint32_t qi = q + k;
                    if (qi >= l_seq) break;
                    const uint8_t b = sq[qi >> 1];
                    const uint8_t base = (uint8_t)SEQ_CODE[(qi & 1) == 0 ? (b >> 4) : (b & 0xF)];
                    const std::vector<uint8_t> &al = cs.alleles.at(g + k);
                    for (size_t i = 0; i < al.size(); ++i)
                        if (base == al[i]) {  // :707-724: the first allele equal to the read base (case-sensitive, as there)
                            fr.cells[it->second] = std::make_pair((uint8_t)i, ql[qi]);
                            fr.first = std::min(fr.first, it->second);
                            fr.last = std::max(fr.last, it->second);
                            break;
                        }
                }
                q += ln;
                g += ln;
            } else if (op == 1 || op == 4) {
                q += ln;
            } else if (op == 2 || op == 3) {
                g += ln;
            }
        }
        const std::string nm(name, strnlen(name, l_rn));
        auto ins = bucket_of.emplace(nm, buckets.size());
        if (ins.second) buckets.emplace_back();
        buckets[ins.first->second].push_back(std::move(fr));
    }

    // ---- combine_frags (file_reader.rs:491-659) ---------------------------------------------------------------------------------
    std::vector<RFrag> frags;
    for (std::vector<RFrag> &fs : buckets) {
        if (fs.size() == 2 && fs[0].paired && fs[1].paired) {
            // :511 frags.sort() on (flags, Frag): the mate with the smaller flag word first
            if (fs[1].flags < fs[0].flags || (fs[1].flags == fs[0].flags && frag_less(fs[1], fs[0]))) std::swap(fs[0], fs[1]);
            RFrag *first = nullptr, *second = nullptr;
            if (fs[0].flags & 64) {
                first = &fs[0];
                second = &fs[1];
            } else if (fs[0].flags & 128) {
                first = &fs[1];
                second = &fs[0];
            } else {
                continue;  // :534-537
            }
            extend(*first, *second);
            frags.push_back(std::move(*first));
        } else if (fs.size() == 1 && !(fs[0].flags & 2048)) {
            frags.push_back(std::move(fs[0]));
        } else {
            // a long read with supplementary alignments (:565-656)
            std::vector<std::pair<uint32_t, uint32_t>> iv;
            for (const RFrag &f : fs)
                if (!f.cells.empty()) iv.emplace_back(f.first, f.last);
            std::sort(iv.begin(), iv.end());
            bool primary_only = false;
            for (size_t i = 0; i + 1 < iv.size(); ++i)
                if (cs.snp_to_pos.at(iv[i + 1].first) - cs.snp_to_pos.at(iv[i].second) > opt.supp_aln_dist_cutoff) {
                    primary_only = true;
                    break;
                }
            int primary = -1;
            for (size_t i = 0; i < fs.size(); ++i)
                if (!(fs[i].flags & 2048)) primary = (int)i;  // :609-617: the last one wins
            if (primary < 0) continue;                         // only supplementary alignments survived the filter
            if (!primary_only)
                for (size_t i = 0; i < fs.size(); ++i)
                    if ((int)i != primary) extend(fs[primary], fs[i]);
            frags.push_back(std::move(fs[primary]));
        }
    }
    uint64_t n_without = 0;
    {
        std::vector<RFrag> keep;
        for (RFrag &f : frags) {
            if (f.cells.empty())
                ++n_without;
            else
                keep.push_back(std::move(f));
        }
        frags.swap(keep);
    }
    std::sort(frags.begin(), frags.end(), frag_less);  // floria.rs:289-293

    // ---- output ---------------------------------------------------------------------------------------------------------------------
    fb_frag_set *s = (fb_frag_set *)calloc(1, sizeof(fb_frag_set));
    uint64_t nnz = 0;
    for (const RFrag &f : frags) nnz += f.cells.size();
    const uint64_t R = frags.size();
    uint64_t *row_ptr = (uint64_t *)calloc(R + 1, sizeof(uint64_t));
    uint32_t *first = (uint32_t *)calloc(R + 1, sizeof(uint32_t)), *last = (uint32_t *)calloc(R + 1, sizeof(uint32_t));
    uint32_t *pp = (uint32_t *)calloc(nnz + 1, sizeof(uint32_t));
    uint8_t *aa = (uint8_t *)calloc(nnz + 1, 1), *qq = (uint8_t *)calloc(nnz + 1, 1);
    uint64_t x = 0;
    for (uint64_t i = 0; i < R; ++i) {
        row_ptr[i] = x;
        first[i] = frags[i].first;
        last[i] = frags[i].last;
        for (const auto &kv : frags[i].cells) {
            pp[x] = kv.first;
            aa[x] = kv.second.first;
            qq[x] = kv.second.second;
            ++x;
        }
    }
    row_ptr[R] = x;
    s->frags.n_reads = R;
    s->frags.nnz = nnz;
    s->frags.row_ptr = row_ptr;
    s->frags.first = first;
    s->frags.last = last;
    s->frags.pos = pp;
    s->frags.allele = aa;
    s->frags.qual = qq;
    // SNP i + 1 -> genome position; counters are dense 1..n unless the VCF repeats a position
    s->n_snps = cs.snp_to_pos.empty() ? 0 : cs.snp_to_pos.rbegin()->first;
    s->snp_to_genome_pos = (uint64_t *)calloc(s->n_snps + 1, sizeof(uint64_t));
    for (const auto &kv : cs.snp_to_pos) s->snp_to_genome_pos[kv.first - 1] = (uint64_t)kv.second;
    s->n_records = n_records;
    s->n_passed = n_passed;
    s->n_without_snps = n_without;
    std::sort(lens.begin(), lens.end());
    s->read_len_p66 = lens.empty() ? 0u : lens[(size_t)((double)lens.size() * 0.66)];
    snprintf(s->contig, sizeof(s->contig), "%s", refs[tid].c_str());
    *out = s;
    return FB_OK;
}

}  // extern "C"
