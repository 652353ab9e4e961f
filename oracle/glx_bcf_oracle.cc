// glx_bcf_oracle.cc — CPU ORACLE for the BCF2 record encoder. TEST INFRASTRUCTURE ONLY.
//
// Restates what the reference does after the per-dataset loop of genotype_site to turn a site's calls into a BCF
// record (src/genotyper.cc:869-1003) together with the htslib 1.9 serialisation it relies on. htslib is a third-party
// dependency that is NOT vendored under /root/reference (CMakeLists.txt:44-60 downloads htslib 1.9), so its part is
// restated from the published BCF2 specification (VCFv4.2 / BCFv2.2 §6) and htslib 1.9's documented encoder rules:
//   bcf_enc_size / bcf_enc_int1 / bcf_enc_vint / bcf_enc_vfloat / bcf_enc_vchar   (htslib/vcf.h, vcf.c)
//     - a typed value is a descriptor byte (count << 4 | type) followed by the data; counts >= 15 put 15 in the byte
//       and append the count as a typed scalar integer;
//     - an integer vector is stored in the smallest of int8/int16/int32 whose range [-120,127] / [-32760,32767] holds
//       every value that is neither missing nor end-of-vector; sentinels become the type's own (0x80/0x81, ...);
//   bcf_update_genotypes / bcf_update_format_int32|float|string: FORMAT = typed key, typed per-sample vector, GT first;
//   bcf_update_format_string pads every sample's string with NULs to the longest;
//   bcf_trim_alleles -> bcf_remove_allele_set: ALT alleles no GT uses are dropped, GT renumbered, and every INFO /
//     FORMAT field whose header Number is A, R or G loses the dropped alleles' entries and is re-encoded (so the
//     integer width is chosen AFTER trimming); entries after a sample's first end-of-vector stay end-of-vector;
//   bcf1_sync (via bcf_dup, genotyper.cc:995-1003): shared = CHROM POS rlen QUAL n_info|n_allele n_fmt|n_sample, ID,
//     alleles, FILTER, INFO; indiv = the FORMAT blocks in order.
// Parity status: PINNED through the reference's 48 truth pVCFs — tests/test_bcf_oracle.py decodes these bytes with an
// independent BCF2 reader (tests/bcfdec.py) and compares every row with the truth files and, string for string, with the
// text assembler those files pin. Byte-level equality with htslib's own output is UNPINNED (no htslib here); what the
// spec leaves open is listed in DESIGN.md.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/glx.h"

namespace {

enum { BT_NULL = 0, BT_INT8 = 1, BT_INT16 = 2, BT_INT32 = 3, BT_FLOAT = 5, BT_CHAR = 7 };

void put8(std::string& s, uint32_t v) { s.push_back((char)(v & 0xff)); }
void put16(std::string& s, uint32_t v) {
    put8(s, v);
    put8(s, v >> 8);
}
void put32(std::string& s, uint32_t v) {
    put16(s, v);
    put16(s, v >> 16);
}

// bcf_enc_size
void enc_size(std::string& s, int size, int type) {
    if (size >= 15) {
        put8(s, 15u << 4 | type);
        if (size >= 128) {
            if (size >= 32768) {
                put8(s, 1u << 4 | BT_INT32);
                put32(s, (uint32_t)size);
            } else {
                put8(s, 1u << 4 | BT_INT16);
                put16(s, (uint32_t)size);
            }
        } else {
            put8(s, 1u << 4 | BT_INT8);
            put8(s, (uint32_t)size);
        }
    } else
        put8(s, (uint32_t)size << 4 | type);
}

// bcf_enc_int1
void enc_int1(std::string& s, int32_t x) {
    if (x == GLX_INT_VECTOR_END) {
        enc_size(s, 1, BT_INT8);
        put8(s, 0x81);
    } else if (x == GLX_INT_MISSING) {
        enc_size(s, 1, BT_INT8);
        put8(s, 0x80);
    } else if (x <= 127 && x >= -120) {
        enc_size(s, 1, BT_INT8);
        put8(s, (uint32_t)x);
    } else if (x <= 32767 && x >= -32760) {
        enc_size(s, 1, BT_INT16);
        put16(s, (uint32_t)x);
    } else {
        enc_size(s, 1, BT_INT32);
        put32(s, (uint32_t)x);
    }
}

// bcf_enc_vint(s, n, a, wsize)
void enc_vint(std::string& s, const int32_t* a, int n, int wsize) {
    if (n <= 0) {
        enc_size(s, 0, BT_NULL);
        return;
    }
    if (n == 1) {
        enc_int1(s, a[0]);
        return;
    }
    if (wsize <= 0) wsize = n;
    int32_t mx = INT32_MIN + 1, mn = INT32_MAX;
    for (int i = 0; i < n; i++) {
        if (a[i] == GLX_INT_MISSING || a[i] == GLX_INT_VECTOR_END) continue;
        if (mx < a[i]) mx = a[i];
        if (mn > a[i]) mn = a[i];
    }
    if (mx <= 127 && mn >= -120) {
        enc_size(s, wsize, BT_INT8);
        for (int i = 0; i < n; i++) put8(s, a[i] == GLX_INT_VECTOR_END ? 0x81u : a[i] == GLX_INT_MISSING ? 0x80u : (uint32_t)a[i]);
    } else if (mx <= 32767 && mn >= -32760) {
        enc_size(s, wsize, BT_INT16);
        for (int i = 0; i < n; i++) put16(s, a[i] == GLX_INT_VECTOR_END ? 0x8001u : a[i] == GLX_INT_MISSING ? 0x8000u : (uint32_t)a[i]);
    } else {
        enc_size(s, wsize, BT_INT32);
        for (int i = 0; i < n; i++) put32(s, (uint32_t)a[i]);
    }
}

void enc_vchar(std::string& s, const char* p, size_t n) {
    enc_size(s, (int)n, BT_CHAR);
    s.append(p, n);
}

std::string pool_str(const char* pool, const uint32_t* off, size_t i) { return std::string(pool + off[i], pool + off[i + 1]); }

}  // namespace

extern "C" {

// The BCF2 records of one batch, uncompressed, in site order: bytes [rec_off[s], rec_off[s+1]) are site s's record
// (empty when the site is dropped after trimming, genotyper.cc:963-965). *bytes is malloc'd; rec_off holds S+1 entries.
int glx_oracle_bcf_records(const glx_config* cfg, const glx_sites* sites, const glx_bcf_meta* meta, const glx_out* out, uint8_t** bytes, uint64_t* n_bytes,
                           uint64_t* rec_off) {
    if (!cfg || !sites || !meta || !out || !bytes || !n_bytes || !rec_off) return GLX_E_INVALID;
    const size_t S = sites->n_sites, N = sites->n_samples;
    std::string o;
    rec_off[0] = 0;
    for (size_t si = 0; si < S; si++) {
        const int A = sites->n_allele[si];
        const uint32_t a0 = sites->allele_off[si];
        // bcf_trim_alleles: ALT alleles with no GT entry go (REF always stays)
        std::vector<int> keep{0};
        for (int a = 1; a < A; a++)
            if (!cfg->trim_uncalled_alleles || (out->allele_used[si] >> a & 1)) keep.push_back(a);
        if ((int)keep.size() < 2) {  // ans.reset()
            rec_off[si + 1] = o.size();
            continue;
        }
        std::vector<int> renum(A, -1);
        for (size_t i = 0; i < keep.size(); i++) renum[keep[i]] = (int)i;
        const bool trimmed = (int)keep.size() != A;
        const std::string ref = pool_str(meta->dna_pool, meta->dna_off, a0);

        // ---- shared ----
        std::string sh;
        put32(sh, (uint32_t)meta->rid[si]);
        put32(sh, (uint32_t)sites->beg[si]);
        put32(sh, (uint32_t)ref.size());  // rlen: bcf_update_alleles sets it from REF
        uint32_t qb;
        const float q = meta->qual[si];
        memcpy(&qb, &q, 4);
        put32(sh, qb);
        bool output_af = true, any_aq = false;  // genotyper.cc:884-912, over ALL unified alleles
        for (int a = 1; a < A; a++) {
            const float f = sites->allele_af[a0 + a];
            if (f != f) output_af = false;
            if (meta->aq[a0 + a]) any_aq = true;
        }
        const uint32_t n_info = (output_af ? 1u : 0u) + (any_aq ? 1u : 0u);
        const uint32_t n_fmt = 2 + cfg->n_fields;
        put32(sh, (uint32_t)keep.size() << 16 | n_info);
        put32(sh, n_fmt << 24 | (uint32_t)N);
        // ID (genotyper.cc:968-993)
        std::string id;
        const std::string chrom = pool_str(meta->contig_pool, meta->contig_off, (size_t)meta->rid[si]);
        for (size_t i = 1; i < keep.size(); i++) {
            const int a = keep[i];
            const int nb = meta->norm_beg[a0 + a], ne = meta->norm_end[a0 + a];
            if (i > 1) id += ";";
            id += chrom + "_" + std::to_string(nb + 1) + "_" + ref.substr((size_t)(nb - sites->beg[si]), (size_t)(ne - nb)) + "_" +
                  pool_str(meta->norm_pool, meta->norm_off, a0 + a);
        }
        enc_vchar(sh, id.data(), id.size());
        for (int a : keep) {
            const std::string d = pool_str(meta->dna_pool, meta->dna_off, a0 + a);
            enc_vchar(sh, d.data(), d.size());
        }
        if (sites->monoallelic[si]) {
            const int32_t f = meta->id_MONOALLELIC;
            enc_vint(sh, &f, 1, -1);
        } else
            enc_vint(sh, nullptr, 0, -1);
        if (output_af) {
            enc_int1(sh, meta->id_AF);
            enc_size(sh, (int)keep.size() - 1, BT_FLOAT);
            for (size_t i = 1; i < keep.size(); i++) {
                uint32_t b;
                memcpy(&b, &sites->allele_af[a0 + keep[i]], 4);
                put32(sh, b);
            }
        }
        if (any_aq) {
            enc_int1(sh, meta->id_AQ);
            std::vector<int32_t> v;
            for (size_t i = 1; i < keep.size(); i++) v.push_back(meta->aq[a0 + keep[i]]);
            enc_vint(sh, v.data(), (int)v.size(), -1);
        }

        // ---- indiv ----
        std::string iv;
        {  // GT (bcf_update_genotypes: htslib GT ints, renumbered by bcf_remove_allele_set)
            std::vector<int32_t> g(2 * N);
            for (size_t n = 0; n < 2 * N; n++) {
                const int a = out->gt[si * N * 2 + n];
                g[n] = a < 0 ? 0 : ((trimmed ? renum[a] : a) + 1) << 1;
            }
            enc_int1(iv, meta->id_GT);
            enc_vint(iv, g.data(), (int)g.size(), 2);
        }
        for (uint32_t k = 0; k < cfg->n_fields; k++) {
            const glx_field& f = cfg->fields[k];
            const unsigned cnt = sites->fld_cnt[k][si];
            const int32_t* base = out->fld[k] + sites->fld_off[k][si];
            enc_int1(iv, meta->id_field[k]);
            if (f.kind == GLX_FK_FT || f.kind == GLX_FK_STRING) {  // bcf_update_format_string
                std::vector<std::string> strs(N);
                size_t max_len = 0;
                for (size_t n = 0; n < N; n++) {
                    const uint32_t m = (uint32_t)base[n * cnt];
                    std::string t;
                    for (uint32_t i = 0; i < meta->n_filters; i++)
                        if (m >> i & 1) {
                            if (!t.empty()) t += ";";
                            t += pool_str(meta->filter_pool, meta->filter_off, i);
                        }
                    if (t.empty()) t = ".";
                    max_len = std::max(max_len, t.size());
                    strs[n] = t;
                }
                enc_size(iv, (int)max_len, BT_CHAR);
                for (size_t n = 0; n < N; n++) {
                    iv += strs[n];
                    iv.append(max_len - strs[n].size(), '\0');
                }
                continue;
            }
            const bool is_float = f.type == GLX_TYPE_FLOAT;
            const int32_t vend = is_float ? (int32_t)GLX_FLOAT_VECTOR_END_BITS : GLX_INT_VECTOR_END;
            const uint8_t vl = meta->field_vl[k];
            // source slots that survive bcf_remove_allele_set
            std::vector<unsigned> src;
            if (!trimmed || (vl != 'A' && vl != 'R' && vl != 'G')) {
                for (unsigned j = 0; j < cnt; j++) src.push_back(j);
            } else if (vl == 'G') {
                unsigned j = 0;
                for (int ib = 0; ib < A && j < cnt; ib++)
                    for (int ia = 0; ia <= ib && j < cnt; ia++, j++)
                        if (renum[ia] >= 0 && renum[ib] >= 0) src.push_back(j);
            } else {
                const int shift = vl == 'A' ? 1 : 0;
                for (unsigned j = 0; j < cnt; j++)
                    if ((int)j + shift < A && renum[j + shift] >= 0) src.push_back(j);
            }
            const size_t cn = src.size();
            const bool recoded = trimmed && (vl == 'A' || vl == 'R' || vl == 'G');  // bcf_remove_allele_set rewrites the vector
            std::vector<int32_t> v(N * cn);
            for (size_t n = 0; n < N; n++) {
                const int32_t* p = base + n * cnt;
                unsigned size = cnt;  // entries before the sample's first end-of-vector
                for (unsigned j = 0; recoded && j < cnt; j++)
                    if (p[j] == vend) {
                        size = j;
                        break;
                    }
                for (size_t i = 0; i < cn; i++) v[n * cn + i] = src[i] < size ? p[src[i]] : vend;
            }
            if (is_float) {
                enc_size(iv, (int)cn, BT_FLOAT);
                for (int32_t x : v) put32(iv, (uint32_t)x);
            } else
                enc_vint(iv, v.data(), (int)v.size(), (int)cn);
        }
        enc_int1(iv, meta->id_RNC);  // bcf_update_format_string of 2N one-character strings = 2 chars per sample
        enc_size(iv, 2, BT_CHAR);
        iv.append(reinterpret_cast<const char*>(out->rnc + si * N * 2), N * 2);

        put32(o, (uint32_t)sh.size());
        put32(o, (uint32_t)iv.size());
        o += sh;
        o += iv;
        rec_off[si + 1] = o.size();
    }
    *n_bytes = o.size();
    *bytes = (uint8_t*)malloc(o.size() ? o.size() : 1);
    if (!*bytes) return GLX_E_FAILURE;
    memcpy(*bytes, o.data(), o.size());
    return GLX_OK;
}

void glx_oracle_free(void* p) { free(p); }
}
