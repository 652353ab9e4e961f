// glx_bcf.cuh — BCF2 record encoder on the device: the packed FORMAT columns of a batch become the
// finished, byte-packed BCF records the reference's writer thread would emit (src/genotyper.cc:869-1003 + htslib's
// bcf_update_* / bcf_trim_alleles / bcf1_sync; rules restated in oracle/glx_bcf_oracle.cc, which pins them).
//
// Stages (kernels in glx_bcf.cu; every stage is a block-level function of (tid, nt) here so that the host build of this
// header — tests/emu — can run it with one virtual thread and diff it against the oracle):
//   scan    per (site, field): min / max of the values that survive allele trimming  -> integer width per SITE (bcf_enc_vint)
//   layout  per site: surviving alleles, widths, FORMAT block offsets, l_shared / l_indiv, record length
//   offsets exclusive scans: record offsets in the stream, offsets of the header+shared bytes in a side pool
//   headers per site: l_shared, l_indiv and the shared part (ID, alleles, FILTER, INFO AF/AQ) into the side pool
//   span    per 65,280-byte span of the output stream (= the payload of one BGZF block): gather the span in shared
//           memory — header bytes, FORMAT block headers, typed per-sample vectors converted from the int32 columns —
//           then hand it to the writer (raw: one bulk store; BGZF: deflate + CRC in glx_bgzf.cuh)
#pragma once
#include "glx_dev_types.cuh"

namespace glxd {

#define GLX_BCF_SPAN 65280u  // uncompressed bytes per span / BGZF block (htslib's BGZF_BLOCK_SIZE 0xff00)
#define GLX_BCF_MAXBLK (GLX_MAX_FIELDS + 2)
#define GLX_BCF_G_MAX (GLX_MAX_ALLELES * (GLX_MAX_ALLELES + 1) / 2)
#define GLX_BCF_HDR_MAX 12  // typed key (<= 5 bytes) + typed count descriptor (<= 6 bytes)

#if defined(__CUDA_ARCH__)
#define GLX_BLOCK_SYNC() __syncthreads()
#else
#define GLX_BLOCK_SYNC() ((void)0)
#endif

#define GLXB_MISSING ((int32_t)0x80000000)  // bcf_int32_missing
#define GLXB_VEND ((int32_t)0x80000001)     // bcf_int32_vector_end
enum { BCF_BT_NULL = 0, BCF_BT_INT8 = 1, BCF_BT_INT16 = 2, BCF_BT_INT32 = 3, BCF_BT_FLOAT = 5, BCF_BT_CHAR = 7 };

struct DBcfMeta {
    const int32_t* rid;
    const float* qual;
    const uint32_t* dna_off;
    const char* dna_pool;
    const int32_t *norm_beg, *norm_end;
    const uint32_t* norm_off;
    const char* norm_pool;
    const int32_t* aq;
    const float* af;
    const uint32_t* contig_off;
    const char* contig_pool;
    const uint32_t* filter_off;
    const char* filter_pool;
    uint32_t n_filters;
    int32_t id_AF, id_AQ, id_MONO, id_GT, id_RNC;
    int32_t id_field[GLX_MAX_FIELDS];
};

struct BcfSitePlan {
    uint32_t keep_mask;  // surviving unified alleles (bit 0 = REF); 0 = the site has no record
    uint32_t l_shared, l_indiv;
    uint8_t n_keep, trimmed, n_blk, _pad;
    uint32_t blk_off[GLX_BCF_MAXBLK];  // FORMAT blocks in record order (GT, lifted fields, RNC): offset of the block's key byte in indiv
    uint16_t blk_cnt[GLX_BCF_MAXBLK];  // values (chars) per sample after trimming
    uint8_t blk_width[GLX_BCF_MAXBLK];
    uint8_t blk_hdr_len[GLX_BCF_MAXBLK];
    uint8_t blk_hdr[GLX_BCF_MAXBLK][GLX_BCF_HDR_MAX];  // typed key + typed descriptor
};

struct BcfEnc {
    uint32_t n_sites, n_samples, n_fields;
    uint8_t trim, _pad[3];
    uint8_t fkind[GLX_MAX_FIELDS], ftype[GLX_MAX_FIELDS], fvl[GLX_MAX_FIELDS];
    const int32_t* beg;
    const uint8_t *n_allele, *monoallelic;
    const uint32_t* allele_off;
    const uint16_t* fld_cnt[GLX_MAX_FIELDS];
    const uint64_t* fld_off[GLX_MAX_FIELDS];
    const int8_t* gt;
    const uint8_t* rnc;
    const int32_t* fld[GLX_MAX_FIELDS];
    const uint32_t* allele_used;
    DBcfMeta m;
    int32_t* mm;  // [S][n_fields][2]: min, max over surviving non-sentinel values (string fields: [1] = longest string)
    BcfSitePlan* plan;
    unsigned long long* rec_off;  // [S+1] record lengths, then their exclusive scan
    unsigned long long* hdr_off;  // [S+1] likewise for 8 + l_shared
    uint8_t* hdr_pool;
};

// host side: the scalar part of BcfEnc from the ABI structs (pointers are bound by the caller)
inline void bcf_enc_scalars(BcfEnc& E, const glx_config& c, const glx_bcf_meta& m, uint32_t n_sites, uint32_t n_samples) {
    E.n_sites = n_sites;
    E.n_samples = n_samples;
    E.n_fields = c.n_fields;
    E.trim = c.trim_uncalled_alleles;
    for (uint32_t k = 0; k < c.n_fields; k++) {
        E.fkind[k] = c.fields[k].kind;
        E.ftype[k] = c.fields[k].type;
        E.fvl[k] = m.field_vl[k];
        E.m.id_field[k] = m.id_field[k];
    }
    E.m.n_filters = m.n_filters;
    E.m.id_AF = m.id_AF;
    E.m.id_AQ = m.id_AQ;
    E.m.id_MONO = m.id_MONOALLELIC;
    E.m.id_GT = m.id_GT;
    E.m.id_RNC = m.id_RNC;
}

// ---- small helpers ----
__host__ __device__ __forceinline__ uint32_t bcf_keep_mask(const BcfEnc& E, uint32_t s) {
    const uint32_t A = E.n_allele[s];
    const uint32_t all = A >= 32 ? 0xffffffffu : (1u << A) - 1u;
    return E.trim ? ((E.allele_used[s] | 1u) & all) : all;
}
__host__ __device__ __forceinline__ int bcf_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// genotype index j -> (ia <= ib), src/diploid.cc:28-50
__host__ __device__ __forceinline__ void bcf_gt_alleles(uint32_t j, uint32_t& ia, uint32_t& ib) {
    uint32_t b = (uint32_t)((sqrtf(8.0f * (float)j + 1.0f) - 1.0f) * 0.5f);
    while ((b + 1) * (b + 2) / 2 <= j) b++;
    while (b * (b + 1) / 2 > j) b--;
    ib = b;
    ia = j - b * (b + 1) / 2;
}
__host__ __device__ __forceinline__ bool bcf_is_arg(uint8_t vl) { return vl == 'A' || vl == 'R' || vl == 'G'; }
// does source slot j of a Number=vl field survive the trim to keep mask km?
__host__ __device__ __forceinline__ bool bcf_slot_kept(uint8_t vl, uint32_t j, uint32_t km) {
    if (vl == 'R') return j < 32 && (km >> j & 1u);
    if (vl == 'A') return j + 1 < 32 && (km >> (j + 1) & 1u);
    uint32_t ia, ib;
    bcf_gt_alleles(j, ia, ib);
    return ib < 32 && (km >> ia & 1u) && (km >> ib & 1u);
}
// its slot after the trim
__host__ __device__ __forceinline__ uint32_t bcf_slot_new(uint8_t vl, uint32_t j, uint32_t km) {
    if (vl == 'R') return (uint32_t)bcf_popc(km & ((1u << j) - 1u));
    if (vl == 'A') return (uint32_t)bcf_popc(km & ((1u << (j + 1)) - 1u)) - 1u;
    uint32_t ia, ib;
    bcf_gt_alleles(j, ia, ib);
    const uint32_t ra = (uint32_t)bcf_popc(km & ((1u << ia) - 1u)), rb = (uint32_t)bcf_popc(km & ((1u << ib) - 1u));
    return rb * (rb + 1) / 2 + ra;
}
__host__ __device__ __forceinline__ uint32_t bcf_count_new(uint8_t vl, uint32_t cnt, uint32_t n_keep, bool trimmed) {
    if (!trimmed || !bcf_is_arg(vl)) return cnt;
    return vl == 'R' ? n_keep : (vl == 'A' ? n_keep - 1 : n_keep * (n_keep + 1) / 2);
}
// length of the FT string of a filter mask ("." when empty; names joined by ';')
__host__ __device__ __forceinline__ uint32_t bcf_ft_len(const DBcfMeta& m, uint32_t mask) {
    if (!mask) return 1;
    uint32_t n = 0, k = 0;
    for (uint32_t i = 0; i < m.n_filters; i++)
        if (mask >> i & 1u) {
            n += m.filter_off[i + 1] - m.filter_off[i];
            k++;
        }
    return n ? n + k - 1 : 1;
}
__host__ __device__ __forceinline__ uint8_t bcf_ft_char(const DBcfMeta& m, uint32_t mask, uint32_t ci) {
    if (!mask) return ci == 0 ? (uint8_t)'.' : 0;
    uint32_t pos = 0;
    bool first = true;
    for (uint32_t i = 0; i < m.n_filters; i++)
        if (mask >> i & 1u) {
            if (!first) {
                if (ci == pos) return (uint8_t)';';
                pos++;
            }
            first = false;
            const uint32_t len = m.filter_off[i + 1] - m.filter_off[i];
            if (ci < pos + len) return (uint8_t)m.filter_pool[m.filter_off[i] + (ci - pos)];
            pos += len;
        }
    if (pos == 0) return ci == 0 ? (uint8_t)'.' : 0;  // mask bits beyond the dictionary
    return 0;
}

// ---- typed-value encoders over a byte sink (count or write) ----
struct BcfCount {
    uint32_t n = 0;
    __host__ __device__ void put(uint32_t) { n++; }
};
struct BcfWrite {
    uint8_t* p;
    __host__ __device__ void put(uint32_t b) { *p++ = (uint8_t)b; }
};
template <class Sink>
__host__ __device__ inline void bcf_put32(Sink& o, uint32_t v) {
    o.put(v & 0xff);
    o.put(v >> 8 & 0xff);
    o.put(v >> 16 & 0xff);
    o.put(v >> 24);
}
template <class Sink>
__host__ __device__ inline void bcf_enc_size(Sink& o, uint32_t size, int type) {
    if (size >= 15) {
        o.put(15u << 4 | (uint32_t)type);
        if (size >= 128) {
            if (size >= 32768) {
                o.put(1u << 4 | BCF_BT_INT32);
                bcf_put32(o, size);
            } else {
                o.put(1u << 4 | BCF_BT_INT16);
                o.put(size & 0xff);
                o.put(size >> 8);
            }
        } else {
            o.put(1u << 4 | BCF_BT_INT8);
            o.put(size);
        }
    } else
        o.put(size << 4 | (uint32_t)type);
}
__host__ __device__ __forceinline__ int bcf_int_width(int32_t mn, int32_t mx) {  // bcf_enc_vint's type choice
    if (mx <= 127 && mn >= -120) return 1;
    if (mx <= 32767 && mn >= -32760) return 2;
    return 4;
}
template <class Sink>
__host__ __device__ inline void bcf_enc_int1(Sink& o, int32_t x) {
    if (x == GLXB_VEND || x == GLXB_MISSING) {
        bcf_enc_size(o, 1, BCF_BT_INT8);
        o.put(x == GLXB_VEND ? 0x81u : 0x80u);
        return;
    }
    const int w = bcf_int_width(x, x);
    bcf_enc_size(o, 1, w == 1 ? BCF_BT_INT8 : (w == 2 ? BCF_BT_INT16 : BCF_BT_INT32));
    for (int i = 0; i < w; i++) o.put((uint32_t)x >> (8 * i) & 0xff);
}
template <class Sink>
__host__ __device__ inline void bcf_enc_pool_str(Sink& o, const char* pool, uint32_t lo, uint32_t hi) {
    bcf_enc_size(o, hi - lo, BCF_BT_CHAR);
    for (uint32_t i = lo; i < hi; i++) o.put((uint8_t)pool[i]);
}
template <class Sink>
__host__ __device__ inline void bcf_put_decimal(Sink& o, uint32_t v) {
    char d[10];
    int n = 0;
    do {
        d[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) o.put((uint8_t)d[--n]);
}
__host__ __device__ __forceinline__ uint32_t bcf_decimal_len(uint32_t v) {
    uint32_t n = 1;
    while (v >= 10) {
        v /= 10;
        n++;
    }
    return n;
}

// The shared part of site s (src/genotyper.cc:869-912,950-993 through bcf1_sync): fixed fields, ID, alleles, FILTER, INFO.
template <class Sink>
__host__ __device__ inline void bcf_emit_shared(const BcfEnc& E, uint32_t s, uint32_t km, uint32_t n_fmt, Sink& o) {
    const DBcfMeta& m = E.m;
    const uint32_t A = E.n_allele[s], a0 = E.allele_off[s];
    const uint32_t n_keep = (uint32_t)bcf_popc(km);
    const int32_t beg = E.beg[s];
    bool output_af = true, any_aq = false;
    for (uint32_t a = 1; a < A; a++) {
        const float f = m.af[a0 + a];
        if (f != f) output_af = false;
        if (m.aq[a0 + a]) any_aq = true;
    }
    bcf_put32(o, (uint32_t)m.rid[s]);
    bcf_put32(o, (uint32_t)beg);
    bcf_put32(o, m.dna_off[a0 + 1] - m.dna_off[a0]);  // rlen = strlen(REF)
    uint32_t qb;
    {
        const float q = m.qual[s];
        memcpy(&qb, &q, 4);
    }
    bcf_put32(o, qb);
    bcf_put32(o, n_keep << 16 | ((output_af ? 1u : 0u) + (any_aq ? 1u : 0u)));
    bcf_put32(o, n_fmt << 24 | E.n_samples);
    // ID: chrom_pos_ref_alt of each surviving ALT's normalized form, ';'-joined
    const uint32_t c_lo = m.contig_off[m.rid[s]], c_hi = m.contig_off[m.rid[s] + 1];
    uint32_t id_len = 0;
    for (uint32_t a = 1; a < A; a++)
        if (km >> a & 1u) {
            const uint32_t nb = (uint32_t)m.norm_beg[a0 + a], ne = (uint32_t)m.norm_end[a0 + a];
            id_len += (id_len ? 1u : 0u) + (c_hi - c_lo) + 1 + bcf_decimal_len(nb + 1) + 1 + (ne - nb) + 1 + (m.norm_off[a0 + a + 1] - m.norm_off[a0 + a]);
        }
    bcf_enc_size(o, id_len, BCF_BT_CHAR);
    bool first = true;
    for (uint32_t a = 1; a < A; a++)
        if (km >> a & 1u) {
            const uint32_t nb = (uint32_t)m.norm_beg[a0 + a], ne = (uint32_t)m.norm_end[a0 + a];
            if (!first) o.put(';');
            first = false;
            for (uint32_t i = c_lo; i < c_hi; i++) o.put((uint8_t)m.contig_pool[i]);
            o.put('_');
            bcf_put_decimal(o, nb + 1);
            o.put('_');
            const uint32_t r0 = m.dna_off[a0] + (nb - (uint32_t)beg);
            for (uint32_t i = 0; i < ne - nb; i++) o.put((uint8_t)m.dna_pool[r0 + i]);
            o.put('_');
            for (uint32_t i = m.norm_off[a0 + a]; i < m.norm_off[a0 + a + 1]; i++) o.put((uint8_t)m.norm_pool[i]);
        }
    for (uint32_t a = 0; a < A; a++)
        if (km >> a & 1u) bcf_enc_pool_str(o, m.dna_pool, m.dna_off[a0 + a], m.dna_off[a0 + a + 1]);
    if (E.monoallelic[s]) bcf_enc_int1(o, m.id_MONO);  // FILTER: bcf_enc_vint of one id
    else bcf_enc_size(o, 0, BCF_BT_NULL);
    if (output_af) {
        bcf_enc_int1(o, m.id_AF);
        bcf_enc_size(o, n_keep - 1, BCF_BT_FLOAT);
        for (uint32_t a = 1; a < A; a++)
            if (km >> a & 1u) {
                uint32_t b;
                const float f = m.af[a0 + a];
                memcpy(&b, &f, 4);
                bcf_put32(o, b);
            }
    }
    if (any_aq) {
        bcf_enc_int1(o, m.id_AQ);
        if (n_keep - 1 == 1) {
            for (uint32_t a = 1; a < A; a++)
                if (km >> a & 1u) bcf_enc_int1(o, m.aq[a0 + a]);
        } else {
            int32_t mn = INT32_MAX, mx = INT32_MIN + 1;
            for (uint32_t a = 1; a < A; a++)
                if (km >> a & 1u) {
                    const int32_t v = m.aq[a0 + a];
                    if (v == GLXB_MISSING || v == GLXB_VEND) continue;
                    mn = v < mn ? v : mn;
                    mx = v > mx ? v : mx;
                }
            const int w = bcf_int_width(mn, mx);
            bcf_enc_size(o, n_keep - 1, w == 1 ? BCF_BT_INT8 : (w == 2 ? BCF_BT_INT16 : BCF_BT_INT32));
            for (uint32_t a = 1; a < A; a++)
                if (km >> a & 1u) {
                    const int32_t v = m.aq[a0 + a];
                    uint32_t x = (uint32_t)v;
                    if (w < 4 && v == GLXB_MISSING) x = w == 1 ? 0x80u : 0x8000u;
                    if (w < 4 && v == GLXB_VEND) x = w == 1 ? 0x81u : 0x8001u;
                    for (int i = 0; i < w; i++) o.put(x >> (8 * i) & 0xff);
                }
        }
    }
}

struct BcfVec4 {
    int32_t v[4];
};
__host__ __device__ __forceinline__ BcfVec4 bcf_load16i(const int32_t* p) {
    BcfVec4 r;
#if defined(__CUDA_ARCH__)
    const int4 x = __ldg(reinterpret_cast<const int4*>(p));
    r.v[0] = x.x;
    r.v[1] = x.y;
    r.v[2] = x.z;
    r.v[3] = x.w;
#else
    memcpy(&r, p, 16);
#endif
    return r;
}

// ---- stage: scan. One block takes samples [n0, n1) of site s. s_mm = [n_fields][2] block-shared ints. ----
__host__ __device__ inline void bcf_scan_init(const BcfEnc& E, int32_t* s_mm, uint32_t tid, uint32_t nt) {
    for (uint32_t i = tid; i < E.n_fields; i += nt) {
        s_mm[2 * i] = INT32_MAX;
        s_mm[2 * i + 1] = INT32_MIN + 1;
    }
}
__host__ __device__ __forceinline__ void bcf_mm_merge(int32_t* dst, int32_t mn, int32_t mx) {
#if defined(__CUDA_ARCH__)
    if (mn != INT32_MAX) atomicMin(dst, mn);
    if (mx != INT32_MIN + 1) atomicMax(dst + 1, mx);
#else
    if (mn < dst[0]) dst[0] = mn;
    if (mx > dst[1]) dst[1] = mx;
#endif
}
__host__ __device__ inline void bcf_scan_accum(const BcfEnc& E, uint32_t s, uint32_t n0, uint32_t n1, int32_t* s_mm, uint32_t tid, uint32_t nt) {
    const uint32_t km = bcf_keep_mask(E, s);
    const bool trimmed = (uint32_t)bcf_popc(km) != E.n_allele[s];
    for (uint32_t k = 0; k < E.n_fields; k++) {
        const uint32_t cnt = E.fld_cnt[k][s];
        const int32_t* __restrict__ base = E.fld[k] + E.fld_off[k][s];
        int32_t mn = INT32_MAX, mx = INT32_MIN + 1;
        if (E.fkind[k] == GLX_FK_FT || E.fkind[k] == GLX_FK_STRING) {
            for (uint32_t n = n0 + tid; n < n1; n += nt) {
                const int32_t len = (int32_t)bcf_ft_len(E.m, (uint32_t)base[(size_t)n * cnt]);
                mx = len > mx ? len : mx;
            }
        } else if (E.ftype[k] == GLX_TYPE_FLOAT) {
            continue;
        } else if (trimmed && bcf_is_arg(E.fvl[k])) {
            // bcf_remove_allele_set rewrites the vector: only surviving slots before the sample's first end-of-vector count
            const uint8_t vl = E.fvl[k];
            for (uint32_t n = n0 + tid; n < n1; n += nt) {
                const int32_t* p = base + (size_t)n * cnt;
                for (uint32_t j = 0; j < cnt; j++) {
                    const int32_t v = p[j];
                    if (v == GLXB_VEND) break;
                    if (v == GLXB_MISSING || !bcf_slot_kept(vl, j, km)) continue;
                    mn = v < mn ? v : mn;
                    mx = v > mx ? v : mx;
                }
            }
        } else {
            const size_t e0 = (size_t)n0 * cnt, e1 = (size_t)n1 * cnt;
            auto fold = [&](int32_t v) {
                if (v == GLXB_MISSING || v == GLXB_VEND) return;
                mn = v < mn ? v : mn;
                mx = v > mx ? v : mx;
            };
            size_t a0 = e0 + ((4 - ((reinterpret_cast<uintptr_t>(base + e0) >> 2) & 3)) & 3);  // first 16-byte aligned element
            if (a0 > e1) a0 = e1;
            for (size_t e = e0 + tid; e < a0; e += nt) fold(base[e]);
            const size_t ng = (e1 - a0) / 4;
            size_t g = tid;
            for (; g + nt < ng; g += 2 * (size_t)nt) {  // two 16-byte loads in flight per thread
                const BcfVec4 x = bcf_load16i(base + a0 + g * 4), y = bcf_load16i(base + a0 + (g + nt) * 4);
                for (int i = 0; i < 4; i++) fold(x.v[i]);
                for (int i = 0; i < 4; i++) fold(y.v[i]);
            }
            if (g < ng) {
                const BcfVec4 x = bcf_load16i(base + a0 + g * 4);
                for (int i = 0; i < 4; i++) fold(x.v[i]);
            }
            for (size_t e = a0 + ng * 4 + tid; e < e1; e += nt) fold(base[e]);
        }
        bcf_mm_merge(s_mm + 2 * k, mn, mx);
    }
}
__host__ __device__ inline void bcf_scan_flush(const BcfEnc& E, uint32_t s, const int32_t* s_mm, uint32_t tid, uint32_t nt) {
    for (uint32_t i = tid; i < E.n_fields; i += nt) bcf_mm_merge(E.mm + ((size_t)s * E.n_fields + i) * 2, s_mm[2 * i], s_mm[2 * i + 1]);
}

// ---- stage: layout. One thread per site. ----
__host__ __device__ inline void bcf_layout_site(const BcfEnc& E, uint32_t s) {
    BcfSitePlan& P = E.plan[s];
    const uint32_t km = bcf_keep_mask(E, s);
    const uint32_t n_keep = (uint32_t)bcf_popc(km), N = E.n_samples;
    if (n_keep < 2) {  // ans.reset(): no record
        P.keep_mask = 0;
        P.l_shared = P.l_indiv = 0;
        P.n_keep = (uint8_t)n_keep;
        P.trimmed = 1;
        P.n_blk = 0;
        E.rec_off[s] = 0;
        E.hdr_off[s] = 0;
        return;
    }
    const bool trimmed = n_keep != E.n_allele[s];
    const uint32_t n_blk = E.n_fields + 2;
    P.keep_mask = km;
    P.n_keep = (uint8_t)n_keep;
    P.trimmed = trimmed ? 1 : 0;
    P.n_blk = (uint8_t)n_blk;
    uint32_t off = 0;
    for (uint32_t b = 0; b < n_blk; b++) {
        int32_t id;
        uint32_t cnt, width;
        int type;
        if (b == 0) {  // GT: htslib GT ints of alleles < 32 always fit int8
            id = E.m.id_GT;
            cnt = 2;
            width = 1;
            type = BCF_BT_INT8;
        } else if (b == n_blk - 1) {
            id = E.m.id_RNC;
            cnt = 2;
            width = 1;
            type = BCF_BT_CHAR;
        } else {
            const uint32_t k = b - 1;
            id = E.m.id_field[k];
            const int32_t* mm = E.mm + ((size_t)s * E.n_fields + k) * 2;
            if (E.fkind[k] == GLX_FK_FT || E.fkind[k] == GLX_FK_STRING) {
                cnt = mm[1] < 1 ? 1u : (uint32_t)mm[1];
                width = 1;
                type = BCF_BT_CHAR;
            } else {
                cnt = bcf_count_new(E.fvl[k], E.fld_cnt[k][s], n_keep, trimmed);
                if (E.ftype[k] == GLX_TYPE_FLOAT) {
                    width = 4;
                    type = BCF_BT_FLOAT;
                } else {
                    width = (uint32_t)bcf_int_width(mm[0], mm[1]);
                    type = width == 1 ? BCF_BT_INT8 : (width == 2 ? BCF_BT_INT16 : BCF_BT_INT32);
                }
            }
        }
        BcfWrite w{P.blk_hdr[b]};
        bcf_enc_int1(w, id);
        bcf_enc_size(w, cnt, type);
        const uint32_t hl = (uint32_t)(w.p - P.blk_hdr[b]);
        P.blk_off[b] = off;
        P.blk_cnt[b] = (uint16_t)cnt;
        P.blk_width[b] = (uint8_t)width;
        P.blk_hdr_len[b] = (uint8_t)hl;
        off += hl + N * cnt * width;
    }
    P.l_indiv = off;
    BcfCount c;
    bcf_emit_shared(E, s, km, n_blk, c);
    P.l_shared = c.n;
    E.rec_off[s] = 8ull + c.n + off;
    E.hdr_off[s] = 8ull + c.n;
}

// ---- stage: headers. One thread per site (after the offset scans). ----
__host__ __device__ inline void bcf_header_site(const BcfEnc& E, uint32_t s) {
    const BcfSitePlan& P = E.plan[s];
    if (!P.keep_mask) return;
    BcfWrite w{E.hdr_pool + E.hdr_off[s]};
    bcf_put32(w, P.l_shared);
    bcf_put32(w, P.l_indiv);
    bcf_emit_shared(E, s, P.keep_mask, P.n_blk, w);
}

// ---- stage: span. Gather bytes [lo, hi) of the record stream into buf[0 .. hi-lo). s_src: GLX_BCF_G_MAX uint16 ----
__host__ __device__ __forceinline__ uint32_t bcf_first_site(const BcfEnc& E, unsigned long long pos) {  // first s with rec_off[s+1] > pos
    uint32_t a = 0, b = E.n_sites;
    while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if (E.rec_off[mid + 1] > pos) b = mid;
        else a = mid + 1;
    }
    return a;
}
__host__ __device__ __forceinline__ uint32_t bcf_conv(int32_t v, uint32_t width) {
    // bcf_int32_missing / vector_end (0x80000000 / 0x80000001) -> the narrow type's own sentinels (0x80 / 0x81, 0x8000 / 0x8001);
    // every other value of a vector typed int8 / int16 already carries its sign bit in the low byte(s)
    if (width == 1) return ((uint32_t)v & 0xffu) | ((uint32_t)v >> 24 & 0x80u);
    if (width == 2) return ((uint32_t)v & 0xffffu) | ((uint32_t)v >> 16 & 0x8000u);
    return (uint32_t)v;
}
// where byte i of the span lives in the staging buffer: linear (raw stream), or with 4 bytes of padding after every 64
// (BGZF stage: its workers walk 64-byte chunks, and a 68-byte stride spreads them over the shared-memory banks)
struct BcfLinear {
    static __host__ __device__ __forceinline__ size_t phys(unsigned long long i) { return (size_t)i; }
    static __host__ __device__ __forceinline__ uint32_t phys32(uint32_t i) { return i; }
};
struct BcfPadded {
    static __host__ __device__ __forceinline__ size_t phys(unsigned long long i) { return (size_t)(i + ((i >> 6) << 2)); }
    static __host__ __device__ __forceinline__ uint32_t phys32(uint32_t i) { return i + ((i >> 6) << 2); }
};

// The span's piece table (shared memory): every run of bytes of the records overlapping the span — header + shared part,
// each FORMAT block's key/descriptor bytes, each FORMAT block's data — with where its bytes come from. Built by the whole
// block (one thread per piece, so the plan loads of all pieces are in flight together), then the span is filled word by
// word: a thread finds its word's piece by binary search and converts the source values, 16 independent words per thread.
// profiling builds (-DGLX_FILL_PROFILE, glx_bcf.cu): thread 0's cycles per part of the gather into glx_bgzf_phase_cycles[10..]
#if defined(GLX_FILL_PROFILE) && defined(GLX_FILL_PROFILE_HERE) && defined(__CUDA_ARCH__)
#define GLX_FILL_MARK(i)                                                                               \
    do {                                                                                               \
        if (tid == 0) {                                                                                \
            const long long now_ = clock64();                                                          \
            atomicAdd(&::glx_bgzf_phase_cycles[10 + (i)], (unsigned long long)(now_ - glx_fill_mark_t)); \
            glx_fill_mark_t = now_;                                                                    \
        }                                                                                              \
    } while (0)
#define GLX_FILL_MARK_INIT() long long glx_fill_mark_t = clock64()
#else
#define GLX_FILL_MARK(i) ((void)0)
#define GLX_FILL_MARK_INIT() ((void)0)
#endif
#ifndef GLX_BCF_FILL_UNROLL
#define GLX_BCF_FILL_UNROLL 4
#endif
#ifndef GLX_BCF_ITEM_SHIFT
#define GLX_BCF_ITEM_SHIFT 11  // log2 of the bytes of one fill item (measured on configs[1]: 256 B items 10 % slower than 1 KB — the per-item piece search — 2 KB 1.6 % faster)
#endif
#define GLX_BCF_TAB_MAX 256
#define GLX_BCF_REC_MAX 24
enum { BCF_PK_BYTES = 0, BCF_PK_GT = 1, BCF_PK_INT = 2, BCF_PK_RAW4 = 3, BCF_PK_TRIM = 4, BCF_PK_STR = 5 };
struct BcfSpanTab {
    int32_t start[GLX_BCF_TAB_MAX + 1];    // span-relative offset of the piece's first byte (negative: begins before the span); non-decreasing; [n] = end
    const void* src[GLX_BCF_TAB_MAX];
    uint32_t info[GLX_BCF_TAB_MAX];  // kind | width << 4 | values per sample after trimming << 8 | values per sample in the source << 20
    uint32_t km[GLX_BCF_TAB_MAX];    // keep mask of the site (GT renumbering, trimmed vectors)
    uint32_t dist[GLX_BCF_TAB_MAX];  // FORMAT data: distance to the same block's data in the previous record (0: none / other layout)
    uint8_t fld[GLX_BCF_TAB_MAX];    // lifted field index (trimmed vectors, strings)
    // staged copies of what the pieces are made from: one round of independent global loads instead of dependent chains
    BcfSitePlan plan[GLX_BCF_REC_MAX + 1];           // [0] = the site before the first, [1 + j] = site site0 + j
    unsigned long long roff[GLX_BCF_REC_MAX + 3];    // [0] = rec_off[site0 - 1], [1 + j] = rec_off[site0 + j]
    unsigned long long hoff[GLX_BCF_REC_MAX + 1];
    unsigned long long foff[GLX_BCF_REC_MAX][GLX_MAX_FIELDS];
    uint16_t fcnt[GLX_BCF_REC_MAX][GLX_MAX_FIELDS];
    uint32_t n, n_rec, site0, site_next, complete;  // complete: the table holds every piece of the span (one round)
    uint32_t next_item;                             // bcf_tab_fill: the next 1 KB item nobody has taken yet
    long long cov_lo, cov_hi;                       // span-relative byte range the table covers
};

// source slot of new slot jn of a trimmed Number=vl vector
__host__ __device__ __forceinline__ uint32_t bcf_nth_kept(uint32_t km, uint32_t r) {  // index of the r-th (0-based) set bit
    for (uint32_t a = 0; a < 32; a++)
        if (km >> a & 1u) {
            if (r == 0) return a;
            r--;
        }
    return 31;
}
__host__ __device__ __forceinline__ uint32_t bcf_slot_src(uint8_t vl, uint32_t jn, uint32_t km) {
    if (vl == 'R') return bcf_nth_kept(km, jn);
    if (vl == 'A') return bcf_nth_kept(km, jn + 1) - 1;
    uint32_t ra, rb;
    bcf_gt_alleles(jn, ra, rb);
    const uint32_t ia = bcf_nth_kept(km, ra), ib = bcf_nth_kept(km, rb);
    return ib * (ib + 1) / 2 + ia;
}

// Pieces of the records of sites [s_begin, ...) that overlap span [lo, hi): as many records as the table holds. Every thread calls it.
__host__ __device__ inline void bcf_tab_build(const BcfEnc& E, BcfSpanTab& T, unsigned long long lo, unsigned long long hi, uint32_t s_begin, uint32_t tid,
                                              uint32_t nt) {
    const uint32_t F = E.n_fields, slots = 1 + 2 * (F + 2);  // pieces per record
    const uint32_t max_rec = GLX_BCF_TAB_MAX / slots < GLX_BCF_REC_MAX ? GLX_BCF_TAB_MAX / slots : GLX_BCF_REC_MAX;
    GLX_FILL_MARK_INIT();
    // ---- stage: offsets, plans and field layouts of sites s_begin - 1 .. s_begin + max_rec (every load independent) ----
    for (uint32_t j = tid; j <= max_rec + 1; j += nt) {
        const long long s = (long long)s_begin - 1 + j;
        T.roff[j] = s < 0 ? 0ull : (s <= (long long)E.n_sites ? E.rec_off[s] : ~0ull);
        if (j >= 1 && j <= max_rec) T.hoff[j - 1] = s < (long long)E.n_sites ? E.hdr_off[s] : 0ull;
    }
    {
        constexpr uint32_t PW = sizeof(BcfSitePlan) / 4;
        static_assert(sizeof(BcfSitePlan) % 4 == 0, "plans are copied as words");
        uint32_t* dst = reinterpret_cast<uint32_t*>(T.plan);
        for (uint32_t i = tid; i < (max_rec + 1) * PW; i += nt) {
            const long long s = (long long)s_begin - 1 + i / PW;
            dst[i] = (s >= 0 && s < (long long)E.n_sites) ? reinterpret_cast<const uint32_t*>(E.plan + s)[i % PW] : 0u;
        }
    }
    for (uint32_t i = tid; i < max_rec * F; i += nt) {
        const uint32_t j = i / F, k = i - j * F, s = s_begin + j;
        T.fcnt[j][k] = s < E.n_sites ? E.fld_cnt[k][s] : (uint16_t)0;
        T.foff[j][k] = s < E.n_sites ? E.fld_off[k][s] : 0ull;
    }
    GLX_BLOCK_SYNC();
    GLX_FILL_MARK(0);
    if (tid == 0) {
        uint32_t nr = 0;
        while (nr < max_rec && s_begin + nr < E.n_sites && T.roff[1 + nr] < hi) nr++;
        T.n_rec = nr;
        T.next_item = 0;
        T.site0 = s_begin;
        T.site_next = s_begin + nr;
        T.n = nr * slots;
        T.cov_lo = (long long)(T.roff[1] > lo ? T.roff[1] - lo : 0);
        const unsigned long long end = nr ? T.roff[1 + nr] : lo;  // start of the first record not in the table
        T.cov_hi = (long long)((end < hi ? end : hi) - lo);
        T.start[nr * slots] = (int32_t)(long long)(end - lo);
    }
    GLX_BLOCK_SYNC();
    GLX_FILL_MARK(1);
    const uint32_t N = E.n_samples;
    for (uint32_t idx = tid; idx < T.n; idx += nt) {
        const uint32_t j = idx / slots, p = idx - j * slots, s = s_begin + j;
        const unsigned long long r0 = T.roff[1 + j];
        const bool has = T.roff[2 + j] > r0;
        const BcfSitePlan& P = T.plan[1 + j];
        long long st = (long long)(r0 - lo);
        const void* src = nullptr;
        uint32_t info = BCF_PK_BYTES | 1u << 4, dist = 0, fld = 0;
        if (has) {
            const uint32_t hdr_len = 8 + P.l_shared;
            if (p == 0) {
                src = E.hdr_pool + T.hoff[j];
            } else {
                const uint32_t b = (p - 1) >> 1;
                if (b >= P.n_blk) {
                    st = (long long)(T.roff[2 + j] - lo);  // unused slot: empty piece at the record's end
                } else {
                    const unsigned long long bstart = r0 + hdr_len + P.blk_off[b];
                    if ((p - 1) & 1) {  // data
                        const unsigned long long dstart = bstart + P.blk_hdr_len[b];
                        st = (long long)(dstart - lo);
                        const uint32_t cn = P.blk_cnt[b], w = P.blk_width[b];
                        if (b == 0) {
                            src = E.gt + (size_t)s * N * 2;
                            info = BCF_PK_GT | 1u << 4 | 2u << 8 | 2u << 20;
                        } else if (b == (uint32_t)P.n_blk - 1) {
                            src = E.rnc + (size_t)s * N * 2;
                            info = BCF_PK_BYTES | 1u << 4 | 2u << 8 | 2u << 20;
                        } else {
                            const uint32_t k = b - 1, cnt = T.fcnt[j][k];
                            fld = k;
                            src = E.fld[k] + T.foff[j][k];
                            uint32_t kind;
                            if (E.fkind[k] == GLX_FK_FT || E.fkind[k] == GLX_FK_STRING) kind = BCF_PK_STR;
                            else if (P.trimmed && bcf_is_arg(E.fvl[k])) kind = BCF_PK_TRIM;
                            else if (w == 4) kind = BCF_PK_RAW4;
                            else kind = BCF_PK_INT;
                            info = kind | w << 4 | (cn & 0xfffu) << 8 | (cnt & 0xfffu) << 20;
                        }
                        // the same block's data in the record just before, if that site has a record of the same layout
                        int u = (int)j;  // staged index of the nearest earlier site that has a record (u = 0: the site before s_begin)
                        while (u >= 0 && !(T.roff[u + 1] > T.roff[u])) u--;
                        if (u >= 0 && (u > 0 || s_begin > 0)) {
                            const BcfSitePlan& Q = T.plan[u];
                            if (Q.keep_mask && Q.n_blk == P.n_blk && Q.blk_cnt[b] == P.blk_cnt[b] && Q.blk_width[b] == P.blk_width[b]) {
                                const unsigned long long pstart = T.roff[u] + 8 + Q.l_shared + Q.blk_off[b] + Q.blk_hdr_len[b];
                                if (dstart - pstart <= 32768) dist = (uint32_t)(dstart - pstart);
                            }
                        }
                    } else {  // the block's typed key + descriptor
                        st = (long long)(bstart - lo);
                        src = P.blk_hdr[b];  // the staged copy: valid until the table is rebuilt
                    }
                }
            }
        }
        T.start[idx] = (int32_t)st;
        T.src[idx] = src;
        T.info[idx] = info;
        T.km[idx] = has ? P.keep_mask : 0u;
        T.dist[idx] = dist;
        T.fld[idx] = (uint8_t)fld;
    }
    GLX_BLOCK_SYNC();
    GLX_FILL_MARK(2);
}

// piece holding span byte x (x inside the table's coverage): last piece with start <= x
__host__ __device__ __forceinline__ uint32_t bcf_tab_find(const BcfSpanTab& T, long long x_) {
    const int32_t x = (int32_t)x_;
    uint32_t a = 0, b = T.n;
    while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if (T.start[mid] <= x) a = mid + 1;
        else b = mid;
    }
    return a - 1;
}
// one byte of a piece, any kind (word-straddling positions, trimmed vectors, strings)
__host__ __device__ inline uint8_t bcf_piece_byte(const BcfEnc& E, const BcfSpanTab& T, uint32_t p, unsigned long long off) {
    const uint32_t info = T.info[p], kind = info & 15u, w = info >> 4 & 15u, cn = info >> 8 & 0xfffu, cnt = info >> 20;
    switch (kind) {
        case BCF_PK_BYTES: return static_cast<const uint8_t*>(T.src[p])[off];
        case BCF_PK_GT: {
            const int a = static_cast<const int8_t*>(T.src[p])[off];
            return a < 0 ? (uint8_t)0 : (uint8_t)((bcf_popc(T.km[p] & ((1u << a) - 1u)) + 1) << 1);
        }
        case BCF_PK_INT: return (uint8_t)(bcf_conv(static_cast<const int32_t*>(T.src[p])[off / w], w) >> (8 * (off % w)));
        case BCF_PK_RAW4: return (uint8_t)((uint32_t) static_cast<const int32_t*>(T.src[p])[off >> 2] >> (8 * (off & 3)));
        case BCF_PK_STR: {
            const unsigned long long n = off / cn;
            return bcf_ft_char(E.m, (uint32_t) static_cast<const int32_t*>(T.src[p])[n * cnt], (uint32_t)(off - n * cn));
        }
        default: {  // BCF_PK_TRIM: bcf_remove_allele_set's rewrite — surviving slots, end-of-vector after the sample's first one
            const uint32_t k = T.fld[p];
            const bool flt = E.ftype[k] == GLX_TYPE_FLOAT;
            const int32_t vend = flt ? (int32_t)GLX_FLOAT_VECTOR_END_BITS : GLXB_VEND;
            const unsigned long long e = off / w, n = e / cn;
            const uint32_t jn = (uint32_t)(e - n * cn), j = bcf_slot_src(E.fvl[k], jn, T.km[p]);
            const int32_t* v = static_cast<const int32_t*>(T.src[p]) + n * cnt;
            int32_t x = v[j];
            for (uint32_t i = 0; i < j; i++)
                if (v[i] == vend) {
                    x = vend;
                    break;
                }
            return (uint8_t)((flt ? (uint32_t)x : bcf_conv(x, w)) >> (8 * (off % w)));
        }
    }
}

// byte `off` of a plainly typed piece (INT / RAW4 / GT / BYTES), no divisions: w is 1, 2 or 4
__host__ __device__ __forceinline__ uint32_t bcf_fast_byte(const void* src, uint32_t kind, uint32_t w, uint32_t km, uint32_t off) {
    if (kind == BCF_PK_BYTES) return static_cast<const uint8_t*>(src)[off];
    if (kind == BCF_PK_GT) {
        const int a = static_cast<const int8_t*>(src)[off];
        return a < 0 ? 0u : (uint32_t)((bcf_popc(km & ((1u << a) - 1u)) + 1) << 1);
    }
    const uint32_t sh = w >> 1;  // log2(w)
    return (bcf_conv(static_cast<const int32_t*>(src)[off >> sh], w) >> (8 * (uint32_t)(off & (w - 1)))) & 0xffu;
}
// One aligned 4-byte word of the span taken wholly from a plainly typed piece: byte offset `off` of the piece's data.
__host__ __device__ __forceinline__ uint32_t bcf_piece_word(const void* src, uint32_t kind, uint32_t w, uint32_t km, uint32_t off) {
    if (kind == BCF_PK_INT && w == 1) {
        const int32_t* v = static_cast<const int32_t*>(src) + off;
        int32_t a, b, c, d;
#if defined(__CUDA_ARCH__)
        if ((reinterpret_cast<uintptr_t>(v) & 15) == 0) {
            const int4 q = __ldg(reinterpret_cast<const int4*>(v));
            a = q.x, b = q.y, c = q.z, d = q.w;
        } else
#endif
        {
            a = v[0], b = v[1], c = v[2], d = v[3];
        }
        return bcf_conv(a, 1) | bcf_conv(b, 1) << 8 | bcf_conv(c, 1) << 16 | bcf_conv(d, 1) << 24;
    }
    if (kind == BCF_PK_INT && w == 2) {  // three values when the word starts in the middle of one
        const int32_t* v = static_cast<const int32_t*>(src) + (off >> 1);
        unsigned long long x = (unsigned long long)bcf_conv(v[0], 2) | (unsigned long long)bcf_conv(v[1], 2) << 16;
        if (off & 1) x |= (unsigned long long)bcf_conv(v[2], 2) << 32;
        return (uint32_t)(x >> (8 * (off & 1)));
    }
    if (kind == BCF_PK_INT || kind == BCF_PK_RAW4) {
        const int32_t* v = static_cast<const int32_t*>(src) + (off >> 2);
        unsigned long long x = (uint32_t)v[0];
        if (off & 3) x |= (unsigned long long)(uint32_t)v[1] << 32;
        return (uint32_t)(x >> (8 * (off & 3)));
    }
    if (kind == BCF_PK_GT) {
        const int8_t* g = static_cast<const int8_t*>(src) + off;
        uint32_t word = 0;
        for (int i = 0; i < 4; i++) {
            const int al = g[i];
            word |= (al < 0 ? 0u : (uint32_t)((bcf_popc(km & ((1u << al) - 1u)) + 1) << 1)) << (8 * i);
        }
        return word;
    }
    const uint8_t* c = static_cast<const uint8_t*>(src) + off;
    return c[0] | c[1] << 8 | c[2] << 16 | (uint32_t)c[3] << 24;
}
template <class Addr>
__host__ __device__ __forceinline__ void bcf_put_word(uint8_t* buf, uint32_t x0, uint32_t word) {  // both layouts keep an aligned word contiguous
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint32_t*>(buf + Addr::phys32(x0)) = word;
#else
    memcpy(buf + Addr::phys32(x0), &word, 4);  // the host build also runs odd span sizes
#endif
}

// Fill the covered bytes of the span from the table. The span is cut into items of 2^GLX_BCF_ITEM_SHIFT bytes; a warp takes the next untaken item (items
// cost up to 4x one another: 1 KB of 1-byte values reads 4 KB of int32 columns, 1 KB of GT reads 1 KB), walks the pieces that
// overlap it and converts their words lane by lane, GLX_BCF_FILL_UNROLL independent words per lane in flight.
template <class Addr>
__host__ __device__ inline void bcf_tab_fill(const BcfEnc& E, BcfSpanTab& T, uint8_t* buf, uint32_t tid, uint32_t nt) {
    if (T.cov_hi <= T.cov_lo) return;
    const uint32_t n_lanes = nt < 32 ? nt : 32, lane = tid % n_lanes;
    constexpr int IS = GLX_BCF_ITEM_SHIFT;
    // span-relative positions fit 32 bits (the span is 64 KB; a piece may begin up to its own length before the span)
    const int32_t cov_lo = (int32_t)T.cov_lo, cov_hi = (int32_t)T.cov_hi;
    const int32_t it0 = cov_lo >> IS, it1 = (cov_hi + (1 << IS) - 1) >> IS;
    for (;;) {
        int32_t it;
#if defined(__CUDA_ARCH__)
        {
            uint32_t k = 0;
            if (lane == 0) k = atomicAdd(&T.next_item, 1u);
            it = it0 + (int32_t)__shfl_sync(0xffffffffu, k, 0);
        }
#else
        it = it0 + (int32_t)T.next_item++;
#endif
        if (it >= it1) break;
        const int32_t a_it = (it << IS) < cov_lo ? cov_lo : (it << IS), e_it = ((it + 1) << IS) > cov_hi ? cov_hi : ((it + 1) << IS);
        for (uint32_t p = bcf_tab_find(T, a_it); p < T.n && T.start[p] < e_it; p++) {
            const int32_t base = T.start[p];
            const int32_t ps = base < a_it ? a_it : base, pe = T.start[p + 1] > e_it ? e_it : T.start[p + 1];
            if (pe <= ps) continue;
            const uint32_t info = T.info[p], kind = info & 15u, w = info >> 4 & 15u, km = T.km[p];
            const void* src = T.src[p];
            if (kind == BCF_PK_TRIM || kind == BCF_PK_STR) {  // rewritten vectors and strings: byte by byte
                for (int32_t x = ps + (int32_t)lane; x < pe; x += (int32_t)n_lanes) buf[Addr::phys32((uint32_t)x)] = bcf_piece_byte(E, T, p, (unsigned long long)(uint32_t)(x - base));
                continue;
            }
            int32_t wa = (ps + 3) & ~3, we = pe & ~3;
            if (wa > we) wa = we = pe;
            // the up to 3 + 3 bytes around the aligned words, in one go
            {
                const int32_t nh = wa - ps, ntl = pe - we;
                if ((int32_t)lane < nh + ntl) {
                    const int32_t x = (int32_t)lane < nh ? ps + (int32_t)lane : we + ((int32_t)lane - nh);
                    buf[Addr::phys32((uint32_t)x)] = (uint8_t)bcf_fast_byte(src, kind, w, km, (uint32_t)(x - base));
                }
                if (n_lanes < 6)  // the host build's single lane
                    for (int32_t i = (int32_t)n_lanes; i < nh + ntl; i++) {
                        const int32_t x = i < nh ? ps + i : we + (i - nh);
                        if (lane == 0) buf[Addr::phys32((uint32_t)x)] = (uint8_t)bcf_fast_byte(src, kind, w, km, (uint32_t)(x - base));
                    }
            }
            // aligned words, lane-strided; each lane keeps GLX_BCF_FILL_UNROLL independent loads in flight (the fill is bound by
            // memory latency, not by instructions: one dependent round trip per piece part otherwise)
            const int32_t step = 4 * (int32_t)n_lanes;
            for (int32_t xb = wa + 4 * (int32_t)lane; xb < we; xb += GLX_BCF_FILL_UNROLL * step) {
                uint32_t v[GLX_BCF_FILL_UNROLL];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int u = 0; u < GLX_BCF_FILL_UNROLL; u++) {
                    const int32_t x0 = xb + u * step;
                    v[u] = x0 < we ? bcf_piece_word(src, kind, w, km, (uint32_t)(x0 - base)) : 0u;
                }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int u = 0; u < GLX_BCF_FILL_UNROLL; u++) {
                    const int32_t x0 = xb + u * step;
                    if (x0 < we) bcf_put_word<Addr>(buf, (uint32_t)x0, v[u]);
                }
            }
        }
    }
}

// Gather bytes [lo, hi) of the record stream into buf (layout Addr). s_first = first site whose record reaches into the span
// (bcf_first_site(E, lo), or the span index table). Every thread calls it.
template <class Addr>
__host__ __device__ inline void bcf_fill_span(const BcfEnc& E, unsigned long long lo, unsigned long long hi, uint32_t s_first, uint8_t* buf, BcfSpanTab& T,
                                              uint32_t tid, uint32_t nt) {
    uint32_t s = s_first, rounds = 0;
    GLX_FILL_MARK_INIT();
    for (;;) {
        bcf_tab_build(E, T, lo, hi, s, tid, nt);
        GLX_FILL_MARK(3);
        bcf_tab_fill<Addr>(E, T, buf, tid, nt);
        GLX_FILL_MARK(4);
        const uint32_t next = T.site_next;
        const bool more = next < E.n_sites && T.n_rec > 0 && T.roff[1 + T.n_rec] < hi;
        rounds++;
        GLX_BLOCK_SYNC();  // the table is rebuilt (or read by the caller) after everyone has finished with it
        GLX_FILL_MARK(5);
        if (!more) break;
        s = next;
    }
    if (tid == 0) T.complete = rounds == 1 ? 1u : 0u;
    GLX_BLOCK_SYNC();
}

}  // namespace glxd
