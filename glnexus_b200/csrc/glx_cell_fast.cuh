// glx_cell_fast.cuh — per-lane logic of the tiled kernel: a lane owns one sample, walks a tile of consecutive sites
// with a monotone cursor into that sample's sorted records, and finishes in registers every cell whose query range
// holds at most one record, that record being a reference band (`REF,<*>`). Those are >95% of the cells of a gVCF
// cohort. Everything else (variant records, multi-record cells, monoallelic sites, malformed fields) is handed to
// the general path (glx_cell_general.cuh) through a worklist, so the two paths together are exactly genotype_site.
//
// For a single reference-band record the reference's per-cell pipeline collapses to closed forms
// (src/genotyper.cc:262-352,410-426; src/genotyper_utils.h:285-355,417-517,555-675):
//   kept records      = {band} iff it overlaps site.pos, else MissingData
//   PartialData       iff !allow_partial_data and the band does not contain site.pos
//   InputNonCalled    iff the band's GT is not 0/0 (and it was not haploid)
//   0/0 or Depth      by min_ref_depth = MIN_DP vs required_dp
//   lifted fields     = the band's own values; allele map is [0,-1] so only slot 0 of allele/genotype-numbered
//                       fields receives a value; PL2 back-fills through the symbolic allele
// The band's field values are resolved once per (lane, record) and cached; a lane re-reads nothing while
// consecutive sites fall in the same band.
#pragma once
#include "glx_cell_general.cuh"

namespace glxd {

#define GLX_FAST_RV 8       // resolved values cached per lane
#define GLX_TILE_SITES 64   // sites walked by one CTA
#define GLX_TILE_THREADS 128
#define GLX_STAGE_MAX 32    // per-sample vector length up to which stores are transposed through shared memory

struct FastPlan {
    uint8_t enabled, n_rv;
    uint8_t rv0[GLX_MAX_FIELDS], nrv[GLX_MAX_FIELDS];
};

inline FastPlan make_fast_plan(const DCfg& c) {
    FastPlan p;
    memset(&p, 0, sizeof p);
    p.enabled = !c.squeeze && !c.has_string_field;
    int n = 0;
    for (uint32_t k = 0; k < c.n_fields; k++) {
        const DFld& f = c.f[k];
        int need = 0;
        switch (f.kind) {
            case GLX_FK_NUMERIC:
            case GLX_FK_DP: need = f.number == GLX_NUM_BASIC ? f.count : 1; break;
            case GLX_FK_AD: need = 1; break;
            case GLX_FK_PL: need = 0; break;
            case GLX_FK_PL2: need = 3; break;
            case GLX_FK_FT: need = 1; break;
            default: p.enabled = 0;
        }
        p.rv0[k] = (uint8_t)n;
        p.nrv[k] = (uint8_t)need;
        n += need;
        if (n > GLX_FAST_RV) {
            p.enabled = 0;
            n = 0;
        }
    }
    p.n_rv = (uint8_t)n;
    return p;
}

// first record of [lo,hi) whose running max end exceeds qbeg (records with maxend <= qbeg cannot overlap the query)
__device__ __forceinline__ uint64_t first_candidate(const int32_t* __restrict__ maxend, uint64_t lo, uint64_t hi, int qbeg) {
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (maxend[mid] <= qbeg) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

struct LaneCursor {
    uint64_t base;  // first record of the lane's sample
    uint32_t n;     // records of the sample
    uint32_t c;     // first record (relative) with maxend > current qbeg
    int beg_c, end_c, maxend_c, beg_n;
};

__device__ __forceinline__ void cursor_load(const DRecs& R, LaneCursor& L) {
    if (L.c < L.n) {
        const uint64_t r = L.base + L.c;
        L.beg_c = R.beg[r];
        L.end_c = R.end[r];
        L.maxend_c = R.maxend[r];
        L.beg_n = (L.c + 1 < L.n) ? R.beg[r + 1] : 0x7fffffff;
    } else {
        L.beg_c = L.beg_n = L.maxend_c = 0x7fffffff;
        L.end_c = 0x7fffffff;
    }
}

__device__ __forceinline__ void cursor_init(const DRecs& R, LaneCursor& L, uint32_t sample, int qbeg) {
    L.base = R.ds_rec_off[sample];
    const uint64_t hi = R.ds_rec_off[sample + 1];
    L.n = (uint32_t)(hi - L.base);
    L.c = (uint32_t)(first_candidate(R.maxend, L.base, hi, qbeg) - L.base);
    cursor_load(R, L);
}

// `back`: this site's qbeg is smaller than the previous site's (query hulls are not monotone in general)
__device__ __forceinline__ void cursor_seek(const DRecs& R, LaneCursor& L, int qbeg, bool back) {
    if (back) {
        bool moved = false;
        while (L.c > 0 && R.maxend[L.base + L.c - 1] > qbeg) {
            L.c--;
            moved = true;
        }
        if (moved) cursor_load(R, L);
    }
    while (L.c < L.n && L.maxend_c <= qbeg) {
        L.c++;
        cursor_load(R, L);
    }
}

enum { CT_SLOW = 0, CT_MISSING = 1, CT_PARTIAL = 2, CT_REF = 3 };
enum { RB_SLOW = 1, RB_NONCALLED = 2 };

// Resolve a reference band into the lane's value cache. Returns RB_* bits; mrd = its MIN_DP as min_ref_depth.
__device__ inline uint32_t resolve_ref_band(const DCfg& cfg, const FastPlan& plan, const DRecs& R, uint64_t r, int32_t* rv, int rvs, int& mrd) {
    const uint8_t flags = R.flags[r];
    const int na = R.n_allele[r];
    if (!(flags & GLX_RF_IS_REF) || na != 2 || (flags & GLX_RF_NO_GT)) return RB_SLOW;
    int g0, g1;
    decode_gt(R.gt[r], flags, g0, g1);
    if ((!gt_is_missing(g0) && (g0 == GLXD_INT_VEND || gt_allele(g0) >= na)) || (!gt_is_missing(g1) && (g1 == GLXD_INT_VEND || gt_allele(g1) >= na)))
        return RB_SLOW;
    uint32_t bits = 0;
    if (!(flags & GLX_RF_HAPLOID) && (gt_is_missing(g0) || gt_allele(g0) != 0 || gt_is_missing(g1) || gt_allele(g1) != 0)) bits |= RB_NONCALLED;
    const int32_t* p = nullptr;
    if (col_get(R, cfg.col_ref_dp, r, p) != 1) return RB_SLOW;  // "reference depth FORMAT field is missing or malformed"
    mrd = (int)(unsigned)p[0];
    for (uint32_t k = 0; k < cfg.n_fields; k++) {
        const DFld& f = cfg.f[k];
        int32_t* c = rv + (size_t)plan.rv0[k] * rvs;
        const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
        const int32_t dflt = f.default_zero ? 0 : missing;
        switch (f.kind) {
            case GLX_FK_NUMERIC:
            case GLX_FK_DP:
            case GLX_FK_AD:
            case GLX_FK_PL: {
                int n_val = f.count;
                if (f.number == GLX_NUM_ALT) n_val = 1;
                else if (f.number == GLX_NUM_ALLELES) n_val = 2;
                else if (f.number == GLX_NUM_GENOTYPE) n_val = 3;
                const int32_t* v = nullptr;
                bool found = false;
                for (int ci = 0; ci < f.n_cols && !found; ci++) {
                    const int rvv = col_get(R, f.cols[ci], r, v);
                    if (rvv == -2) return RB_SLOW;
                    if (rvv < 0) continue;
                    if (rvv != n_val) return RB_SLOW;
                    found = true;
                }
                if (!found && f.kind == GLX_FK_AD) {  // AD falls back to the reference depth as a 1-value field
                    const int rvv = col_get(R, cfg.col_ref_dp, r, v);
                    if (rvv == -2) return RB_SLOW;
                    if (rvv >= 0) {
                        if (rvv != 1) return RB_SLOW;
                        found = true;
                    }
                }
                if (f.kind == GLX_FK_PL) break;  // a lone band never yields a liftable PL vector; validated only
                if (f.number == GLX_NUM_BASIC) {
                    for (int j = 0; j < f.count; j++) c[(size_t)j * rvs] = found ? v[j] : dflt;
                } else {
                    c[0] = (found && n_val > 0) ? v[0] : dflt;
                }
                break;
            }
            case GLX_FK_PL2: {
                const int32_t* buf = nullptr;
                const int rvv = col_get(R, cfg.col_pl, r, buf);
                int32_t p0 = 0, p1 = 0, p2 = 0;
                if (rvv > 0) {
                    if (rvv != 3) return RB_SLOW;
                    p0 = buf[0];
                    p1 = buf[1];
                    p2 = buf[2];
                    const bool zero = p0 == 0 || p1 == 0 || p2 == 0;
                    const bool other = (p0 != 0 && p0 != GLXD_INT_MISSING) || (p1 != 0 && p1 != GLXD_INT_MISSING) || (p2 != 0 && p2 != GLXD_INT_MISSING);
                    if (!(zero && other)) p0 = p1 = p2 = 0;
                    else {
                        if (p0 == GLXD_INT_MISSING) p0 = 990;
                        if (p1 == GLXD_INT_MISSING) p1 = 990;
                        if (p2 == GLXD_INT_MISSING) p2 = 990;
                    }
                }
                c[0] = p0;
                c[(size_t)rvs] = p1;
                c[(size_t)2 * rvs] = p2;
                break;
            }
            case GLX_FK_FT: {
                const uint32_t m = R.filt[r];
                const int nu = __popc(m);
                int32_t v = 0;
                if (nu == 0) v = 0;
                else if (nu == 1 || f.combi == GLX_COMBI_SEMICOLON) v = (int32_t)m;
                else if (f.combi == GLX_COMBI_MISSING) v = 0;
                else return RB_SLOW;
                c[0] = v;
                break;
            }
            default: return RB_SLOW;
        }
    }
    return bits;
}

struct FastCell {
    int type;        // CT_*
    int a0, a1;      // output GT allele indices (-1 missing)
    int rnc0, rnc1;  // RNC_* codes
};

// Classify the lane's cell at a site and, unless CT_SLOW, decide GT / RNC. `resolved` caches which record rv[] holds.
__device__ __forceinline__ FastCell fast_classify(const DCfg& cfg, const FastPlan& plan, const DRecs& R, LaneCursor& L, int sbeg, int send, int qbeg, int qend,
                                                  bool slow_site, bool back, uint32_t& resolved, uint32_t& rbits, int& mrd, int32_t* rv, int rvs) {
    FastCell out;
    out.a0 = out.a1 = -1;
    out.rnc0 = out.rnc1 = RNC_M;
    out.type = CT_MISSING;
    cursor_seek(R, L, qbeg, back);
    if (slow_site) {
        out.type = CT_SLOW;
        return out;
    }
    if (L.c >= L.n || L.beg_c >= qend) return out;  // nothing overlaps the query range: MissingData
    if (L.beg_n < qend) {                            // a second record may overlap it
        out.type = CT_SLOW;
        return out;
    }
    if (resolved != L.c) {
        rbits = resolve_ref_band(cfg, plan, R, L.base + L.c, rv, rvs, mrd);
        resolved = L.c;
    }
    if (rbits & RB_SLOW) {
        out.type = CT_SLOW;
        return out;
    }
    if (!(L.end_c > sbeg && L.beg_c < send)) return out;  // band only inside the unification hull: not kept
    if (!cfg.allow_partial && !(L.beg_c <= sbeg && L.end_c >= send)) {
        out.type = CT_PARTIAL;
        out.rnc0 = out.rnc1 = RNC_P;
        return out;
    }
    out.type = CT_REF;
    if (rbits & RB_NONCALLED) {
        out.rnc0 = out.rnc1 = RNC_I;
    } else if ((unsigned long long)(long long)mrd >= (unsigned long long)cfg.required_dp) {
        out.a0 = out.a1 = 0;
        out.rnc0 = out.rnc1 = RNC_NA;
    } else {
        out.rnc0 = out.rnc1 = RNC_D;
    }
    return out;
}

// value of slot j = genotype (a,b), b <= a (or allele/basic index j) of field k for a finished fast cell
__device__ __forceinline__ int32_t fast_value(const DFld& f, const FastPlan& plan, int k, int type, int count, int j, int a, int b, const int32_t* rv, int rvs) {
    const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
    const bool vec_rule = f.type == GLX_TYPE_INT && f.number != GLX_NUM_BASIC && count > 1;  // all-missing vector => [missing, vector_end]
    if (type != CT_REF) {  // every field censored (MissingData / PartialData)
        if (f.kind == GLX_FK_PL2 || f.kind == GLX_FK_FT) return 0;
        return (vec_rule && j == 1) ? GLXD_INT_VEND : missing;
    }
    const int32_t* c = rv + (size_t)plan.rv0[k] * rvs;
    switch (f.kind) {
        case GLX_FK_PL2: return a == 0 ? c[0] : (b == 0 ? c[(size_t)rvs] : c[(size_t)2 * rvs]);
        case GLX_FK_FT: return c[0];
        case GLX_FK_PL: return (vec_rule && j == 1) ? GLXD_INT_VEND : GLXD_INT_MISSING;
        default: break;
    }
    if (f.number == GLX_NUM_BASIC) return c[(size_t)j * rvs];
    const int32_t dflt = f.default_zero ? 0 : missing;
    const int32_t v0 = c[0];
    if (j == 0) return v0;
    if (vec_rule && j == 1 && v0 == GLXD_INT_MISSING && dflt == GLXD_INT_MISSING) return GLXD_INT_VEND;
    return dflt;
}

}  // namespace glxd
