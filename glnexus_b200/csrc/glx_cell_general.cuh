// glx_cell_general.cuh — the fully general (site, sample) cell: every branch of genotype_site for one
// single-sample dataset, written as a streaming computation over the sample's sorted record array.
//
// What it replaces, per cell (reference file:line, /root/reference):
//   record selection + span check      src/genotyper.cc:262-310 (prepare_dataset_records), types.h:174-215
//   preprocess_record                  src/genotyper.cc:35-80
//   revise_genotypes                   src/genotyper.cc:95-225, src/diploid.cc:107-140
//   update_min_ref_depth / depth gate  src/genotyper_utils.h:982-1152
//   translate_genotypes / _monoallelic src/genotyper.cc:363-680
//   update_format_fields + helpers     src/genotyper_utils.h:138-675, 933-976
//   censoring, squeeze, slot swap      src/genotyper.cc:796-830, 862-867
// No per-cell record list is materialised: the kept records are re-walked (they are a contiguous, short
// span of the sample's records) and per-slot value counts live in a byte scratch owned by the thread.
#pragma once
#include "glx_dev_types.cuh"

namespace glxd {

#define GLXD_INT_MISSING ((int32_t)0x80000000)
#define GLXD_INT_VEND ((int32_t)0x80000001)

enum { RNC_NA = 0, RNC_M, RNC_P, RNC_D, RNC_LOSTDEL, RNC_LOST, RNC_U, RNC_O, RNC_MONO, RNC_I };
__device__ __constant__ const char kRncChar[10] = {'.', 'M', 'P', 'D', '-', 'L', 'U', 'O', '1', 'I'};

__device__ __forceinline__ bool gt_is_missing(int v) { return (v >> 1) == 0; }
__device__ __forceinline__ int gt_allele(int v) { return (v >> 1) - 1; }
__device__ __forceinline__ int gt_unphased(int a) { return (a + 1) << 1; }
__device__ __forceinline__ unsigned n_genotypes(unsigned n) { return (n + 1) * n / 2; }
__device__ __forceinline__ int alleles_gt(int i, int j) { return i <= j ? j * (j + 1) / 2 + i : i * (i + 1) / 2 + j; }

// The batch fails with the status the reference would return: that of the first failing site and, within it, of the
// first failing dataset (genotype_site walks datasets in order). Cells are numbered site-major, so the minimum over
// (cell << 8 | -status) is that error.
__device__ __forceinline__ void report_error(const DOut& O, unsigned long long cell, int code /* positive = -status */) {
    atomicMin(O.error, (cell << 8) | (unsigned)code);
}

// decode the inline GT pair of a single-sample record as preprocess_record sees it (genotyper.cc:44-57)
__device__ __forceinline__ void decode_gt(uint32_t g, uint8_t flags, int& g0, int& g1) {
    int16_t a = (int16_t)(g & 0xffff), b = (int16_t)(g >> 16);
    int x = a == GLX_GT16_VECTOR_END ? GLXD_INT_VEND : (int)a;
    int y = b == GLX_GT16_VECTOR_END ? GLXD_INT_VEND : (int)b;
    if (flags & GLX_RF_HAPLOID) {
        g0 = 0;
        g1 = x;
    } else {
        g0 = x;
        g1 = y;
    }
}

// bcf_get_format_int32 / bcf_get_info_* through the record store: rv as htslib (-3 absent, -2 type clash)
__device__ __forceinline__ int col_get(const DRecs& R, int col, uint64_t r, const int32_t*& p) {
    if (col < 0) return -1;
    const size_t idx = (size_t)col * R.n_records + r;
    const uint16_t len = R.col_len[idx];
    if (len == GLX_LEN_ABSENT) return -3;
    if (len == 0xFFFE) return -2;
    if (len == 0xFFFD) {
        p = nullptr;
        return 0x7fffffff;
    }
    p = (len == 1) ? &R.col_val[idx] : R.pool + R.col_val[idx];
    return (int)len;
}

// One sample of a MULTI-sample dataset, presented as if its records were single-sample: `si` of the dataset's `ns` samples.
// The per-sample logic (allele depths, likelihood rescoring, FORMAT lifting) runs unchanged on this view; what the reference
// decides per DATASET across its samples (src/genotyper.cc:95-115, 335-347, 380-392, 474-476) is restated in
// genotype_cell_general's MULTI branches. Values of a column sit sample-major in the pool (bcf_get_format_*'s layout), the GT
// pairs in gt_pool.
struct MRecs : DRecs {
    const int32_t* gt_pool;
    uint32_t ns, si;
};
__device__ __forceinline__ int col_get(const MRecs& R, int col, uint64_t r, const int32_t*& p) {
    if (R.ns == 1) return col_get(static_cast<const DRecs&>(R), col, r, p);
    if (col < 0) return -1;
    const size_t idx = (size_t)col * R.n_records + r;
    const uint16_t len = R.col_len[idx];
    if (len == GLX_LEN_ABSENT) return -3;
    if (len == 0xFFFE) return -2;
    if (len == 0xFFFD) {
        p = nullptr;
        return 0x7fffffff;
    }
    p = R.pool + R.col_val[idx] + (size_t)R.si * len;
    return (int)len;
}
// the record's GT pair as preprocess_record sees it (htslib GT ints, phase bit cleared)
__device__ __forceinline__ void rec_gt(const DRecs& R, uint64_t r, uint8_t flags, int& g0, int& g1) { decode_gt(R.gt[r], flags, g0, g1); }
__device__ __forceinline__ void rec_gt_of(const MRecs& R, uint64_t r, uint8_t flags, uint32_t k, int& g0, int& g1) {
    if (R.ns == 1) {
        decode_gt(R.gt[r], flags, g0, g1);
        return;
    }
    const int32_t* gp = R.gt_pool + R.gt[r] + 2 * (size_t)k;
    g0 = gp[0];
    g1 = gp[1];
}
__device__ __forceinline__ void rec_gt(const MRecs& R, uint64_t r, uint8_t flags, int& g0, int& g1) { rec_gt_of(R, r, flags, R.si, g0, g1); }

// site.unification.find(allele(record range, dna)) — src/genotyper.cc:65-72
template <class RT, class ST>
__device__ inline int unif_lookup(const RT& R, const ST& S, uint32_t u0, uint32_t u1, int rbeg, int rend, uint32_t a) {
    const uint32_t len = R.al_len[a];
    const uint64_t pre = R.al_prefix[a];
    for (uint32_t u = u0; u < u1; u++) {
        if (S.unif_beg[u] != rbeg || S.unif_end[u] != rend || S.unif_len[u] != len || S.unif_prefix[u] != pre) continue;
        bool eq = true;
        if (len > 8) {
            const char* x = R.al_pool + R.al_str[a];
            const char* y = S.unif_pool + S.unif_str[u];
            for (uint32_t i = 8; i < len; i++)
                if (x[i] != y[i]) {
                    eq = false;
                    break;
                }
        }
        if (eq) return S.unif_to[u];
    }
    return -1;
}

struct SiteCtx {
    uint32_t s;
    int beg, end, qbeg, qend;
    int A;
    bool mono;
    uint32_t al0;     // offset of the site's alleles in allele_is_indel / hom_log_prior
    uint32_t u0, u1;  // unification entries
};

__device__ __forceinline__ SiteCtx load_site(const DSites& S, uint32_t s) {
    SiteCtx c;
    c.s = s;
    c.beg = S.beg[s];
    c.end = S.end[s];
    c.qbeg = S.qbeg[s];
    c.qend = S.qend[s];
    c.A = S.n_allele[s];
    c.mono = S.monoallelic[s] != 0;
    c.al0 = S.allele_off[s];
    c.u0 = S.unif_off[s];
    c.u1 = S.unif_off[s + 1];
    return c;
}

// allele_mapping / deletion_allele of preprocess_record. map[] has n_allele entries.
template <class RT, class ST>
__device__ inline void compute_allele_map(const RT& R, const ST& S, const SiteCtx& st, uint64_t r, int na, bool is_ref, int8_t* map,
                                          uint32_t& del_mask) {
    map[0] = 0;
    del_mask = 0;
    for (int i = 1; i < na; i++) map[i] = -1;
    if (is_ref) return;
    const int rbeg = R.beg[r], rend = R.end[r];
    const uint32_t a0 = R.allele_off[r];
    for (int i = 1; i < na; i++) {
        const uint8_t fl = R.al_flags[a0 + i];
        if (fl & GLX_AF_IS_DNA) map[i] = (int8_t)unif_lookup(R, S, st.u0, st.u1, rbeg, rend, a0 + i);
        if (fl & GLX_AF_DELETION) del_mask |= 1u << i;
    }
}

// keep test of prepare_dataset_records (genotyper.cc:283-300) for a record already known to overlap the query range
__device__ inline bool record_kept(const DRecs& R, const DSites& S, const SiteCtx& st, uint64_t r) {
    const int rbeg = R.beg[r], rend = R.end[r];
    if (rend > st.beg && rbeg < st.end) return true;
    if (R.flags[r] & GLX_RF_IS_REF) return false;
    const int na = R.n_allele[r];
    const uint32_t a0 = R.allele_off[r];
    for (int i = 1; i < na; i++)
        if ((R.al_flags[a0 + i] & GLX_AF_IS_DNA) && unif_lookup(R, S, st.u0, st.u1, rbeg, rend, a0 + i) >= 0) return true;
    return false;
}

// (int) of an out-of-range double is INT_MIN on x86-64 (cvttsd2si); the oracle pins the same
__device__ __forceinline__ int to_int_x86(double x) {
    if (!(x > -2147483649.0 && x < 2147483648.0)) return (int)0x80000000;
    return (int)x;
}

// revise_genotypes for one single-sample variant record. Returns 0 ok / positive error code.
// g0,g1 in/out (htslib GT ints), gq out (valid when revised).
template <class RT, class ST>
// `forced`: -1 the record's own GT decides whether it is rescored (single-sample dataset); 0 no; 1 yes; 2 yes, as a homozygous-ALT
// call (multi-sample dataset: the first sample with a reason decides for all of them, genotyper.cc:97-115 — revise_forced).
__device__ inline int revise_record(const DCfg& cfg, const RT& R, const ST& S, const SiteCtx& st, uint64_t r, int na, const int8_t* map, int& g0,
                                    int& g1, int& gq, bool& revised, int forced = -1) {
    revised = false;
    bool needs = false, homalt = false;
    if (forced < 0) {
        if (!gt_is_missing(g0) && map[gt_allele(g0)] == -1 && !st.mono) needs = true;
        else if (!gt_is_missing(g1) && map[gt_allele(g1)] == -1 && !st.mono) needs = true;
        else if (!gt_is_missing(g0) && !gt_is_missing(g1) && gt_allele(g0) > 0 && gt_allele(g0) == gt_allele(g1)) needs = homalt = true;
    } else {
        needs = forced > 0;
        homalt = forced == 2;
    }
    if (!needs) return 0;
    const int nGT = (int)n_genotypes(na);
    const int32_t* pl = nullptr;
    if (col_get(R, cfg.col_pl, r, pl) != nGT) return 1;  // Failure: couldn't find genotype likelihoods
    for (int k = 0; k < nGT; k++) {
        const int32_t x = pl[k];
        if (x != GLXD_INT_MISSING && x != GLXD_INT_VEND && x < 0) return 1;  // Invalid -> Failure (genotyper.cc:141-144)
    }
    const int32_t* gqp = nullptr;
    if (col_get(R, cfg.col_gq, r, gqp) != 1 || !gqp) return 1;
    const double lost_prior = (double)S.lost_log_prior[st.s];
    const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
    double map_gll = ninf, silver_gll = ninf;
    int map_i = -1, map_j = -1, map_gt = -1, silver_i = -1, silver_j = -1, silver_gt = -1;
    int g = 0;
    for (int j = 0; j < na; j++) {
        for (int i = 0; i <= j; i++, g++) {
            double prior = 0.0;
            const int mi = map[i], mj = map[j];
            if (mi == -1 || mj == -1) {
                if (!st.mono) prior = lost_prior;
            } else if (i > 0 && i == j) {
                prior = (double)S.hom_log_prior[st.al0 + mi];
            }
            if (cfg.calibration && prior != 0.0) {
                const bool has_indel = (mi > 0 && S.allele_is_indel[st.al0 + mi]) || (mj > 0 && S.allele_is_indel[st.al0 + mj]);
                const double x = __dmul_rn((double)(has_indel ? cfg.indel_cal : cfg.snv_cal), prior);
                prior = (x < 0.0) ? x : 0.0;  // std::min(0.0, x)
            }
            const int32_t x = pl[g];
            const double gll = (x == GLXD_INT_MISSING || x == GLXD_INT_VEND) ? ninf : __ddiv_rn((double)x, cfg.neg10_log10e);
            const double v = __dadd_rn(gll, prior);
            if (v > map_gll) {
                silver_gll = map_gll;
                silver_gt = map_gt;
                silver_i = map_i;
                silver_j = map_j;
                map_gll = v;
                map_gt = g;
                map_i = i;
                map_j = j;
            } else if (v > silver_gll) {
                silver_gll = v;
                silver_gt = g;
                silver_i = i;
                silver_j = j;
            }
        }
    }
    if (map_gt < 0 || silver_gt < 0) return 1;  // the reference asserts; the oracle fails the site
    if (homalt && map_gt == 0) {
        map_i = silver_i;
        map_j = silver_j;
    }
    g0 = gt_unphased(map_i);
    g1 = gt_unphased(map_j);
    const double q = round(__ddiv_rn(__dmul_rn(10.0, __dsub_rn(map_gll, silver_gll)), cfg.ln10));
    const int qi = to_int_x86(q);
    gq = qi < 99 ? qi : 99;
    revised = true;
    return 0;
}

// Who decides the rescoring of record r: its own GT (-1), or — multi-sample dataset — the dataset's samples in order, the first
// one with a reason (a called allele that is not unified; a homozygous-ALT call) deciding for all (genotyper.cc:97-115).
__device__ __forceinline__ int revise_forced(const DRecs&, const SiteCtx&, uint64_t, int, uint8_t, const int8_t*) { return -1; }
__device__ inline int revise_forced(const MRecs& R, const SiteCtx& st, uint64_t r, int na, uint8_t flags, const int8_t* map) {
    if (R.ns == 1) return -1;
    for (uint32_t k = 0; k < R.ns; k++) {
        int g0, g1;
        rec_gt_of(R, r, flags, k, g0, g1);
        const bool v0 = !gt_is_missing(g0) && g0 != GLXD_INT_VEND && gt_allele(g0) < na, v1 = !gt_is_missing(g1) && g1 != GLXD_INT_VEND && gt_allele(g1) < na;
        if (v0 && map[gt_allele(g0)] == -1 && !st.mono) return 1;
        if (v1 && map[gt_allele(g1)] == -1 && !st.mono) return 1;
        if (v0 && v1 && gt_allele(g0) > 0 && gt_allele(g0) == gt_allele(g1)) return 2;
    }
    return 0;
}
// the view's own GT pair of variant record r after revise_genotypes (when configured); 0 or an error code
template <class RT, class ST>
__device__ inline int rec_gt_revised(const DCfg& cfg, const RT& R, const ST& S, const SiteCtx& st, uint64_t r, int na, uint8_t flags, const int8_t* map, int& g0,
                                     int& g1, int& gq, bool& revised) {
    rec_gt(R, r, flags, g0, g1);
    revised = false;
    if (!cfg.revise) return 0;
    return revise_record(cfg, R, S, st, r, na, map, g0, g1, gq, revised, revise_forced(R, st, r, na, flags, map));
}
// Is variant record r "non-0/0" for its dataset: some sample's (revised) GT entry missing or not REF (genotyper.cc:380-392).
// With `first_only` the answer's GT pair is sample 0's, the one the half-call rules read (genotyper.cc:474-476).
template <class ST>
__device__ inline bool rec_non00_any(const DCfg& cfg, const MRecs& R, const ST& S, const SiteCtx& st, uint64_t r, int na, uint8_t flags, const int8_t* map,
                                     int& g0_first, int& g1_first) {
    const int forced = cfg.revise ? revise_forced(R, st, r, na, flags, map) : 0;
    bool non00 = false;
    MRecs Rk = R;
    for (uint32_t k = 0; k < R.ns; k++) {
        Rk.si = k;
        int g0, g1, gq;
        bool revised;
        rec_gt(Rk, r, flags, g0, g1);
        if (forced > 0) {
            int h0 = g0, h1 = g1;
            if (revise_record(cfg, Rk, S, st, r, na, map, h0, h1, gq, revised, forced) == 0) {  // (a failing sample fails the site from its own cell)
                g0 = h0;
                g1 = h1;
            }
        }
        if (k == 0) {
            g0_first = g0;
            g1_first = g1;
        }
        if (gt_is_missing(g0) || gt_allele(g0) != 0 || gt_is_missing(g1) || gt_allele(g1) != 0) non00 = true;
    }
    return non00;
}

// AlleleDepthHelper::Load + get (genotyper_utils.h:982-1114) for a single-sample record.
// Returns 0 ok / error code 2 (Invalid). depth out as the reference's `unsigned`.
template <class RT>
__device__ inline int depth_get(const DCfg& cfg, const RT& R, uint64_t r, int na, bool is_ref, int allele, unsigned& depth) {
    const int32_t* p = nullptr;
    depth = 0;
    if (cfg.xatlas) {
        if (col_get(R, cfg.col_ref_dp, r, p) != 1) return 2;
        const int32_t rr = p[0];
        int32_t vr = 0;
        if (!is_ref) {
            if (col_get(R, cfg.col_allele_dp, r, p) != 1) return 2;
            vr = p[0];
        }
        if (allele < 0 || allele >= na) return 0;
        if (allele == 0) depth = (unsigned)rr;
        else if (!is_ref) depth = (unsigned)vr;
        return 0;
    }
    if (is_ref) {
        if (col_get(R, cfg.col_ref_dp, r, p) != 1) return 2;
        if (allele == 0) depth = (unsigned)p[0];
        return 0;
    }
    const int nv = col_get(R, cfg.col_allele_dp, r, p);
    if (nv == -1 || nv == -3) return 0;  // zeros
    if (nv != na) return 2;
    if (allele >= 0 && allele < na) depth = (unsigned)p[allele];
    return 0;
}

struct CellScratch {
    uint8_t* cnt;      // this thread's count bytes: cnt[slot * stride]
    uint32_t stride;
};

// one value arriving at (field k, slot j): NumericFormatFieldHelper::add_record_data + combine_f, streaming
__device__ __forceinline__ void push_value(const DFld& f, int32_t* out, uint8_t* cnt, uint32_t stride, int j, int32_t v, bool squeeze_dp) {
    if (squeeze_dp && v > 2) {  // DPFieldHelper::squeeze, genotyper_utils.h:403-413
        int32_t d = 2;
        while ((long long)d * 2 <= v) d *= 2;
        v = d;
    }
    uint8_t& c = cnt[(size_t)j * stride];
    if (c == 0) {
        out[j] = v;
        c = 1;
        return;
    }
    c = 2;
    if (f.kind == GLX_FK_PL || f.combi == GLX_COMBI_MISSING) return;  // >1 value: decided at finalize
    if (f.type == GLX_TYPE_FLOAT) {
        const float a = __int_as_float(v), b = __int_as_float(out[j]);
        if (f.combi == GLX_COMBI_MIN ? (a < b) : (b < a)) out[j] = v;
    } else {
        if (f.combi == GLX_COMBI_MIN ? (v < out[j]) : (out[j] < v)) out[j] = v;
    }
}

// NumericFormatFieldHelper::add_record_data for one record (genotyper_utils.h:285-355, 87-135).
// rv: 0 found, -3 not found, >0 error code
template <class RT>
__device__ inline int numeric_add(const DFld& f, const RT& R, uint64_t r, int na, const int8_t* map, int count, const int16_t* cols, int ncols,
                                  int n_val_forced, int32_t* out, uint8_t* cnt, uint32_t stride, bool squeeze_dp, bool gq_revised, int gq_value,
                                  int col_gq) {
    int n_val = n_val_forced;
    if (n_val < 0) {
        n_val = count;
        if (f.number == GLX_NUM_ALT) n_val = na - 1;
        else if (f.number == GLX_NUM_ALLELES) n_val = na;
        else if (f.number == GLX_NUM_GENOTYPE) n_val = (int)n_genotypes(na > 2 ? na : 2);
    }
    for (int c = 0; c < ncols; c++) {
        const int col = cols[c];
        const int32_t* v = nullptr;
        int rv;
        int32_t gqv;
        if (gq_revised && col == col_gq) {  // bcf_update_format_int32(GQ) on the duplicated record, genotyper.cc:217
            gqv = gq_value;
            v = &gqv;
            rv = 1;
        } else
            rv = col_get(R, col, r, v);
        if (rv == -2) return 2;
        if (rv < 0) continue;
        if (rv != n_val) return 2;
        if (f.number == GLX_NUM_GENOTYPE) {
            const int ne = na > 2 ? na : 2;
            int g = 0;
            for (int j = 0; j < ne; j++)
                for (int i = 0; i <= j; i++, g++) {
                    int m1, m2;
                    if (na > 1) {
                        m1 = map[i];
                        m2 = map[j];
                    } else {
                        m1 = i <= 1 ? i : -1;
                        m2 = j <= 1 ? j : -1;
                    }
                    if (m1 < 0 || m2 < 0) continue;
                    const int o = alleles_gt(m1, m2);
                    if (o < count) push_value(f, out, cnt, stride, o, v[g], squeeze_dp);
                }
        } else if (f.number == GLX_NUM_ALT || f.number == GLX_NUM_ALLELES) {
            for (int j = 0; j < n_val; j++) {
                // NB the reference indexes allele_mapping with the value index for Number=ALT too (genotyper_utils.h:101)
                const int m = j < na ? map[j] : -1;
                if (m < 0 || m >= count) continue;  // m >= count: the reference writes out of bounds; dropped here
                push_value(f, out, cnt, stride, m, v[j], squeeze_dp);
            }
        } else {
            for (int j = 0; j < n_val; j++) push_value(f, out, cnt, stride, j, v[j], squeeze_dp);
        }
        return 0;
    }
    return -3;
}

// PLFieldHelper2::add_record_data (genotyper_utils.h:555-651). returns error code or 0; sets self_censor.
template <class RT>
__device__ inline int pl2_add(const DCfg& cfg, const RT& R, uint64_t r, int na, uint8_t flags, const int8_t* map, int A, int count, int32_t* out,
                              bool& self_censor) {
    const int32_t* buf = nullptr;
    const int rv = col_get(R, cfg.col_pl, r, buf);
    if (rv <= 0) return 0;
    const int n_val = (int)n_genotypes(na > 2 ? na : 2);
    if (rv != n_val) return 2;
    const bool has_sym = flags & GLX_RF_SYMBOLIC_LAST;
    bool all_mapped = true;
    int8_t rev[GLX_MAX_ALLELES];
    for (int a = 0; a < A; a++) rev[a] = -1;
    for (int i = 0; i < na; i++) {
        if (map[i] >= 0) rev[map[i]] = (int8_t)i;
        else if (i < na - 1 || !has_sym) all_mapped = false;
    }
    if (has_sym && all_mapped)
        for (int a = 0; a < A; a++)
            if (rev[a] == -1) rev[a] = (int8_t)(na - 1);
    int32_t v0 = out[0];
    if (v0 == 0) {
        if (buf[0] == 0) {
            int32_t margin = GLXD_INT_MISSING;
            for (int j = 1; j < n_val; j++)
                if (buf[j] != GLXD_INT_MISSING) margin = (margin != GLXD_INT_MISSING) ? min(margin, buf[j]) : buf[j];
            if (margin != GLXD_INT_MISSING) {
                int32_t old_margin = GLXD_INT_MISSING;
                for (int j = 1; j < count; j++)
                    if (out[j] != GLXD_INT_MISSING) old_margin = (old_margin != GLXD_INT_MISSING) ? min(old_margin, out[j]) : out[j];
                if (old_margin != GLXD_INT_MISSING && margin < old_margin) v0 = GLXD_INT_MISSING;
            }
        }
    } else if (v0 != GLXD_INT_MISSING) {
        self_censor = true;
    }
    if (v0 == GLXD_INT_MISSING) {
        for (int a = 0; a < A; a++) {
            const int ai = rev[a];
            for (int b = 0; b <= a; b++) {
                const int bi = rev[b];
                int32_t v = GLXD_INT_MISSING;
                if (ai >= 0 && bi >= 0) v = buf[alleles_gt(ai, bi)];
                out[alleles_gt(a, b)] = v;
            }
        }
    }
    return 0;
}

struct RecBasic {
    uint64_t r;
    int na;
    uint8_t flags;
    bool is_ref;
};

// FORMAT lifting of one record into every field that takes it (update_format_fields inner loop)
template <class RT>
__device__ inline int lift_record(const DCfg& cfg, const RT& R, const DSites& S, const DOut& O, const SiteCtx& st, uint32_t sample, uint64_t r,
                                  bool from_used_set, bool have_used, bool squeeze_mode, const int* slot_base, uint8_t* cnt, uint32_t stride,
                                  uint32_t& self_censor_mask) {
    const int na = R.n_allele[r];
    const uint8_t flags = R.flags[r];
    const bool is_ref = flags & GLX_RF_IS_REF;
    // does any field take this record in this role?
    bool any = false;
    for (uint32_t k = 0; k < cfg.n_fields; k++) {
        const DFld& f = cfg.f[k];
        if (squeeze_mode && !(f.name_flags & NF_DP)) continue;
        const bool wants_used = (cfg.allow_partial || f.ignore_non_variants) && have_used;
        if (wants_used == from_used_set) any = true;
    }
    if (!any) return 0;
    int8_t map[GLX_MAX_ALLELES];
    uint32_t del_mask;
    compute_allele_map(R, S, st, r, na, is_ref, map, del_mask);
    bool gq_revised = false;
    int gq_value = 0;
    if (!is_ref && cfg.revise) {
        int g0, g1;
        const int e = rec_gt_revised(cfg, R, S, st, r, na, flags, map, g0, g1, gq_value, gq_revised);
        if (e) return e;
    }
    for (uint32_t k = 0; k < cfg.n_fields; k++) {
        const DFld& f = cfg.f[k];
        if (squeeze_mode && !(f.name_flags & NF_DP)) continue;
        const bool wants_used = (cfg.allow_partial || f.ignore_non_variants) && have_used;
        if (wants_used != from_used_set) continue;
        const int count = S.fld_cnt[k][st.s];
        int32_t* out = O.fld[k] + S.fld_off[k][st.s] + (uint64_t)sample * count;
        uint8_t* c = cnt + (size_t)slot_base[k] * stride;
        int e = 0;
        switch (f.kind) {
            case GLX_FK_PL2: {
                bool sc = false;
                e = pl2_add(cfg, R, r, na, flags, map, st.A, count, out, sc);
                if (sc) self_censor_mask |= 1u << k;
                break;
            }
            case GLX_FK_FT: {
                const uint32_t m = R.filt[r];
                if (m) out[0] |= (int32_t)m;  // set union of FILTER ids (sorted dictionary bits)
                break;
            }
            case GLX_FK_STRING: e = 6; break;  // NotImplemented, genotyper_utils.h:713-716
            case GLX_FK_AD: {
                e = numeric_add(f, R, r, na, map, count, f.cols, f.n_cols, -1, out, c, stride, false, gq_revised, gq_value, cfg.col_gq);
                if (e == -3) {  // ADFieldHelper: fall back to ref_dp_format as a 1-value field (genotyper_utils.h:430-446)
                    const int16_t rc = cfg.col_ref_dp;
                    e = numeric_add(f, R, r, na, map, count, &rc, 1, 1, out, c, stride, false, gq_revised, gq_value, cfg.col_gq);
                }
                break;
            }
            default:
                e = numeric_add(f, R, r, na, map, count, f.cols, f.n_cols, -1, out, c, stride, squeeze_mode && f.kind == GLX_FK_DP, gq_revised, gq_value,
                                cfg.col_gq);
        }
        if (e > 0) return e;
    }
    return 0;
}

// The whole cell. `first` = index of the sample's first record with maxend > site.qbeg; `hi` = end of the sample's records.
// RT = DRecs: the sample's own single-sample dataset. RT = MRecs: output sample `sample` is sample R.si of a dataset of R.ns
// samples whose records are [first.., hi); the decisions the reference takes per dataset across its samples are the MULTI branches;
// errors are keyed by `err_cell` (site-major, datasets in order within the site: the first failing dataset's status wins).
template <class RT>
__device__ inline void genotype_cell_general(const DCfg& cfg, const RT& R, const DSites& S, const DOut& O, uint32_t s, uint32_t sample, uint64_t first,
                                             uint64_t hi, uint8_t* cnt, uint32_t stride, unsigned long long err_cell = ~0ull) {
    constexpr bool MULTI = std::is_same<RT, MRecs>::value;
    const SiteCtx st = load_site(S, s);
    const uint32_t N = S.n_samples;
    if (err_cell == ~0ull) err_cell = (unsigned long long)s * N + sample;

    // ---- pass 0: kept records, span test (types.h:174-215 on a beg-sorted stream) ----
    int n_kept = 0;
    bool spanned = false, have_cur = false;
    int cur_b = 0, cur_e = 0;
    for (uint64_t r = first; r < hi && R.beg[r] < st.qend; r++) {
        if (!(R.end[r] > st.qbeg) || !record_kept(R, S, st, r)) continue;
        n_kept++;
        const int b = R.beg[r], e = R.end[r];
        if (!have_cur) {
            cur_b = b;
            cur_e = e;
            have_cur = true;
        } else if (b <= cur_e) {
            cur_e = max(cur_e, e);
        } else {
            if (cur_b <= st.beg && cur_e >= st.end) spanned = true;
            cur_b = b;
            cur_e = e;
        }
    }
    if (have_cur && cur_b <= st.beg && cur_e >= st.end) spanned = true;

    int rnc0 = RNC_M, rnc1 = RNC_M;
    int al0 = 0, al1 = 0;  // raw GT ints, 0 = missing
    bool hc0 = false, hc1 = false;
    bool all_records_nonempty = false;
    int n_var = 0;
    uint64_t used1 = ~0ull, used2 = ~0ull;
    int err = 0;

    if (n_kept == 0) {
        // MissingData
    } else if (!cfg.allow_partial && !spanned) {
        rnc0 = rnc1 = RNC_P;
    } else {
        all_records_nonempty = true;
        // ---- pass 1: preprocess/revise every kept record, gather the calling state ----
        int mrd = -1;
        bool ref_noncalled = false;
        int n_non00 = 0, n_00 = 0;
        int is_b = 0, is_e = 0;
        uint64_t r_single = ~0ull;
        bool hc_abort = false;
        int n_usable = 0;
        float b1q = 0, b2q = 0;
        int b1i = 0, b2i = 0, b1w = 0, b2w = 0;
        uint64_t b1r = ~0ull, b2r = ~0ull;
        uint64_t mono_rec = ~0ull;
        bool mono_multi = false;
        int deferred_err = 0;
        for (uint64_t r = first; r < hi && R.beg[r] < st.qend && !err; r++) {
            if (!(R.end[r] > st.qbeg) || !record_kept(R, S, st, r)) continue;
            const uint8_t flags = R.flags[r];
            const int na = R.n_allele[r];
            const bool is_ref = flags & GLX_RF_IS_REF;
            if (flags & GLX_RF_NO_GT) {
                err = 1;
                break;
            }
            int g0, g1;
            rec_gt(R, r, flags, g0, g1);
            // GT entries index the record's alleles (ingest-time invariant, src/BCF_utils.cc:126-140); the reference
            // would throw from vector::at on anything else
            if ((!gt_is_missing(g0) && (g0 == GLXD_INT_VEND || gt_allele(g0) >= na)) || (!gt_is_missing(g1) && (g1 == GLXD_INT_VEND || gt_allele(g1) >= na))) {
                err = 2;
                break;
            }
            if (is_ref) {
                unsigned d;
                // update_min_ref_depth runs after every record was preprocessed/revised (genotyper.cc:312-330): its error
                // only counts if none of those failed
                const int de = depth_get(cfg, R, r, na, true, 0, d);
                if (de) {
                    if (!deferred_err) deferred_err = de;
                    continue;
                }
                mrd = (mrd < 0) ? (int)d : min(mrd, (int)d);
                if (!(flags & GLX_RF_HAPLOID) && (gt_is_missing(g0) || gt_allele(g0) != 0 || gt_is_missing(g1) || gt_allele(g1) != 0)) ref_noncalled = true;
                if constexpr (MULTI) {  // any sample of the dataset, genotyper.cc:335-347
                    for (uint32_t k = 0; k < R.ns && !ref_noncalled; k++) {
                        int h0, h1;
                        rec_gt_of(R, r, flags, k, h0, h1);
                        if (!(flags & GLX_RF_HAPLOID) && (gt_is_missing(h0) || gt_allele(h0) != 0 || gt_is_missing(h1) || gt_allele(h1) != 0)) ref_noncalled = true;
                    }
                }
                continue;
            }
            n_var++;
            int8_t map[GLX_MAX_ALLELES];
            uint32_t del_mask;
            compute_allele_map(R, S, st, r, na, false, map, del_mask);
            if (cfg.revise) {
                int gq;
                bool revised;
                err = revise_record(cfg, R, S, st, r, na, map, g0, g1, gq, revised, revise_forced(R, st, r, na, flags, map));
                if (err) break;
            }
            for (int a = 1; a < na; a++)
                if (map[a] == 1) {
                    if (mono_rec == ~0ull) mono_rec = r;
                    else mono_multi = true;
                }
            bool non00 = gt_is_missing(g0) || gt_allele(g0) != 0 || gt_is_missing(g1) || gt_allele(g1) != 0;
            if constexpr (MULTI) {
                // the record is sorted by the whole dataset's calls, and the half-call rules below read sample 0's pair
                if (R.ns > 1) non00 = rec_non00_any(cfg, R, S, st, r, na, flags, map, g0, g1);
            }
            if (!non00) {
                n_00++;
                continue;
            }
            const int rb = R.beg[r], re = R.end[r];
            if (n_non00 == 0) {
                is_b = rb;
                is_e = re;
                r_single = r;
            } else {
                is_b = max(is_b, rb);
                is_e = min(is_e, re);
            }
            n_non00++;
            if (!hc_abort) {
                const int a0 = gt_is_missing(g0) ? -1 : gt_allele(g0), a1 = gt_is_missing(g1) ? -1 : gt_allele(g1);
                if (a0 > 0 && a1 > 0) {
                    hc_abort = true;
                    n_usable = 0;
                } else {
                    const float nq = -1.0f * R.qual[r];
                    for (int w = 0; w < 2; w++) {
                        const int a = w ? a1 : a0;
                        if (!(a > 0 && a < na && map[a] > 0)) continue;
                        const int idx = n_usable++;
                        // keep the two smallest (−QUAL, insertion index) tuples (std::sort of genotyper.cc:497)
                        const bool lt1 = b1r == ~0ull || nq < b1q || (!(b1q < nq) && idx < b1i);
                        if (lt1) {
                            b2q = b1q; b2i = b1i; b2w = b1w; b2r = b1r;
                            b1q = nq; b1i = idx; b1w = w; b1r = r;
                        } else {
                            const bool lt2 = b2r == ~0ull || nq < b2q || (!(b2q < nq) && idx < b2i);
                            if (lt2) {
                                b2q = nq; b2i = idx; b2w = w; b2r = r;
                            }
                        }
                    }
                }
            }
        }

        if (!err) err = deferred_err;
        if (!err) {
            if ((n_var == 0 || !cfg.allow_partial) && ref_noncalled) {
                rnc0 = rnc1 = RNC_I;
            } else if (!st.mono) {
                // ---- translate_genotypes ----
                if (n_00 > 0) {  // 0/0 variant records lower min_ref_depth with AD[0] (genotyper.cc:396-397)
                    for (uint64_t r = first; r < hi && R.beg[r] < st.qend && !err; r++) {
                        if (!(R.end[r] > st.qbeg) || (R.flags[r] & GLX_RF_IS_REF) || !record_kept(R, S, st, r)) continue;
                        const uint8_t flags = R.flags[r];
                        const int na = R.n_allele[r];
                        int g0, g1;
                        rec_gt(R, r, flags, g0, g1);
                        bool multi_ds = false;
                        if constexpr (MULTI) multi_ds = R.ns > 1;
                        if (cfg.revise || multi_ds) {
                            int8_t map[GLX_MAX_ALLELES];
                            uint32_t del_mask;
                            compute_allele_map(R, S, st, r, na, false, map, del_mask);
                            if (cfg.revise) {
                                int gq;
                                bool revised;
                                err = revise_record(cfg, R, S, st, r, na, map, g0, g1, gq, revised, revise_forced(R, st, r, na, flags, map));
                                if (err) break;
                            }
                            if constexpr (MULTI) {
                                if (multi_ds) {
                                    int f0, f1;
                                    if (rec_non00_any(cfg, R, S, st, r, na, flags, map, f0, f1)) continue;
                                }
                            }
                        }
                        if (gt_is_missing(g0) || gt_allele(g0) != 0 || gt_is_missing(g1) || gt_allele(g1) != 0) continue;
                        unsigned d;
                        err = depth_get(cfg, R, r, na, false, 0, d);
                        if (err) break;
                        mrd = (mrd < 0) ? (int)d : min(mrd, (int)d);
                    }
                }
                int mode1 = -1, mode2 = -1;
                if (err) {
                } else if (n_non00 == 0) {
                    // `rd >= cfg.required_dp` is int vs size_t (genotyper.cc:417): rd converts to unsigned 64-bit
                    if ((unsigned long long)(long long)mrd >= (unsigned long long)cfg.required_dp) {
                        al0 = al1 = gt_unphased(0);
                        rnc0 = rnc1 = RNC_NA;
                    } else
                        rnc0 = rnc1 = RNC_D;
                } else if (n_non00 == 1) {
                    used1 = r_single;
                    mode1 = 2;
                } else if (is_b >= is_e) {
                    rnc0 = rnc1 = RNC_U;
                } else if (n_usable == 0) {
                    rnc0 = rnc1 = RNC_O;
                } else {
                    used1 = b1r;
                    mode1 = b1w;
                    if (n_usable == 2 || (cfg.top_two && n_usable > 2)) {
                        used2 = b2r;
                        mode2 = b2w;
                    }
                }
                if (mode1 >= 0) {
                    for (int pass = 0; pass < 2 && !err; pass++) {
                        // pass 0: record 1 (both slots in mode 2, else slot 0); pass 1: record 2 -> slot 1
                        const uint64_t r = pass ? used2 : used1;
                        if (pass == 1 && (mode1 == 2 || used2 == ~0ull)) {
                            if (mode1 != 2) rnc1 = RNC_O;
                            break;
                        }
                        const uint8_t flags = R.flags[r];
                        const int na = R.n_allele[r];
                        int8_t map[GLX_MAX_ALLELES];
                        uint32_t del_mask;
                        compute_allele_map(R, S, st, r, na, false, map, del_mask);
                        int g[2];
                        {
                            int gq;
                            bool revised;
                            err = rec_gt_revised(cfg, R, S, st, r, na, flags, map, g[0], g[1], gq, revised);
                            if (err) break;
                        }
                        unsigned dchk;
                        err = depth_get(cfg, R, r, na, false, 0, dchk);  // AlleleDepthHelper::Load errors
                        if (err) break;
                        const int mode = pass ? mode2 : mode1;
                        for (int k = 0; k < 2; k++) {
                            int in_ofs, out_ofs;
                            if (mode == 2) {
                                in_ofs = out_ofs = k;
                            } else {
                                if (k) break;
                                in_ofs = mode;
                                out_ofs = pass;
                            }
                            // fill_allele, genotyper.cc:531-555
                            int rnc = RNC_I, al = 0;
                            const int gv = g[in_ofs];
                            if (gv != GLXD_INT_VEND && !gt_is_missing(gv)) {
                                const int a = gt_allele(gv);
                                if (a >= na) {
                                    err = 2;
                                    break;
                                }
                                unsigned d;
                                depth_get(cfg, R, r, na, false, a, d);
                                if (d >= cfg.required_dp && (mrd < 0 || (unsigned)mrd >= cfg.required_dp)) {
                                    if (map[a] >= 0) {
                                        al = gt_unphased(map[a]);
                                        rnc = RNC_NA;
                                    } else
                                        rnc = (del_mask >> a & 1) ? RNC_LOSTDEL : RNC_LOST;
                                } else
                                    rnc = RNC_D;
                            }
                            if (out_ofs == 0) {
                                al0 = al;
                                rnc0 = rnc;
                                if (mode != 2) hc0 = true;
                            } else {
                                al1 = al;
                                rnc1 = rnc;
                                if (mode != 2) hc1 = true;
                            }
                        }
                    }
                }
            } else {
                // ---- translate_monoallelic ----
                if (mono_multi) {
                    rnc0 = rnc1 = RNC_O;
                    used1 = mono_rec;  // the first carrier is already in variant_records_used when the second one aborts (genotyper.cc:620-628)
                } else if (mono_rec == ~0ull) rnc0 = rnc1 = RNC_MONO;
                else {
                    used1 = mono_rec;
                    const uint64_t r = mono_rec;
                    const uint8_t flags = R.flags[r];
                    const int na = R.n_allele[r];
                    int8_t map[GLX_MAX_ALLELES];
                    uint32_t del_mask;
                    compute_allele_map(R, S, st, r, na, false, map, del_mask);
                    int g[2];
                    {
                        int gq;
                        bool revised;
                        err = rec_gt_revised(cfg, R, S, st, r, na, flags, map, g[0], g[1], gq, revised);
                    }
                    unsigned dchk;
                    if (!err) err = depth_get(cfg, R, r, na, false, 0, dchk);
                    for (int k = 0; k < 2 && !err; k++) {
                        int rnc = RNC_I, al = 0;
                        const int gv = g[k];
                        if (gv != GLXD_INT_VEND && !gt_is_missing(gv)) {
                            const int a = gt_allele(gv);
                            if (a >= na) {
                                err = 2;
                                break;
                            }
                            if (map[a] > 0) {
                                unsigned d;
                                depth_get(cfg, R, r, na, false, a, d);
                                if (d >= cfg.required_dp) {
                                    al = gt_unphased(map[a]);
                                    rnc = RNC_NA;
                                } else
                                    rnc = RNC_D;
                            } else
                                rnc = RNC_MONO;
                        }
                        if (k == 0) {
                            al0 = al;
                            rnc0 = rnc;
                        } else {
                            al1 = al;
                            rnc1 = rnc;
                        }
                    }
                }
            }
        }
    }

    // ---- FORMAT lifting (update_format_fields) ----
    int slot_base[GLX_MAX_FIELDS + 1];
    slot_base[0] = 0;
    for (uint32_t k = 0; k < cfg.n_fields; k++) slot_base[k + 1] = slot_base[k] + S.fld_cnt[k][s];
    for (int j = 0; j < slot_base[cfg.n_fields]; j++) cnt[(size_t)j * stride] = 0;
    for (uint32_t k = 0; k < cfg.n_fields; k++) {
        const DFld& f = cfg.f[k];
        const int count = S.fld_cnt[k][s];
        int32_t* out = O.fld[k] + S.fld_off[k][s] + (uint64_t)sample * count;
        if (f.kind == GLX_FK_PL2)
            for (int j = 0; j < count; j++) out[j] = GLXD_INT_MISSING;
        else if (f.kind == GLX_FK_FT || f.kind == GLX_FK_STRING)
            out[0] = 0;
    }
    const bool squeeze_mode = cfg.squeeze && n_var == 0 && all_records_nonempty;
    const bool have_used = used1 != ~0ull;
    uint32_t self_censor = 0;
    if (!err && all_records_nonempty) {
        for (uint64_t r = first; r < hi && R.beg[r] < st.qend && !err; r++) {
            if (!(R.end[r] > st.qbeg) || !record_kept(R, S, st, r)) continue;
            err = lift_record(cfg, R, S, O, st, sample, r, false, have_used, squeeze_mode, slot_base, cnt, stride, self_censor);
        }
        if (!err && have_used) err = lift_record(cfg, R, S, O, st, sample, used1, true, true, squeeze_mode, slot_base, cnt, stride, self_censor);
        if (!err && used2 != ~0ull) err = lift_record(cfg, R, S, O, st, sample, used2, true, true, squeeze_mode, slot_base, cnt, stride, self_censor);
    }
    if (err) {
        report_error(O, err_cell, err);
        return;
    }

    // ---- censoring decision (genotyper.cc:796-830) ----
    // per field: 0 = not censored, 1 = censor(half_call=false), 2 = censor(half_call=true)
    const bool half = st.mono || hc0 || hc1;
    int censor_all = 0, censor_but_dp_ft = 0, censor_but_dp_gq_ft = 0;
    if (squeeze_mode) {
        rnc0 = rnc1 = RNC_NA;
    } else if (rnc0 == RNC_M || rnc0 == RNC_P) {
        censor_all = 1;
    } else if (rnc0 == RNC_U || rnc1 == RNC_U || rnc0 == RNC_O || rnc1 == RNC_O) {
        censor_but_dp_ft = half ? 2 : 1;
    } else if (half) {
        censor_but_dp_gq_ft = 2;
    }

    // ---- finalize each field (combine_format_data, perform_censor, update_record_format) ----
    for (uint32_t k = 0; k < cfg.n_fields; k++) {
        const DFld& f = cfg.f[k];
        const int count = S.fld_cnt[k][s];
        int32_t* out = O.fld[k] + S.fld_off[k][s] + (uint64_t)sample * count;
        const uint8_t* c = cnt + (size_t)slot_base[k] * stride;
        int cz = 0;
        if (squeeze_mode) cz = (f.name_flags & NF_DP) ? 0 : 1;
        else if (censor_all) cz = censor_all;
        else if (censor_but_dp_ft && !(f.name_flags & (NF_DP | NF_FT))) cz = censor_but_dp_ft;
        else if (censor_but_dp_gq_ft && !(f.name_flags & (NF_DP | NF_GQ | NF_FT))) cz = censor_but_dp_gq_ft;

        if (f.kind == GLX_FK_PL2) {
            bool censor = cz != 0 || (self_censor >> k & 1);
            if (!censor) {
                bool zero = false, other = false;
                for (int j = 0; j < count; j++) {
                    const int32_t v = out[j];
                    if (v == 0) zero = true;
                    else if (v != GLXD_INT_MISSING) other = true;
                }
                censor = !(zero && other);
            }
            for (int j = 0; j < count; j++) {
                if (censor) out[j] = 0;
                else if (out[j] == GLXD_INT_MISSING) out[j] = 990;
            }
            continue;
        }
        if (f.kind == GLX_FK_FT || f.kind == GLX_FK_STRING) {
            const uint32_t uniq = (uint32_t)out[0];
            const int n_uniq = __popc(uniq);
            int32_t v = 0;
            if (n_uniq == 0) v = 0;
            else if (n_uniq == 1 || f.combi == GLX_COMBI_SEMICOLON) v = (int32_t)uniq;
            else if (f.combi == GLX_COMBI_MISSING) v = 0;
            else {
                report_error(O, err_cell, 2);
                return;
            }
            if (cz) v = 0;
            out[0] = v;
            continue;
        }
        const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
        if (f.kind == GLX_FK_PL) {  // PLFieldHelper::combine_format_data, genotyper_utils.h:480-515
            int zeroes = 0, nonzeroes = 0;
            bool multi = false;
            for (int j = 0; j < count; j++) {
                const uint8_t n = c[(size_t)j * stride];
                if (n == 1) {
                    if (out[j] == 0) zeroes++;
                    else nonzeroes++;
                } else if (n > 1)
                    multi = true;
            }
            const bool drop = multi || zeroes == 0 || (zeroes == 1 && nonzeroes == 0);
            for (int j = 0; j < count; j++)
                if (drop || c[(size_t)j * stride] != 1) out[j] = GLXD_INT_MISSING;
        } else {
            const int32_t dflt = f.default_zero ? 0 : missing;
            for (int j = 0; j < count; j++) {
                const uint8_t n = c[(size_t)j * stride];
                if (n == 0) out[j] = dflt;
                else if (n > 1 && f.combi == GLX_COMBI_MISSING) out[j] = missing;
            }
        }
        if (cz) {
            const int upto = (f.kind == GLX_FK_AD && cz == 2) ? 1 : count;  // ADFieldHelper::perform_censor
            for (int j = 0; j < upto; j++) out[j] = missing;
        }
        if (f.type == GLX_TYPE_INT && f.number != GLX_NUM_BASIC && count > 1) {
            bool all_missing = true;
            for (int j = 0; j < count; j++)
                if (out[j] != GLXD_INT_MISSING) all_missing = false;
            if (all_missing) out[1] = GLXD_INT_VEND;
        }
    }

    // ---- slot order (genotyper.cc:862-867), GT / RNC / site reductions ----
    if ((al0 != 0 && al1 == 0) || al0 > al1) {
        int t = al0; al0 = al1; al1 = t;
        t = rnc0; rnc0 = rnc1; rnc1 = t;
    }
    const int a0 = gt_is_missing(al0) ? -1 : gt_allele(al0), a1 = gt_is_missing(al1) ? -1 : gt_allele(al1);
    const size_t cell = ((size_t)s * N + sample) * 2;
    *reinterpret_cast<char2*>(O.gt + cell) = make_char2((signed char)a0, (signed char)a1);
    *reinterpret_cast<uchar2*>(O.rnc + cell) = make_uchar2((unsigned char)kRncChar[rnc0], (unsigned char)kRncChar[rnc1]);
    uint32_t used = 0;
    if (a0 >= 0) used |= 1u << a0;
    if (a1 >= 0) used |= 1u << a1;
    if (used) atomicOr(O.allele_used + s, used);
    // residuals flagging rule, genotyper.cc:837-849
    auto lost = [](int r) { return !(r == RNC_NA || r == RNC_M || r == RNC_P || r == RNC_D || r == RNC_MONO); };
    if (lost(rnc0) || lost(rnc1)) {
        // byte-wide atomicOr: site_flags is uint8; use a 32-bit CAS-free path on the aligned word
        uint32_t* w = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(O.site_flags + s) & ~(uintptr_t)3);
        atomicOr(w, 1u << (8 * (reinterpret_cast<uintptr_t>(O.site_flags + s) & 3)));
    }
}

// ---------------------------------------------------------------------------------------------------------
// Closed form for the commonest non-band cell: the query range holds exactly ONE record and it is a variant
// record whose alleles map to distinct unified alleles (site not monoallelic, squeeze off). Then all_records =
// variant_records = {r}, each output slot receives at most one value, nothing is censored, and the cell is one pass
// over the record with no count scratch. Same reference lines as genotype_cell_general. Returns false (nothing
// written) when the cell is not of this shape.
template <class RT, class ST>
__device__ inline bool genotype_cell_single(const DCfg& cfg, const RT& R, const ST& S, const DOut& O, uint32_t s, uint32_t sample, uint64_t r) {
    if (cfg.squeeze) return false;
    const SiteCtx st = load_site(S, s);
    if (st.mono) return false;
    const uint8_t flags = R.flags[r];
    if (flags & GLX_RF_IS_REF) return false;
    const int na = R.n_allele[r];
    if (na < 2) return false;  // REF-only record: numeric_add pads it to two alleles; left to the general path
    const int rbeg = R.beg[r], rend = R.end[r];
    int8_t map[GLX_MAX_ALLELES];
    uint32_t del_mask;
    compute_allele_map(R, S, st, r, na, false, map, del_mask);
    uint32_t seen = 1u;  // unified alleles already targeted (map[0] = 0)
    bool any_mapped = false;
    for (int i = 1; i < na; i++) {
        if (map[i] < 0) continue;
        any_mapped = true;
        if (seen >> map[i] & 1) return false;  // two record alleles on one unified allele: slots would hold >1 value
        seen |= 1u << map[i];
    }
    const uint32_t N = S.n_samples;
    int rnc0 = RNC_M, rnc1 = RNC_M, al0 = 0, al1 = 0;
    bool lift = false;
    int g[2] = {0, 0};
    int gq_value = 0;
    bool gq_revised = false;
    int err = 0;
    const bool kept = (rend > st.beg && rbeg < st.end) || any_mapped;
    if (!kept) {
        // MissingData
    } else if (!cfg.allow_partial && !(rbeg <= st.beg && rend >= st.end)) {
        rnc0 = rnc1 = RNC_P;
    } else {
        lift = true;
        if (flags & GLX_RF_NO_GT) err = 1;
        decode_gt(R.gt[r], flags, g[0], g[1]);
        if (!err && ((!gt_is_missing(g[0]) && (g[0] == GLXD_INT_VEND || gt_allele(g[0]) >= na)) || (!gt_is_missing(g[1]) && (g[1] == GLXD_INT_VEND || gt_allele(g[1]) >= na))))
            err = 2;
        if (!err && cfg.revise) err = revise_record(cfg, R, S, st, r, na, map, g[0], g[1], gq_value, gq_revised);
        unsigned d0 = 0;
        if (!err) err = depth_get(cfg, R, r, na, false, 0, d0);  // AlleleDepthHelper::Load
        if (!err) {
            const bool non00 = gt_is_missing(g[0]) || gt_allele(g[0]) != 0 || gt_is_missing(g[1]) || gt_allele(g[1]) != 0;
            if (!non00) {
                const int mrd = (int)d0;  // a 0/0 variant record only lowers min_ref_depth (genotyper.cc:396-426)
                if ((unsigned long long)(long long)mrd >= (unsigned long long)cfg.required_dp) {
                    al0 = al1 = gt_unphased(0);
                    rnc0 = rnc1 = RNC_NA;
                } else
                    rnc0 = rnc1 = RNC_D;
            } else {
                for (int k = 0; k < 2; k++) {  // fill_allele, min_ref_depth = -1 (no reference records)
                    int rnc = RNC_I, al = 0;
                    if (!gt_is_missing(g[k])) {
                        const int a = gt_allele(g[k]);
                        unsigned d;
                        depth_get(cfg, R, r, na, false, a, d);
                        if (d >= cfg.required_dp) {
                            if (map[a] >= 0) {
                                al = gt_unphased(map[a]);
                                rnc = RNC_NA;
                            } else
                                rnc = (del_mask >> a & 1) ? RNC_LOSTDEL : RNC_LOST;
                        } else
                            rnc = RNC_D;
                    }
                    if (k == 0) {
                        al0 = al;
                        rnc0 = rnc;
                    } else {
                        al1 = al;
                        rnc1 = rnc;
                    }
                }
            }
        }
    }
    // ---- fields ----
    for (uint32_t k = 0; k < cfg.n_fields && !err; k++) {
        const DFld& f = cfg.f[k];
        const int count = S.fld_cnt[k][s];
        int32_t* out = O.fld[k] + S.fld_off[k][s] + (uint64_t)sample * count;
        const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
        const bool vec_rule = f.type == GLX_TYPE_INT && f.number != GLX_NUM_BASIC && count > 1;
        if (!lift) {  // every field censored
            const int32_t c = (f.kind == GLX_FK_PL2 || f.kind == GLX_FK_FT || f.kind == GLX_FK_STRING) ? 0 : missing;
            for (int j = 0; j < count; j++) out[j] = c;
            if (vec_rule && c == GLXD_INT_MISSING) out[1] = GLXD_INT_VEND;
            continue;
        }
        if (f.kind == GLX_FK_STRING) {
            err = 6;
            break;
        }
        if (f.kind == GLX_FK_FT) {
            const uint32_t m = R.filt[r];
            const int nu = __popc(m);
            int32_t v = 0;
            if (nu == 0) v = 0;
            else if (nu == 1 || f.combi == GLX_COMBI_SEMICOLON) v = (int32_t)m;
            else if (f.combi == GLX_COMBI_MISSING) v = 0;
            else err = 2;
            out[0] = v;
            continue;
        }
        if (f.kind == GLX_FK_PL2) {
            for (int j = 0; j < count; j++) out[j] = GLXD_INT_MISSING;
            bool sc = false;
            err = pl2_add(cfg, R, r, na, flags, map, st.A, count, out, sc);
            if (err) break;
            bool zero = false, other = false;
            for (int j = 0; j < count; j++) {
                const int32_t v = out[j];
                if (v == 0) zero = true;
                else if (v != GLXD_INT_MISSING) other = true;
            }
            const bool censor = !(zero && other);
            for (int j = 0; j < count; j++) {
                if (censor) out[j] = 0;
                else if (out[j] == GLXD_INT_MISSING) out[j] = 990;
            }
            continue;
        }
        // numeric kinds: locate the source column (first orig_name present; AD falls back to the reference depth)
        int n_val = count;
        if (f.number == GLX_NUM_BASIC) n_val = f.count;
        else if (f.number == GLX_NUM_ALT) n_val = na - 1;
        else if (f.number == GLX_NUM_ALLELES) n_val = na;
        else n_val = (int)n_genotypes(na > 2 ? na : 2);
        const int32_t* v = nullptr;
        int32_t gqv;
        bool found = false;
        for (int ci = 0; ci < f.n_cols && !found && !err; ci++) {
            int rv;
            if (gq_revised && f.cols[ci] == cfg.col_gq) {
                gqv = gq_value;
                v = &gqv;
                rv = 1;
            } else
                rv = col_get(R, f.cols[ci], r, v);
            if (rv == -2) err = 2;
            else if (rv >= 0) {
                if (rv != n_val) err = 2;
                found = true;
            }
        }
        bool fallback = false;
        if (!found && !err && f.kind == GLX_FK_AD) {
            int rv;
            if (gq_revised && cfg.col_ref_dp == cfg.col_gq) {
                gqv = gq_value;
                v = &gqv;
                rv = 1;
            } else
                rv = col_get(R, cfg.col_ref_dp, r, v);
            if (rv == -2) err = 2;
            else if (rv >= 0) {
                if (rv != 1) err = 2;
                found = fallback = true;
                n_val = 1;
            }
        }
        if (err) break;
        const int32_t dflt = f.kind == GLX_FK_PL ? GLXD_INT_MISSING : (f.default_zero ? 0 : missing);
        for (int j = 0; j < count; j++) out[j] = dflt;
        int zeroes = 0, nonzeroes = 0;
        if (found) {
            if (f.number == GLX_NUM_GENOTYPE && !fallback) {
                int gi = 0;
                for (int j = 0; j < na; j++)
                    for (int i = 0; i <= j; i++, gi++) {
                        const int m1 = map[i], m2 = map[j];
                        if (m1 < 0 || m2 < 0) continue;
                        const int o = alleles_gt(m1, m2);
                        if (o >= count) continue;
                        out[o] = v[gi];
                        if (v[gi] == 0) zeroes++;
                        else nonzeroes++;
                    }
            } else if (f.number == GLX_NUM_ALT || f.number == GLX_NUM_ALLELES) {
                for (int j = 0; j < n_val; j++) {
                    const int m = j < na ? map[j] : -1;
                    if (m < 0 || m >= count) continue;
                    out[m] = v[j];
                }
            } else {
                for (int j = 0; j < n_val && j < count; j++) out[j] = v[j];
            }
        }
        if (f.kind == GLX_FK_PL && (zeroes == 0 || (zeroes == 1 && nonzeroes == 0)))
            for (int j = 0; j < count; j++) out[j] = GLXD_INT_MISSING;
        if (vec_rule) {
            bool all_missing = true;
            for (int j = 0; j < count; j++)
                if (out[j] != GLXD_INT_MISSING) all_missing = false;
            if (all_missing) out[1] = GLXD_INT_VEND;
        }
    }
    if (err) {
        report_error(O, (unsigned long long)s * N + sample, err);
        return true;
    }
    if ((al0 != 0 && al1 == 0) || al0 > al1) {
        int tmp = al0; al0 = al1; al1 = tmp;
        tmp = rnc0; rnc0 = rnc1; rnc1 = tmp;
    }
    const int a0 = gt_is_missing(al0) ? -1 : gt_allele(al0), a1 = gt_is_missing(al1) ? -1 : gt_allele(al1);
    const size_t cell = ((size_t)s * N + sample) * 2;
    *reinterpret_cast<char2*>(O.gt + cell) = make_char2((signed char)a0, (signed char)a1);
    *reinterpret_cast<uchar2*>(O.rnc + cell) = make_uchar2((unsigned char)kRncChar[rnc0], (unsigned char)kRncChar[rnc1]);
    uint32_t used = 0;
    if (a0 >= 0) used |= 1u << a0;
    if (a1 >= 0) used |= 1u << a1;
    if (used) atomicOr(O.allele_used + s, used);
    auto lost = [](int x) { return !(x == RNC_NA || x == RNC_M || x == RNC_P || x == RNC_D || x == RNC_MONO); };
    if (lost(rnc0) || lost(rnc1)) {
        uint32_t* w = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(O.site_flags + s) & ~(uintptr_t)3);
        atomicOr(w, 1u << (8 * (reinterpret_cast<uintptr_t>(O.site_flags + s) & 3)));
    }
    return true;
}

}  // namespace glxd
