// glx_cell_variant.cuh — the lone-variant-record cell with its loads hoisted.
//
// genotype_cell_single (glx_cell_general.cuh) is the closed form for a cell whose query range holds one variant
// record. Run against the columnar store it is a chain of ~90 dependent, scattered loads per warp (one per array
// touched, each behind the branch that needs it): the worklist kernel that runs it waits on memory latency
// (long-scoreboard stalls) at one third occupancy. This file states the same function so that
//   * every load is issued in one of three rounds of independent loads —
//       round 1  the site's packed rows (DSites::site_rows)
//       round 2  the record's view (four 16-byte loads of its blob, glx_cell_fast.cuh) and the first unification rows
//       round 3  the record's ALT allele keys, its PL and allele-depth vectors (parked in shared memory, one column
//                per thread, because they are indexed by data-dependent genotype numbers later)
//   * the code stays small (rolled loops over the actual allele counts): a first, fully unrolled version of this
//     function was bound by instruction-cache misses instead of by memory.
// Bounds: at most CAPA record alleles and site alleles (allele maps are nibble-packed, CAPA <= 8). Cells outside the
// bounds and records without a view are declined and go to genotype_cell_single; a cell declined half-way may
// already have stored some field vectors — every later stage rewrites the whole cell.
//
// Reference semantics restated (as in genotype_cell_single): src/genotyper.cc:35-80 (allele map), :95-225 (revise),
// :283-310 (kept / span), :363-592 with one variant record, src/genotyper_utils.h:285-355, 417-459, 472-517, 555-675.
#pragma once
#include "glx_cell_fast.cuh"

namespace glxd {

// nibble-packed allele maps: 0xF = none
__device__ __forceinline__ int nib_get(uint32_t w, int i) {
    const int v = (int)(w >> (4 * i)) & 15;
    return v == 15 ? -1 : v;
}
__device__ __forceinline__ uint32_t nib_set(uint32_t w, int i, int v) { return (w & ~(15u << (4 * i))) | ((uint32_t)(v & 15) << (4 * i)); }

template <int CAPA>
struct VariantSmem {
    static constexpr int G = CAPA * (CAPA + 1) / 2;
    static constexpr int SLOTS = G + CAPA;  // PL then allele depths
};

// `sm`: this thread's column of shared memory (element i at sm[i * sms]), VariantSmem<CAPA>::SLOTS elements.
template <int CAPA, class Sig = void>
__device__ inline bool genotype_cell_variant(const DCfg& cfg, const FastPlan& plan, const DRecs& R, const DSites& S, const DOut& O,
                                             const int32_t* __restrict__ blobs, uint32_t s, uint32_t sample, uint64_t first, int32_t* sm, uint32_t sms) {
    constexpr int CAPG = VariantSmem<CAPA>::G;
    static_assert(CAPA <= 8, "allele maps are nibble-packed");
    if (!plan.vview) return false;
    const uint32_t W = blob_words(plan.n_rv);
    // ---- round 1: site rows ----
    const int4 sr0 = __ldg(S.site_rows + 2 * (size_t)s), sr1 = __ldg(S.site_rows + 2 * (size_t)s + 1);
    // ---- round 2: the record's view, the site's first two unification rows ----
    const int4* bp = reinterpret_cast<const int4*>(blobs + first * W);
    const int4 b0 = __ldg(bp), b1 = __ldg(bp + 1), b2 = __ldg(bp + 2), b3 = __ldg(bp + 3);
    const int sbeg = sr0.x, send = sr0.y;
    const int A = sr1.x & 255;
    const uint32_t al0 = (uint32_t)sr1.y, u0 = (uint32_t)sr1.z, nU = (uint32_t)sr1.w - u0;
    const size_t ur1 = nU > 1 ? u0 + 1 : u0;  // a valid row either way (unif_rows carries spare ones)
    const int4 ua0 = __ldg(S.unif_rows + 2 * (size_t)u0), ub0 = __ldg(S.unif_rows + 2 * (size_t)u0 + 1);
    const int4 ua1 = __ldg(S.unif_rows + 2 * ur1), ub1 = __ldg(S.unif_rows + 2 * ur1 + 1);
    if ((sr1.x & 256) || A > CAPA) return false;
    if (!((uint32_t)b0.w & RB_VIEW)) return false;
    const int rbeg = b0.x, rend = b0.y;
    const uint32_t gtw = (uint32_t)b1.x, a0 = (uint32_t)b1.z;
    const uint32_t pk0 = (uint32_t)b1.w, pk1 = (uint32_t)b2.x, pk2 = (uint32_t)b2.y;
    const uint8_t flags = (uint8_t)(pk0 & 0x7f);
    const int na = (int)((pk0 >> 8) & 0xff);
    if (na < 2 || na > CAPA) return false;
    const int32_t clen[GLX_VIEW_COLS] = {(int32_t)view_len_of((pk0 >> 16) & 0xFFF), (int32_t)view_len_of(((pk0 >> 28) | (pk1 << 4)) & 0xFFF),
                                         (int32_t)view_len_of((pk1 >> 8) & 0xFFF),  (int32_t)view_len_of((pk1 >> 20) & 0xFFF),
                                         (int32_t)view_len_of(pk2 & 0xFFF),         (int32_t)view_len_of((pk2 >> 12) & 0xFFF)};
    const int32_t cval[GLX_VIEW_COLS] = {b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
    // column `col` of the record: htslib-style rv and its col_val; false if the column is not in the view
    auto column = [&](int col, int& rv, int32_t& val) -> bool {
        rv = -1;
        val = 0;
        if (col < 0) return true;
        const int j = plan.vc_slot[col];
        if (j < 0) return false;
        rv = view_rv((uint32_t)pick(clen, j));
        val = pick(cval, j);
        return true;
    };
    // ---- round 3: ALT allele keys (the first three in registers), PL and allele-depth vectors into shared memory ----
    uint32_t afl1, afl2, afl3, aln1, aln2, aln3;
    uint64_t apx1, apx2, apx3;
    {
        const size_t i1 = a0 + 1, i2 = na > 2 ? a0 + 2 : a0, i3 = na > 3 ? a0 + 3 : a0;
        afl1 = __ldg(R.al_flags + i1);
        aln1 = __ldg(R.al_len + i1);
        apx1 = __ldg(R.al_prefix + i1);
        afl2 = __ldg(R.al_flags + i2);
        aln2 = __ldg(R.al_len + i2);
        apx2 = __ldg(R.al_prefix + i2);
        afl3 = __ldg(R.al_flags + i3);
        aln3 = __ldg(R.al_len + i3);
        apx3 = __ldg(R.al_prefix + i3);
    }
    int pl_rv, ad_rv;
    int32_t pl_val, ad_val;
    if (!column(cfg.col_pl, pl_rv, pl_val) || !column(cfg.col_allele_dp, ad_rv, ad_val)) return false;
    if ((pl_rv > CAPG && pl_rv != 0x7fffffff) || (ad_rv > CAPA && ad_rv != 0x7fffffff)) return false;
    int32_t* const s_pl = sm;
    int32_t* const s_ad = sm + (size_t)CAPG * sms;
    {
        // chunks of independent loads first, then the stores: a load-store loop would serialise on memory latency
        const int n_pl = (pl_rv > 1 && pl_rv <= CAPG) ? pl_rv : 0, n_ad = (ad_rv > 1 && ad_rv <= CAPA) ? ad_rv : 0;
        const int32_t* psrc = n_pl ? R.pool + pl_val : blobs;
        const int32_t* asrc = n_ad ? R.pool + ad_val : blobs;
        constexpr int CH = 10;
        int32_t ta[CAPA];
#pragma unroll
        for (int i = 0; i < CAPA; i++) ta[i] = __ldg(asrc + (i < n_ad ? i : 0));
#pragma unroll
        for (int c0 = 0; c0 < CAPG; c0 += CH) {
            if (c0 < n_pl || c0 == 0) {
                int32_t tp[CH];
#pragma unroll
                for (int u = 0; u < CH; u++) tp[u] = __ldg(psrc + (c0 + u < n_pl ? c0 + u : 0));
#pragma unroll
                for (int u = 0; u < CH; u++)
                    if (c0 + u < CAPG && c0 + u < n_pl) s_pl[(size_t)(c0 + u) * sms] = tp[u];
            }
        }
#pragma unroll
        for (int i = 0; i < CAPA; i++)
            if (i < n_ad) s_ad[(size_t)i * sms] = ta[i];
        if (pl_rv == 1) s_pl[0] = pl_val;
        if (ad_rv == 1) s_ad[0] = ad_val;
    }

    // ---- allele map (preprocess_record): record allele -> unified allele, deletion flags ----
    uint32_t mapw = 0xFFFFFFF0u, invw = 0xFFFFFFF0u;  // map[0] = 0 = inv[0]
    uint32_t del_mask = 0;
    bool any_mapped = false;
    {
        uint32_t seen = 1u;
#pragma unroll 1
        for (int i = 1; i < na; i++) {
            uint32_t fl, ln;
            uint64_t px;
            if (i == 1) {
                fl = afl1; ln = aln1; px = apx1;
            } else if (i == 2) {
                fl = afl2; ln = aln2; px = apx2;
            } else if (i == 3) {
                fl = afl3; ln = aln3; px = apx3;
            } else {
                fl = R.al_flags[a0 + i]; ln = R.al_len[a0 + i]; px = R.al_prefix[a0 + i];
            }
            if (fl & GLX_AF_DELETION) del_mask |= 1u << i;
            if (!(fl & GLX_AF_IS_DNA)) continue;
            int m = -1;
#pragma unroll 1
            for (uint32_t j = 0; j < nU && m < 0; j++) {
                int4 ua, ub;
                if (j == 0) {
                    ua = ua0; ub = ub0;
                } else if (j == 1) {
                    ua = ua1; ub = ub1;
                } else {
                    ua = __ldg(S.unif_rows + 2 * (size_t)(u0 + j));
                    ub = __ldg(S.unif_rows + 2 * (size_t)(u0 + j) + 1);
                }
                const uint64_t upx = (uint64_t)(uint32_t)ub.x | (uint64_t)(uint32_t)ub.y << 32;
                if (ua.x != rbeg || ua.y != rend || (uint32_t)ua.z != ln || upx != px) continue;
                bool eq = true;
                if (ln > 8) {  // beyond the 8-byte prefix: compare the rest of the strings
                    const char* x = R.al_pool + R.al_str[a0 + i];
                    const char* y = S.unif_pool + (uint32_t)ub.z;
                    for (uint32_t c = 8; c < ln; c++)
                        if (x[c] != y[c]) {
                            eq = false;
                            break;
                        }
                }
                if (eq) m = ua.w & 255;
            }
            if (m < 0) continue;
            // two record alleles on one unified allele would put two values in a slot: not this closed form
            if (m >= CAPA || (seen >> m & 1)) return false;
            seen |= 1u << m;
            any_mapped = true;
            mapw = nib_set(mapw, i, m);
            invw = nib_set(invw, m, i);
        }
    }

    const uint32_t N = S.n_samples;
    int rnc0 = RNC_M, rnc1 = RNC_M, al_0 = 0, al_1 = 0;
    bool lift = false;
    int g0 = 0, g1 = 0;
    int gq_value = 0;
    bool gq_revised = false;
    int err = 0;
    const bool kept = (rend > sbeg && rbeg < send) || any_mapped;
    if (!kept) {
        // MissingData
    } else if (!cfg.allow_partial && !(rbeg <= sbeg && rend >= send)) {
        rnc0 = rnc1 = RNC_P;
    } else {
        lift = true;
        if (flags & GLX_RF_NO_GT) err = 1;
        decode_gt(gtw, flags, g0, g1);
        if (!err && ((!gt_is_missing(g0) && (g0 == GLXD_INT_VEND || gt_allele(g0) >= na)) || (!gt_is_missing(g1) && (g1 == GLXD_INT_VEND || gt_allele(g1) >= na))))
            err = 2;
        // ---- revise_genotypes ----
        if (!err && cfg.revise) {
            bool needs = false, homalt = false;
            if (!gt_is_missing(g0) && nib_get(mapw, gt_allele(g0)) == -1) needs = true;
            else if (!gt_is_missing(g1) && nib_get(mapw, gt_allele(g1)) == -1) needs = true;
            else if (!gt_is_missing(g0) && !gt_is_missing(g1) && gt_allele(g0) > 0 && gt_allele(g0) == gt_allele(g1)) needs = homalt = true;
            if (needs) {
                const int nGT = (int)n_genotypes(na);
                if (pl_rv != nGT) err = 1;  // Failure: couldn't find genotype likelihoods
                for (int k = 0; k < nGT && !err; k++) {
                    const int32_t x = s_pl[(size_t)k * sms];
                    if (x != GLXD_INT_MISSING && x != GLXD_INT_VEND && x < 0) err = 1;  // Invalid -> Failure (genotyper.cc:141-144)
                }
                int gq_rv;
                int32_t gq_val;
                if (!column(cfg.col_gq, gq_rv, gq_val)) return false;
                if (!err && gq_rv != 1) err = 1;
                if (!err) {
                    const double lost_prior = (double)S.lost_log_prior[s];
                    const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
                    double map_gll = ninf, silver_gll = ninf;
                    int map_i = -1, map_j = -1, map_gt = -1, silver_i = -1, silver_j = -1, silver_gt = -1;
                    int g = 0;
#pragma unroll 1
                    for (int j = 0; j < na; j++) {
                        const int mj = nib_get(mapw, j);
#pragma unroll 1
                        for (int i = 0; i <= j; i++, g++) {
                            double prior = 0.0;
                            const int mi = nib_get(mapw, i);
                            if (mi == -1 || mj == -1) prior = lost_prior;
                            else if (i > 0 && i == j) prior = (double)S.hom_log_prior[al0 + mi];
                            if (cfg.calibration && prior != 0.0) {
                                const bool has_indel = (mi > 0 && S.allele_is_indel[al0 + mi]) || (mj > 0 && S.allele_is_indel[al0 + mj]);
                                const double x = __dmul_rn((double)(has_indel ? cfg.indel_cal : cfg.snv_cal), prior);
                                prior = (x < 0.0) ? x : 0.0;
                            }
                            const int32_t x = s_pl[(size_t)g * sms];
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
                    if (map_gt < 0 || silver_gt < 0) err = 1;
                    else {
                        if (homalt && map_gt == 0) {
                            map_i = silver_i;
                            map_j = silver_j;
                        }
                        g0 = gt_unphased(map_i);
                        g1 = gt_unphased(map_j);
                        const double q = round(__ddiv_rn(__dmul_rn(10.0, __dsub_rn(map_gll, silver_gll)), cfg.ln10));
                        const int qi = to_int_x86(q);
                        gq_value = qi < 99 ? qi : 99;
                        gq_revised = true;
                    }
                }
            }
        }
        // ---- AlleleDepthHelper::Load / get on a variant record ----
        const bool ad_zero = ad_rv == -1 || ad_rv == -3;
        if (!err && !ad_zero && ad_rv != na) err = 2;
        auto depth = [&](int allele) -> unsigned { return (ad_zero || allele < 0 || allele >= na) ? 0u : (unsigned)s_ad[(size_t)allele * sms]; };
        if (!err) {
            const bool non00 = gt_is_missing(g0) || gt_allele(g0) != 0 || gt_is_missing(g1) || gt_allele(g1) != 0;
            if (!non00) {
                const int mrd = (int)depth(0);  // a 0/0 variant record only lowers min_ref_depth (genotyper.cc:396-426)
                if ((unsigned long long)(long long)mrd >= (unsigned long long)cfg.required_dp) {
                    al_0 = al_1 = gt_unphased(0);
                    rnc0 = rnc1 = RNC_NA;
                } else
                    rnc0 = rnc1 = RNC_D;
            } else {
#pragma unroll 1
                for (int k = 0; k < 2; k++) {  // fill_allele, min_ref_depth = -1 (no reference records)
                    const int g = k ? g1 : g0;
                    int rnc = RNC_I, al = 0;
                    if (!gt_is_missing(g)) {
                        const int a = gt_allele(g);
                        if (depth(a) >= cfg.required_dp) {
                            const int m = nib_get(mapw, a);
                            if (m >= 0) {
                                al = gt_unphased(m);
                                rnc = RNC_NA;
                            } else
                                rnc = (del_mask >> a & 1) ? RNC_LOSTDEL : RNC_LOST;
                        } else
                            rnc = RNC_D;
                    }
                    if (k == 0) {
                        al_0 = al;
                        rnc0 = rnc;
                    } else {
                        al_1 = al;
                        rnc1 = rnc;
                    }
                }
            }
        }
    }
    // ---- fields ----
    // One lifted field: `number` is the field's Number class — a compile-time constant when the config's signature is
    // passed (Sig), so that each field keeps only the code of its own shape. Returns false to decline the cell.
    // (The kernels pass Sig = void: unrolled per field the fused tiled kernel measured 4 % slower — larger code, more
    // spills — than with this one rolled, interpreted loop; the Sig form stays covered by the host-compiled tests.)
    const int GA = (int)n_genotypes(A);
    auto do_field = [&](uint32_t k, uint32_t number, int count, uint64_t off) -> bool {
        const DFld& f = cfg.f[k];
        int32_t* __restrict__ out = O.fld[k] + off + (uint64_t)sample * count;
        const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
        const bool vec_rule = f.type == GLX_TYPE_INT && number != GLX_NUM_BASIC && count > 1;
        if (!lift) {  // every field censored
            const int32_t c = f.kind == GLX_FK_PL2 ? 0 : missing;
            for (int j = 0; j < count; j++) out[j] = (j == 1 && vec_rule && c == GLXD_INT_MISSING) ? GLXD_INT_VEND : c;
            return true;
        }
        if (f.kind == GLX_FK_PL2) {
            if (count != GA) return false;
            int fill = -2;  // -2: the record has no PL, every slot stays missing
            if (pl_rv > 0) {
                if (pl_rv != (int)n_genotypes(na)) {
                    err = 2;
                    return true;
                }
                const bool has_sym = flags & GLX_RF_SYMBOLIC_LAST;
                bool all_mapped = true;
                for (int i = 0; i < na; i++)
                    if (nib_get(mapw, i) < 0 && (i < na - 1 || !has_sym)) all_mapped = false;
                fill = (has_sym && all_mapped) ? na - 1 : -1;
            }
            // unified genotype (a, b) takes the record's likelihood of (inv[a], inv[b]); unmapped alleles go through the symbolic one
            auto value = [&](int a, int b) -> int32_t {
                if (fill == -2) return GLXD_INT_MISSING;
                int ai = nib_get(invw, a), bi = nib_get(invw, b);
                if (ai < 0) ai = fill;
                if (bi < 0) bi = fill;
                return (ai >= 0 && bi >= 0) ? s_pl[(size_t)alleles_gt(ai, bi) * sms] : GLXD_INT_MISSING;
            };
            bool zero = false, other = false;
#pragma unroll 1
            for (int a = 0; a < A; a++)
#pragma unroll 1
                for (int b = 0; b <= a; b++) {
                    const int32_t x = value(a, b);
                    if (x == 0) zero = true;
                    else if (x != GLXD_INT_MISSING) other = true;
                }
            const bool censor = !(zero && other);
            int o = 0;
#pragma unroll 1
            for (int a = 0; a < A; a++)
#pragma unroll 1
                for (int b = 0; b <= a; b++, o++) {
                    const int32_t x = censor ? GLXD_INT_MISSING : value(a, b);
                    out[o] = censor ? 0 : (x == GLXD_INT_MISSING ? 990 : x);
                }
            return true;
        }
        // numeric kinds: the source column (first orig_name present; AD falls back to the reference depth)
        int n_val = count;
        if (number == GLX_NUM_BASIC) n_val = f.count;
        else if (number == GLX_NUM_ALT) n_val = na - 1;
        else if (number == GLX_NUM_ALLELES) n_val = na;
        else n_val = (int)n_genotypes(na);
        int src_col = -1;
        int32_t src_val = 0;
        bool found = false, from_gq = false, fallback = false;
        for (int ci = 0; ci < f.n_cols && !found && !err; ci++) {
            int rv;
            int32_t val;
            from_gq = false;
            if (gq_revised && f.cols[ci] == cfg.col_gq) {
                rv = 1;
                val = gq_value;
                from_gq = true;
            } else if (!column(f.cols[ci], rv, val))
                return false;
            if (rv == -2) err = 2;
            else if (rv >= 0) {
                if (rv != n_val) err = 2;
                found = true;
                src_col = f.cols[ci];
                src_val = val;
            }
        }
        if (!found && !err && f.kind == GLX_FK_AD) {
            int rv;
            int32_t val;
            from_gq = false;
            if (gq_revised && cfg.col_ref_dp == cfg.col_gq) {
                rv = 1;
                val = gq_value;
                from_gq = true;
            } else if (!column(cfg.col_ref_dp, rv, val))
                return false;
            if (rv == -2) err = 2;
            else if (rv >= 0) {
                if (rv != 1) err = 2;
                found = fallback = true;
                n_val = 1;
                src_col = cfg.col_ref_dp;
                src_val = val;
            }
        }
        if (err) return true;
        // where the source vector lives: 0 the scalar src_val, 1 shared PL, 2 shared allele depths, 3 the pool
        const int src_mode = n_val == 1 ? 0 : (!from_gq && src_col == cfg.col_pl) ? 1 : (!from_gq && src_col == cfg.col_allele_dp) ? 2 : 3;
        auto src = [&](int j) -> int32_t {
            if (src_mode == 0) return src_val;
            if (src_mode == 1) return s_pl[(size_t)j * sms];
            if (src_mode == 2) return s_ad[(size_t)j * sms];
            return __ldg(R.pool + src_val + j);
        };
        const int32_t dflt = f.kind == GLX_FK_PL ? GLXD_INT_MISSING : (f.default_zero ? 0 : missing);
        const int shape = (number == GLX_NUM_GENOTYPE && !fallback) ? 0 : (number == GLX_NUM_ALT || number == GLX_NUM_ALLELES) ? 1 : 2;
        // output slot o of the field: the source value that lands there, or the default
        // NB the reference indexes allele_mapping with the value index for Number=ALT too (genotyper_utils.h:101)
        auto value = [&](int o, int a, int b) -> int32_t {
            if (!found) return dflt;
            if (shape == 0) {
                if (a >= A) return dflt;
                const int ai = nib_get(invw, a), bi = nib_get(invw, b);
                return (ai >= 0 && bi >= 0) ? src(alleles_gt(ai, bi)) : dflt;
            }
            if (shape == 1) {
                const int ai = o < CAPA ? nib_get(invw, o) : -1;
                return (ai >= 0 && ai < n_val) ? src(ai) : dflt;
            }
            return o < n_val ? src(o) : dflt;
        };
        bool wipe = false, all_missing = false;
#ifndef GLX_ABLATE_GENOTYPE_SCATTER
        if (shape == 0 && found && count == GA) {
            // A genotype-numbered vector through the allele map, walked from the RECORD's side: its n_genotypes(na) entries (6 or 10
            // for a record with one or two ALT alleles) instead of the site's n_genotypes(A) slots (28 at seven alleles) twice over.
            // The allele map is injective here (checked above), so record genotype (i, j) with both alleles mapped is the one source of
            // slot (map[i], map[j]); every other slot holds the default.
            int zeroes = 0, nonzeroes = 0, k_mapped = 0;
            bool none = true;
            {
                int g = 0;
#pragma unroll 1
                for (int j = 0; j < na; j++) {
                    const int mj = nib_get(mapw, j);
                    if (mj >= 0 && mj < A) k_mapped++;
#pragma unroll 1
                    for (int i = 0; i <= j; i++, g++) {
                        const int mi = nib_get(mapw, i);
                        if (mi < 0 || mj < 0 || mi >= A || mj >= A) continue;
                        const int32_t x = src(g);
                        if (x != GLXD_INT_MISSING) none = false;
                        if (x == 0) zeroes++;
                        else nonzeroes++;
                    }
                }
            }
            if (dflt != GLXD_INT_MISSING && (int)n_genotypes((unsigned)k_mapped) < count) none = false;
            if (f.kind == GLX_FK_PL || vec_rule) all_missing = none;
            if (f.kind == GLX_FK_PL && (zeroes == 0 || (zeroes == 1 && nonzeroes == 0))) wipe = all_missing = true;
            const int32_t fillv = wipe ? GLXD_INT_MISSING : dflt;
            if ((count & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
                const int4 f4 = make_int4(fillv, fillv, fillv, fillv);
#pragma unroll 1
                for (int o = 0; o < count; o += 4) *reinterpret_cast<int4*>(out + o) = f4;
            } else {
#pragma unroll 1
                for (int o = 0; o < count; o++) out[o] = fillv;
            }
            if (!wipe) {
                int g = 0;
#pragma unroll 1
                for (int j = 0; j < na; j++) {
                    const int mj = nib_get(mapw, j);
#pragma unroll 1
                    for (int i = 0; i <= j; i++, g++) {
                        const int mi = nib_get(mapw, i);
                        if (mi < 0 || mj < 0 || mi >= A || mj >= A) continue;
                        out[alleles_gt(mi, mj)] = src(g);
                    }
                }
            }
            if (vec_rule && all_missing) out[1] = GLXD_INT_VEND;
            return true;
        }
        if (f.kind != GLX_FK_PL) {  // one pass: the only whole-vector rule is "all missing -> vector end in slot 1", patched afterwards
            bool none = true;
            int a = 0, b = 0;
#pragma unroll 1
            for (int o = 0; o < count; o++) {
                const int32_t x = value(o, a, b);
                if (x != GLXD_INT_MISSING) none = false;
                out[o] = x;
                if (++b > a) {
                    a++;
                    b = 0;
                }
            }
            if (vec_rule && none) out[1] = GLXD_INT_VEND;
            return true;
        }
#endif
        if (f.kind == GLX_FK_PL || vec_rule) {  // pass 1: the censoring rules look at the whole vector
            int zeroes = 0, nonzeroes = 0;
            all_missing = true;
            int a = 0, b = 0;
#pragma unroll 1
            for (int o = 0; o < count; o++) {
                const int32_t x = value(o, a, b);
                if (x != GLXD_INT_MISSING) all_missing = false;
                if (shape == 0 && found && a < A && nib_get(invw, a) >= 0 && nib_get(invw, b) >= 0) {
                    if (x == 0) zeroes++;
                    else nonzeroes++;
                }
                if (++b > a) {
                    a++;
                    b = 0;
                }
            }
            if (f.kind == GLX_FK_PL && (zeroes == 0 || (zeroes == 1 && nonzeroes == 0))) wipe = all_missing = true;
        }
        {
            int a = 0, b = 0;
#pragma unroll 1
            for (int o = 0; o < count; o++) {
                int32_t x = wipe ? GLXD_INT_MISSING : value(o, a, b);
                if (vec_rule && all_missing && o == 1) x = GLXD_INT_VEND;
                out[o] = x;
                if (++b > a) {
                    a++;
                    b = 0;
                }
            }
        }
            return true;
    };
    if constexpr (std::is_void<Sig>::value) {
        int cnt_next = cfg.n_fields ? __ldg(S.fld_cnt[0] + s) : 0;  // the next field's layout is fetched one field ahead
        uint64_t off_next = cfg.n_fields ? __ldg(S.fld_off[0] + s) : 0;
#pragma unroll 1
        for (uint32_t k = 0; k < cfg.n_fields && !err; k++) {
            const int count = cnt_next;
            const uint64_t off = off_next;
            if (k + 1 < cfg.n_fields) {
                cnt_next = __ldg(S.fld_cnt[k + 1] + s);
                off_next = __ldg(S.fld_off[k + 1] + s);
            }
            if (!do_field(k, cfg.f[k].number, count, off)) return false;
        }
    } else {
        int cnt[Sig::nf ? Sig::nf : 1];
        uint64_t off[Sig::nf ? Sig::nf : 1];
#pragma unroll
        for (uint32_t k = 0; k < Sig::nf; k++) {
            cnt[k] = __ldg(S.fld_cnt[k] + s);
            off[k] = __ldg(S.fld_off[k] + s);
        }
        bool declined = false;
        glx_static_for<Sig::nf>([&](auto kc) {
            constexpr uint32_t k = decltype(kc)::value;
            if (!declined && !err && !do_field(k, Sig::num(k), cnt[k], off[k])) declined = true;
        });
        if (declined) return false;
    }
    if (err) {
        report_error(O, (unsigned long long)s * N + sample, err);
        return true;
    }
    if ((al_0 != 0 && al_1 == 0) || al_0 > al_1) {
        int tmp = al_0; al_0 = al_1; al_1 = tmp;
        tmp = rnc0; rnc0 = rnc1; rnc1 = tmp;
    }
    const int o0 = gt_is_missing(al_0) ? -1 : gt_allele(al_0), o1 = gt_is_missing(al_1) ? -1 : gt_allele(al_1);
    const size_t cell = ((size_t)s * N + sample) * 2;
    *reinterpret_cast<char2*>(O.gt + cell) = make_char2((signed char)o0, (signed char)o1);
    *reinterpret_cast<uchar2*>(O.rnc + cell) = make_uchar2((unsigned char)kRncChar[rnc0], (unsigned char)kRncChar[rnc1]);
    uint32_t used = 0;
    if (o0 >= 0) used |= 1u << o0;
    if (o1 >= 0) used |= 1u << o1;
    if (used) atomicOr(O.allele_used + s, used);
    auto lostcall = [](int x) { return !(x == RNC_NA || x == RNC_M || x == RNC_P || x == RNC_D || x == RNC_MONO); };
    if (lostcall(rnc0) || lostcall(rnc1)) {
        uint32_t* w = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(O.site_flags + s) & ~(uintptr_t)3);
        atomicOr(w, 1u << (8 * (reinterpret_cast<uintptr_t>(O.site_flags + s) & 3)));
    }
    return true;
}

// The same closed form for the commonest such cell by far: a biallelic site and a record REF, ALT[, <symbolic>] whose ALT is the
// site's ALT and whose called alleles are REF / ALT, fields lifted (the record kept and spanning the site), nothing malformed.
// Everything above that is a loop over alleles or genotypes is written out for A = 2, na = 2 or 3 and the identity allele map
// (the symbolic allele unmapped), the field table is unrolled over the config's signature, PL / AD stay in registers (no
// shared-memory columns). Any other shape — and every error — is DECLINED (false):
// the caller hands the cell to genotype_cell_variant, which computes or reports it. As there, a declined cell may already
// have had some vectors stored; every later stage rewrites the whole cell. Supported fields: PL2, and numeric kinds numbered
// BASIC with one value or by ALLELES; a signature with anything else always declines.
template <class Sig>
__device__ inline bool genotype_cell_biallelic(const DCfg& cfg, const FastPlan& plan, const DRecs& R, const DSites& S, const DOut& O,
                                               const int32_t* __restrict__ blobs, uint32_t s, uint32_t sample, uint64_t first) {
    if (!plan.vview) return false;
    // ---- loads: site rows, the record's view, the first two unification rows ----
    const int4 sr0 = __ldg(S.site_rows + 2 * (size_t)s), sr1 = __ldg(S.site_rows + 2 * (size_t)s + 1);
    const int4* bp = reinterpret_cast<const int4*>(blobs + first * Sig::W);
    const int4 b0 = __ldg(bp), b1 = __ldg(bp + 1), b2 = __ldg(bp + 2), b3 = __ldg(bp + 3);
    const int sbeg = sr0.x, send = sr0.y;
    const uint32_t al0 = (uint32_t)sr1.y, u0 = (uint32_t)sr1.z, nU = (uint32_t)sr1.w - u0;
    const size_t ur1 = nU > 1 ? u0 + 1 : u0;
    const int4 ua0 = __ldg(S.unif_rows + 2 * (size_t)u0), ub0 = __ldg(S.unif_rows + 2 * (size_t)u0 + 1);
    const int4 ua1 = __ldg(S.unif_rows + 2 * ur1), ub1 = __ldg(S.unif_rows + 2 * ur1 + 1);
    if ((sr1.x & 256) || (sr1.x & 255) != 2) return false;
    if (!((uint32_t)b0.w & RB_VIEW)) return false;
    const int rbeg = b0.x, rend = b0.y;
    const uint32_t gtw = (uint32_t)b1.x, a0 = (uint32_t)b1.z;
    const uint32_t pk0 = (uint32_t)b1.w, pk1 = (uint32_t)b2.x, pk2 = (uint32_t)b2.y;
    const uint8_t flags = (uint8_t)(pk0 & 0x7f);
    const int na = (int)((pk0 >> 8) & 0xff);
    if (!(na == 2 || (na == 3 && (flags & GLX_RF_SYMBOLIC_LAST))) || (flags & GLX_RF_NO_GT)) return false;
    const int nGT = na == 2 ? 3 : 6;
    const uint32_t afl1 = __ldg(R.al_flags + a0 + 1), aln1 = __ldg(R.al_len + a0 + 1), afl2 = __ldg(R.al_flags + a0 + na - 1);
    const uint64_t apx1 = __ldg(R.al_prefix + a0 + 1);
    const int32_t clen[GLX_VIEW_COLS] = {(int32_t)view_len_of((pk0 >> 16) & 0xFFF), (int32_t)view_len_of(((pk0 >> 28) | (pk1 << 4)) & 0xFFF),
                                         (int32_t)view_len_of((pk1 >> 8) & 0xFFF),  (int32_t)view_len_of((pk1 >> 20) & 0xFFF),
                                         (int32_t)view_len_of(pk2 & 0xFFF),         (int32_t)view_len_of((pk2 >> 12) & 0xFFF)};
    const int32_t cval[GLX_VIEW_COLS] = {b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
    auto column = [&](int col, int& rv, int32_t& val) -> bool {
        rv = -1;
        val = 0;
        if (col < 0) return true;
        const int j = plan.vc_slot[col];
        if (j < 0) return false;
        rv = view_rv((uint32_t)pick(clen, j));
        val = pick(cval, j);
        return true;
    };
    int pl_rv, ad_rv;
    int32_t pl_val, ad_val;
    if (!column(cfg.col_pl, pl_rv, pl_val) || !column(cfg.col_allele_dp, ad_rv, ad_val)) return false;
    if (pl_rv > 0 && pl_rv != nGT) return false;  // a PL of another length: for the general form to judge
    const bool ad_zero = ad_rv == -1 || ad_rv == -3;
    if (!ad_zero && ad_rv != na) return false;    // "allele depth vector of the wrong length": general form reports it
    int32_t pl[6] = {0, 0, 0, 0, 0, 0}, ad[3] = {0, 0, 0};
    {
        const bool hp = pl_rv == nGT, ha = ad_rv == na;
        const int32_t* psrc = hp ? R.pool + pl_val : blobs;
        const int32_t* asrc = ha ? R.pool + ad_val : blobs;
        int32_t t[9];
#pragma unroll
        for (int i = 0; i < 6; i++) t[i] = __ldg(psrc + (i < nGT ? i : 0));
#pragma unroll
        for (int i = 0; i < 3; i++) t[6 + i] = __ldg(asrc + (i < na ? i : 0));
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (hp) pl[i] = t[i];
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (ha) ad[i] = t[6 + i];
    }
    // ---- the ALT allele must be the site's ALT (unified allele 1); a third allele must be the symbolic one (unmapped) ----
    if (!(afl1 & GLX_AF_IS_DNA)) return false;
    if (na == 3 && (afl2 & GLX_AF_IS_DNA)) return false;
    {
        int m = -1;
#pragma unroll 1
        for (uint32_t j = 0; j < nU && m < 0; j++) {
            int4 ua, ub;
            if (j == 0) {
                ua = ua0; ub = ub0;
            } else if (j == 1) {
                ua = ua1; ub = ub1;
            } else {
                ua = __ldg(S.unif_rows + 2 * (size_t)(u0 + j));
                ub = __ldg(S.unif_rows + 2 * (size_t)(u0 + j) + 1);
            }
            const uint64_t upx = (uint64_t)(uint32_t)ub.x | (uint64_t)(uint32_t)ub.y << 32;
            if (ua.x != rbeg || ua.y != rend || (uint32_t)ua.z != aln1 || upx != apx1) continue;
            bool eq = true;
            if (aln1 > 8) {
                const char* x = R.al_pool + R.al_str[a0 + 1];
                const char* y = S.unif_pool + (uint32_t)ub.z;
                for (uint32_t c = 8; c < aln1; c++)
                    if (x[c] != y[c]) {
                        eq = false;
                        break;
                    }
            }
            if (eq) m = ua.w & 255;
        }
        if (m != 1) return false;  // lost allele (or a unification the closed form does not take)
    }
    // kept (an allele maps); the fields are lifted only if the record spans the site
    if (!cfg.allow_partial && !(rbeg <= sbeg && rend >= send)) return false;
    int g0, g1;
    decode_gt(gtw, flags, g0, g1);
    // (a called allele >= 2 is malformed or the unmapped symbolic allele: lost-allele accounting and its rescoring are the general form's)
    if ((!gt_is_missing(g0) && (g0 == GLXD_INT_VEND || gt_allele(g0) >= 2)) || (!gt_is_missing(g1) && (g1 == GLXD_INT_VEND || gt_allele(g1) >= 2))) return false;
    // ---- revise_genotypes: the called alleles map, so only a homozygous-ALT call is rescored (three or six genotypes) ----
    int gq_value = 0;
    bool gq_revised = false;
    if (cfg.revise && !gt_is_missing(g0) && !gt_is_missing(g1) && gt_allele(g0) > 0 && gt_allele(g0) == gt_allele(g1)) {
        if (pl_rv != nGT) return false;
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (k < nGT && pl[k] != GLXD_INT_MISSING && pl[k] != GLXD_INT_VEND && pl[k] < 0) return false;
        int gq_rv;
        int32_t gq_val;
        if (!column(cfg.col_gq, gq_rv, gq_val) || gq_rv != 1) return false;
        const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
        double map_gll = ninf, silver_gll = ninf;
        int map_i = -1, map_j = -1, map_gt = -1, silver_i = -1, silver_j = -1, silver_gt = -1;
        const double lost_prior = (double)S.lost_log_prior[s];
#pragma unroll
        for (int g = 0; g < 6; g++) {  // (i, j) = (0,0), (0,1), (1,1), (0,2), (1,2), (2,2); allele 2 is the unmapped symbolic one
            if (g >= nGT) break;
            const int i = g == 2 || g == 4 ? 1 : (g == 5 ? 2 : 0), j = g == 0 ? 0 : (g < 3 ? 1 : 2);
            double prior = 0.0;
            if (j == 2) prior = lost_prior;
            else if (g == 2) prior = (double)S.hom_log_prior[al0 + 1];
            if (cfg.calibration && prior != 0.0) {
                const bool has_indel = (g == 2 || g == 4) && S.allele_is_indel[al0 + 1];  // a mapped non-reference allele that is an indel
                const double x = __dmul_rn((double)(has_indel ? cfg.indel_cal : cfg.snv_cal), prior);
                prior = (x < 0.0) ? x : 0.0;
            }
            const int32_t x = pl[g];
            const double gll = (x == GLXD_INT_MISSING || x == GLXD_INT_VEND) ? ninf : __ddiv_rn((double)x, cfg.neg10_log10e);
            const double v = __dadd_rn(gll, prior);
            if (v > map_gll) {
                silver_gll = map_gll; silver_gt = map_gt; silver_i = map_i; silver_j = map_j;
                map_gll = v; map_gt = g; map_i = i; map_j = j;
            } else if (v > silver_gll) {
                silver_gll = v; silver_gt = g; silver_i = i; silver_j = j;
            }
        }
        if (map_gt < 0 || silver_gt < 0) return false;
        if (map_gt == 0) {  // (homalt)
            map_i = silver_i;
            map_j = silver_j;
        }
        if (map_j >= 2) return false;  // rescored onto the unmapped symbolic allele: a lost call, the general form's
        g0 = gt_unphased(map_i);
        g1 = gt_unphased(map_j);
        const double q = round(__ddiv_rn(__dmul_rn(10.0, __dsub_rn(map_gll, silver_gll)), cfg.ln10));
        const int qi = to_int_x86(q);
        gq_value = qi < 99 ? qi : 99;
        gq_revised = true;
    }
    // ---- depth gate, allele slots ----
    auto depth = [&](int allele) -> unsigned { return (ad_zero || allele < 0 || allele >= na) ? 0u : (unsigned)(allele == 0 ? ad[0] : (allele == 1 ? ad[1] : ad[2])); };
    int rnc0, rnc1, al_0 = 0, al_1 = 0;
    if (!(gt_is_missing(g0) || gt_allele(g0) != 0 || gt_is_missing(g1) || gt_allele(g1) != 0)) {
        const int mrd = (int)depth(0);
        if ((unsigned long long)(long long)mrd >= (unsigned long long)cfg.required_dp) {
            al_0 = al_1 = gt_unphased(0);
            rnc0 = rnc1 = RNC_NA;
        } else
            rnc0 = rnc1 = RNC_D;
    } else {
        auto slot = [&](int g, int& al, int& rnc) {
            rnc = RNC_I;
            al = 0;
            if (!gt_is_missing(g)) {
                const int a = gt_allele(g);
                if (depth(a) >= cfg.required_dp) {
                    al = gt_unphased(a);  // identity map
                    rnc = RNC_NA;
                } else
                    rnc = RNC_D;
            }
        };
        slot(g0, al_0, rnc0);
        slot(g1, al_1, rnc1);
    }
    // ---- fields ----
    const uint32_t N = S.n_samples;
    bool ok = true;
    glx_static_for<Sig::nf>([&](auto kc) {
        constexpr uint32_t k = decltype(kc)::value;
        if (!ok) return;
        const DFld& f = cfg.f[k];
        const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
        if (f.kind == GLX_FK_PL2) {
            if (Sig::num(k) != GLX_NUM_GENOTYPE || __ldg(S.fld_cnt[k] + s) != 3) {
                ok = false;
                return;
            }
            int32_t* __restrict__ out = O.fld[k] + __ldg(S.fld_off[k] + s) + (uint64_t)sample * 3;
            const bool have = pl_rv > 0;  // (its length was checked above; the record's genotypes of REF / ALT are its first three)
            bool zero = false, other = false;
#pragma unroll
            for (int o = 0; o < 3; o++) {
                const int32_t x = have ? pl[o] : GLXD_INT_MISSING;
                if (x == 0) zero = true;
                else if (x != GLXD_INT_MISSING) other = true;
            }
            const bool censor = !(zero && other);
#pragma unroll
            for (int o = 0; o < 3; o++) {
                const int32_t x = have ? pl[o] : GLXD_INT_MISSING;
                out[o] = censor ? 0 : (x == GLXD_INT_MISSING ? 990 : x);
            }
            return;
        }
        if (f.kind == GLX_FK_PL || f.kind == GLX_FK_FT || f.kind == GLX_FK_STRING || (Sig::num(k) != GLX_NUM_BASIC && Sig::num(k) != GLX_NUM_ALLELES) ||
            (Sig::num(k) == GLX_NUM_BASIC && f.count != 1)) {
            ok = false;
            return;
        }
        // numeric kinds: the source column (first orig_name present; AD falls back to the reference depth)
        int n_val = Sig::num(k) == GLX_NUM_BASIC ? 1 : na;
        int src_col = -1;
        int32_t src_val = 0;
        bool found = false, fallback = false;
        for (int ci = 0; ci < f.n_cols && !found; ci++) {
            int rv;
            int32_t val;
            if (gq_revised && f.cols[ci] == cfg.col_gq) {
                rv = 1;
                val = gq_value;
            } else if (!column(f.cols[ci], rv, val)) {
                ok = false;
                return;
            }
            if (rv == -2 || (rv >= 0 && rv != n_val)) {
                ok = false;  // err = 2
                return;
            }
            if (rv >= 0) {
                found = true;
                src_col = f.cols[ci];
                src_val = val;
            }
        }
        if (!found && f.kind == GLX_FK_AD) {
            int rv;
            int32_t val;
            if (gq_revised && cfg.col_ref_dp == cfg.col_gq) {
                rv = 1;
                val = gq_value;
            } else if (!column(cfg.col_ref_dp, rv, val)) {
                ok = false;
                return;
            }
            if (rv == -2 || (rv >= 0 && rv != 1)) {
                ok = false;
                return;
            }
            if (rv >= 0) {
                found = fallback = true;
                n_val = 1;
                src_val = val;
            }
        }
        const int32_t dflt = f.default_zero ? 0 : missing;
        if (Sig::num(k) == GLX_NUM_BASIC) {
            O.fld[k][((uint64_t)s * N + sample)] = found ? src_val : dflt;
            return;
        }
        if (__ldg(S.fld_cnt[k] + s) != 2) {
            ok = false;
            return;
        }
        int32_t* __restrict__ out = O.fld[k] + __ldg(S.fld_off[k] + s) + (uint64_t)sample * 2;
        int32_t x0 = dflt, x1 = dflt;
        if (found) {
            if (n_val == 1) {
                x0 = src_val;  // the scalar (AD's reference-depth fallback: slot 0 only; a 1-value source under shape "alleles" cannot pass the length check)
                if (!fallback) x1 = dflt;
            } else if (src_col == cfg.col_pl) {
                x0 = pl[0];
                x1 = pl[1];
            } else if (src_col == cfg.col_allele_dp) {
                x0 = ad[0];
                x1 = ad[1];
            } else {
                x0 = __ldg(R.pool + src_val);
                x1 = __ldg(R.pool + src_val + 1);
            }
        }
        if (f.type == GLX_TYPE_INT && x0 == GLXD_INT_MISSING && x1 == GLXD_INT_MISSING) x1 = GLXD_INT_VEND;
        out[0] = x0;
        out[1] = x1;
    });
    if (!ok) return false;
    if ((al_0 != 0 && al_1 == 0) || al_0 > al_1) {
        int tmp = al_0; al_0 = al_1; al_1 = tmp;
        tmp = rnc0; rnc0 = rnc1; rnc1 = tmp;
    }
    const int o0 = gt_is_missing(al_0) ? -1 : gt_allele(al_0), o1 = gt_is_missing(al_1) ? -1 : gt_allele(al_1);
    const size_t cell = ((size_t)s * N + sample) * 2;
    *reinterpret_cast<char2*>(O.gt + cell) = make_char2((signed char)o0, (signed char)o1);
    *reinterpret_cast<uchar2*>(O.rnc + cell) = make_uchar2((unsigned char)kRncChar[rnc0], (unsigned char)kRncChar[rnc1]);
    uint32_t used = 0;
    if (o0 >= 0) used |= 1u << o0;
    if (o1 >= 0) used |= 1u << o1;
    if (used) atomicOr(O.allele_used + s, used);
    if (rnc0 == RNC_I || rnc1 == RNC_I) {  // the only "lost call" codes this form produces
        uint32_t* w = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(O.site_flags + s) & ~(uintptr_t)3);
        atomicOr(w, 1u << (8 * (reinterpret_cast<uintptr_t>(O.site_flags + s) & 3)));
    }
    return true;
}

}  // namespace glxd
