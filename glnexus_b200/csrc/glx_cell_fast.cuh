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
// The band's field values are resolved once per RECORD by a separate thread-per-record kernel into a 64-byte blob;
// a lane copies the blob into its cache when its cursor advances and re-reads nothing while consecutive sites fall
// in the same band.
#pragma once
#include <type_traits>
#include <utility>

#include "glx_cell_general.cuh"

namespace glxd {

#define GLX_FAST_RV 24      // resolved values cached per lane (shared memory column per lane)
#ifndef GLX_TILE_SITES
#define GLX_TILE_SITES 64   // sites walked by one CTA (at most 64: per-lane site masks are 64-bit)
#endif
#define GLX_MAX_CHUNKS 8      // a launch covers the tile grid in at most this many chunks (glx_launch)
#define GLX_TILE_THREADS 128
#define GLX_STAGE_MAX 32     // per-sample vectors of 3..32 ints are transposed through shared memory before being stored
#define GLX_WORK_SINGLE 0x80000000u  // worklist cursor flag: the query range holds exactly one record (not a monoallelic site)
#define GLX_WORK_DONE 0x40000000u    // set by the closed-form worklist kernel on an unflagged cell it finished (all-band cell)
#define GLX_WORK_CURSOR 0x3fffffffu
#define GLX_NWORK_MASK 0xffffu      // n_work[region]: low half = entries in the region, high half = entries the tiled kernel finished itself
#define GLX_VIEW_COLS 6     // source columns cached in a variant record's view (glx_cell_variant.cuh)

// How a field's per-sample vector is produced for a finished cell:
//   FM_VEC      BASIC-numbered field: `count` cached values
//   FM_PATTERN  allele/genotype-numbered field: slot 0 = x0, slot 1 = x1, first slot of every later genotype row
//               (0,a) = xr, everything else = xt  (4 cached values; covers [v0, default...] and the PL2 back-fill)
//   FM_CONST    nothing liftable from a lone band (PLFieldHelper): the censored form
enum { FM_VEC = 0, FM_PATTERN = 1, FM_CONST = 2 };

struct FastPlan {
    uint8_t enabled, n_rv;
    uint8_t mode[GLX_MAX_FIELDS], rv0[GLX_MAX_FIELDS], nrv[GLX_MAX_FIELDS];
    uint8_t vec_rule[GLX_MAX_FIELDS];  // int field, not BASIC-numbered: an all-missing vector is [missing, vector_end, ...]
    int32_t cen[GLX_MAX_FIELDS];       // value of a censored slot
    // record views (variant records): the source columns the genotyper reads, cached in the record's blob
    uint8_t vview, n_vc;
    int16_t vc[GLX_VIEW_COLS];        // cached columns
    int8_t vc_slot[GLX_MAX_COLS];     // column -> slot in vc, -1 not cached
    // declined list of the fused tiled kernel (set per launch, null otherwise): worklist entries its biallelic closed form
    // declined, as (region, index in the region); decl_n counts the entries offered, decl_cap bounds the list
    uint2* decl_list;
    uint32_t* decl_n;
    uint32_t decl_cap;
};

inline FastPlan make_fast_plan(const DCfg& c) {
    FastPlan p;
    memset(&p, 0, sizeof p);
    p.enabled = !c.squeeze && !c.has_string_field;
    int n = 0;
    for (uint32_t k = 0; k < c.n_fields; k++) {
        const DFld& f = c.f[k];
        int need = 0;
        const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
        p.cen[k] = missing;
        p.vec_rule[k] = f.type == GLX_TYPE_INT && f.number != GLX_NUM_BASIC;
        switch (f.kind) {
            case GLX_FK_NUMERIC:
            case GLX_FK_DP:
                if (f.number == GLX_NUM_BASIC) {
                    p.mode[k] = FM_VEC;
                    need = f.count;
                } else {
                    p.mode[k] = FM_PATTERN;
                    need = 4;
                }
                break;
            case GLX_FK_AD: p.mode[k] = FM_PATTERN; need = 4; break;
            case GLX_FK_PL: p.mode[k] = FM_CONST; need = 0; break;
            case GLX_FK_PL2: p.mode[k] = FM_PATTERN; need = 4; p.cen[k] = 0; p.vec_rule[k] = 0; break;
            case GLX_FK_FT: p.mode[k] = FM_VEC; need = 1; p.cen[k] = 0; p.vec_rule[k] = 0; break;
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
    // columns a lone variant record is asked for: the genotyper's own four and every field's sources
    for (int i = 0; i < GLX_MAX_COLS; i++) p.vc_slot[i] = -1;
    bool ok = p.enabled && !c.xatlas;
    auto want = [&](int col) {
        if (col < 0 || col >= GLX_MAX_COLS || p.vc_slot[col] >= 0) return;
        if (p.n_vc >= GLX_VIEW_COLS) {
            ok = false;
            return;
        }
        p.vc_slot[col] = (int8_t)p.n_vc;
        p.vc[p.n_vc++] = (int16_t)col;
    };
    want(c.col_pl);
    want(c.col_allele_dp);
    want(c.col_gq);
    want(c.col_ref_dp);
    for (uint32_t k = 0; k < c.n_fields; k++) {
        if (c.f[k].kind == GLX_FK_FT || c.f[k].kind == GLX_FK_STRING) ok = false;
        for (int ci = 0; ci < c.f[k].n_cols; ci++) want(c.f[k].cols[ci]);
    }
    p.vview = ok;
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

// ---- resolved-record blobs --------------------------------------------------------------------------------
// Everything a lane needs from a record to finish reference-band cells, resolved ONCE per record by
// glx_resolve_records_kernel (thread per record) instead of once per (site, sample) cell as the reference does:
//   [0] beg  [1] end  [2] running max end  [3] RB_* bits  [4] MIN_DP as min_ref_depth  [5] beg of the sample's next
//   record (INT_MAX after the last)  [6..6+n_rv) the cached field values of the FastPlan.  16-byte aligned rows.
//
// A record that is not a reference band (RB_SLOW) carries a VIEW of itself instead when the plan has one (RB_VIEW): the
// worklist kernel for lone variant records reads the whole record with four 16-byte loads instead of one scattered
// load per array of the columnar store:
//   [4] GT word  [6] allele_off  [7..9] flags (8 bits), n_allele (8), then GLX_VIEW_COLS 12-bit column lengths
//   [10..16) col_val of the cached columns (FastPlan::vc)
#define GLX_BLOB_HDR 6
#define GLX_BLOB_MIN 16  // words; room for the view
inline __host__ __device__ uint32_t blob_words(uint32_t n_rv) {
    const uint32_t w = (GLX_BLOB_HDR + n_rv + 3u) & ~3u;
    return w < GLX_BLOB_MIN ? GLX_BLOB_MIN : w;
}
// 12-bit code of a glx_records.col_len entry: lengths up to 0xFF0, the three sentinels, or 0xFFC = does not fit
__host__ __device__ inline uint32_t view_len_code(uint32_t len) {
    if (len >= 0xFFFDu) return 0xF00u | (len & 0xFFu);
    return len <= 0xFF0u ? len : 0xFFCu;
}
__host__ __device__ inline uint32_t view_len_of(uint32_t code) { return code >= 0xFFDu ? (0xF000u | code) : code; }

struct LaneCursor {
    uint64_t base;  // first record of the lane's sample
    uint32_t n;     // records of the sample
    uint32_t c;     // first record (relative) with maxend > current qbeg
    int beg_c, end_c, maxend_c, beg_n;
    uint32_t bits;  // RB_* of record c
    int mrd;
};

// load record c's blob: header into registers, field values into the lane's cache column
__device__ __forceinline__ void cursor_load(const int32_t* __restrict__ blobs, uint32_t W, uint32_t n_rv, LaneCursor& L, int32_t* rv, int rvs) {
    if (L.c < L.n) {
        const int4* b = reinterpret_cast<const int4*>(blobs + (L.base + L.c) * W);
        const int4 h = b[0];
        L.beg_c = h.x;
        L.end_c = h.y;
        L.maxend_c = h.z;
        L.bits = (uint32_t)h.w;
        const int4 g = b[1];
        L.mrd = g.x;
        L.beg_n = g.y;
        if (n_rv > 0) rv[0] = g.z;
        if (n_rv > 1) rv[(size_t)rvs] = g.w;
        for (uint32_t q = 2; q * 4 < W; q++) {
            const int4 v = b[q];
            const uint32_t j = q * 4 - GLX_BLOB_HDR;
            if (j < n_rv) rv[(size_t)j * rvs] = v.x;
            if (j + 1 < n_rv) rv[(size_t)(j + 1) * rvs] = v.y;
            if (j + 2 < n_rv) rv[(size_t)(j + 2) * rvs] = v.z;
            if (j + 3 < n_rv) rv[(size_t)(j + 3) * rvs] = v.w;
        }
    } else {
        L.beg_c = L.beg_n = L.maxend_c = L.end_c = 0x7fffffff;
        L.bits = 1;
        L.mrd = -1;
    }
}

__device__ __forceinline__ void cursor_init(const DRecs& R, const int32_t* __restrict__ blobs, uint32_t W, uint32_t n_rv, LaneCursor& L, uint32_t sample,
                                            uint32_t tile_cursor, int32_t* rv, int rvs) {
    L.base = R.ds_rec_off[sample];
    const uint64_t hi = R.ds_rec_off[sample + 1];
    L.n = (uint32_t)(hi - L.base);
    L.c = tile_cursor < L.n ? tile_cursor : L.n;  // from the tile cursor table; the first seek moves it on to the site's own qbeg
    cursor_load(blobs, W, n_rv, L, rv, rvs);
}

// `back`: this site's qbeg is smaller than the previous site's (query hulls are not monotone in general)
__device__ __forceinline__ void cursor_seek(const int32_t* __restrict__ blobs, uint32_t W, uint32_t n_rv, LaneCursor& L, int qbeg, bool back, int32_t* rv, int rvs) {
    if (back) {
        bool moved = false;
        while (L.c > 0 && blobs[(L.base + L.c - 1) * W + 2] > qbeg) {
            L.c--;
            moved = true;
        }
        if (moved) cursor_load(blobs, W, n_rv, L, rv, rvs);
    }
    while (L.c < L.n && L.maxend_c <= qbeg) {
        L.c++;
        cursor_load(blobs, W, n_rv, L, rv, rvs);
    }
}

enum { CT_SLOW = 0, CT_MISSING = 1, CT_PARTIAL = 2, CT_REF = 3 };
enum { RB_SLOW = 1, RB_NONCALLED = 2, RB_VIEW = 4, RB_FOUND0 = 256 };  // RB_FOUND0 << k: field k found a source column on this band

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
                if (found) bits |= (uint32_t)RB_FOUND0 << k;
                if (f.kind == GLX_FK_PL) break;  // a lone band never yields a liftable PL vector; validated only
                if (f.number == GLX_NUM_BASIC) {
                    for (int j = 0; j < f.count; j++) c[(size_t)j * rvs] = found ? v[j] : dflt;
                } else {
                    // allele map of a band is [0,-1]: only slot 0 receives a value; the rest take the default
                    const int32_t x0 = found ? v[0] : dflt;
                    c[0] = x0;
                    c[(size_t)rvs] = (plan.vec_rule[k] && x0 == GLXD_INT_MISSING && dflt == GLXD_INT_MISSING) ? GLXD_INT_VEND : dflt;
                    c[(size_t)2 * rvs] = dflt;
                    c[(size_t)3 * rvs] = dflt;
                }
                break;
            }
            case GLX_FK_PL2: {
                const int32_t* buf = nullptr;
                const int rvv = col_get(R, cfg.col_pl, r, buf);
                int32_t p0 = 0, p1 = 0, p2 = 0;
                if (rvv > 0) {
                    if (rvv != 3) return RB_SLOW;
                    bits |= (uint32_t)RB_FOUND0 << k;
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
                // back-fill through the symbolic allele: (0,0) -> PL[0], (0,a) -> PL[1], (a,b) -> PL[2]
                c[0] = p0;
                c[(size_t)rvs] = p1;
                c[(size_t)2 * rvs] = p1;
                c[(size_t)3 * rvs] = p2;
                break;
            }
            case GLX_FK_FT: {
                const uint32_t m = R.filt[r];
                if (m) bits |= (uint32_t)RB_FOUND0 << k;
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

// Tile cursor table: cursor[tile][sample] = first record (relative to the sample's first) whose running max end exceeds
// the tile's starting query position tq[tile] (tq = suffix-minimum of the tiles' first qbeg, so it is non-decreasing
// and never later than any qbeg in the tile). Record r is that record for exactly the tiles with
// maxend[r-1] <= tq[tile] < maxend[r], so the table is filled by the per-record kernel with two searches over the tiny
// tq array instead of one binary search over the sample's records per (lane, tile) in the tiled kernel.
// Entries no record claims stay 0xFFFFFFFF (= past the end).
__device__ inline void fill_tile_cursors(const DRecs& R, uint64_t r, uint64_t sample_base, uint32_t sample, uint32_t n_samples, const int32_t* __restrict__ tq,
                                         uint32_t n_tiles, uint32_t* __restrict__ cursors) {
    const int cur = R.maxend[r];
    uint32_t lo = 0, hi = n_tiles;  // t_hi = first tile with tq >= cur
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tq[mid] < cur) lo = mid + 1;
        else hi = mid;
    }
    const uint32_t t_hi = lo;
    uint32_t t_lo = 0;
    if (r > sample_base) {  // first tile with tq >= maxend[r-1]
        const int prev = R.maxend[r - 1];
        lo = 0;
        hi = t_hi;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (tq[mid] < prev) lo = mid + 1;
            else hi = mid;
        }
        t_lo = lo;
    }
    for (uint32_t tile = t_lo; tile < t_hi; tile++) cursors[(size_t)tile * n_samples + sample] = (uint32_t)(r - sample_base);
}

// f(integral_constant<uint32_t, 0>) ... f(integral_constant<uint32_t, N - 1>): a loop whose index is a constant expression
template <class F, uint32_t... Ks>
__host__ __device__ __forceinline__ void glx_static_for_impl(F&& f, std::integer_sequence<uint32_t, Ks...>) {
    (f(std::integral_constant<uint32_t, Ks>{}), ...);
}
template <uint32_t N, class F>
__host__ __device__ __forceinline__ void glx_static_for(F&& f) {
    glx_static_for_impl(f, std::make_integer_sequence<uint32_t, N>{});
}

// a[j] for a register array and a run-time j
template <int N>
__device__ __forceinline__ int32_t pick(const int32_t (&a)[N], int j) {
    int32_t r = a[0];
#pragma unroll
    for (int i = 1; i < N; i++) r = (j == i) ? a[i] : r;
    return r;
}

// htslib-style return value of a column length (col_get's contract)
__device__ __forceinline__ int view_rv(uint32_t len) {
    if (len == GLX_LEN_ABSENT) return -3;
    if (len == 0xFFFE) return -2;
    if (len == 0xFFFD) return 0x7fffffff;
    return (int)len;
}

// build_record_blob for a plan with record views (FastPlan::vview: every column the genotyper can ask for is one of
// the plan's n_vc <= GLX_VIEW_COLS view columns). Same result as the general form below, but all of the record's
// scalars and its (length, value) pairs of the view columns are fetched up front as independent loads, and both the
// reference-band resolution and the variant view are computed from those registers; only a band's PL triple and
// multi-value BASIC fields need a second round of loads.
__device__ inline void build_record_blob_cached(const DCfg& cfg, const FastPlan& plan, const DRecs& R, uint64_t r, bool last_of_sample, int32_t* blob) {
    const uint32_t W = blob_words(plan.n_rv);
    const uint8_t flags = R.flags[r];
    const int na = R.n_allele[r];
    const uint32_t gtw = R.gt[r];
    const int rbeg = R.beg[r], rend = R.end[r], rmaxend = R.maxend[r];
    const int beg_next = last_of_sample ? 0x7fffffff : R.beg[r + 1];
    const uint32_t a_off = R.allele_off[r];
    int32_t clen[GLX_VIEW_COLS], cval[GLX_VIEW_COLS];
#pragma unroll
    for (int j = 0; j < GLX_VIEW_COLS; j++) {
        const bool on = j < plan.n_vc;
        const size_t idx = (on ? (size_t)plan.vc[j] * R.n_records : 0) + r;
        const uint16_t l = R.col_len[idx];
        const int32_t v = R.col_val[idx];
        clen[j] = on ? (int32_t)l : (int32_t)GLX_LEN_ABSENT;
        cval[j] = on ? v : 0;
    }
    auto column = [&](int col, int& rv, int32_t& val) {
        rv = -1;
        val = 0;
        if (col < 0) return;
        const int j = plan.vc_slot[col];
        rv = view_rv((uint32_t)pick(clen, j));
        val = pick(cval, j);
    };
    for (uint32_t j = GLX_BLOB_HDR; j < W; j++) blob[j] = 0;
    int mrd = -1;
    uint32_t bits = RB_SLOW;
    // ---- a well-formed reference band? (resolve_ref_band) ----
    do {
        if (!(flags & GLX_RF_IS_REF) || na != 2 || (flags & GLX_RF_NO_GT)) break;
        int g0, g1;
        decode_gt(gtw, flags, g0, g1);
        if ((!gt_is_missing(g0) && (g0 == GLXD_INT_VEND || gt_allele(g0) >= na)) || (!gt_is_missing(g1) && (g1 == GLXD_INT_VEND || gt_allele(g1) >= na))) break;
        uint32_t b = 0;
        if (!(flags & GLX_RF_HAPLOID) && (gt_is_missing(g0) || gt_allele(g0) != 0 || gt_is_missing(g1) || gt_allele(g1) != 0)) b |= RB_NONCALLED;
        int ref_rv;
        int32_t ref_val;
        column(cfg.col_ref_dp, ref_rv, ref_val);
        if (ref_rv != 1) break;  // "reference depth FORMAT field is missing or malformed"
        bool ok = true;
        int32_t* rvp = blob + GLX_BLOB_HDR;
        for (uint32_t k = 0; k < cfg.n_fields && ok; k++) {
            const DFld& f = cfg.f[k];
            int32_t* c = rvp + plan.rv0[k];
            const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
            const int32_t dflt = f.default_zero ? 0 : missing;
            if (f.kind == GLX_FK_PL2) {
                int rvv;
                int32_t val;
                column(cfg.col_pl, rvv, val);
                int32_t p0 = 0, p1 = 0, p2 = 0;
                if (rvv > 0) {
                    if (rvv != 3) {
                        ok = false;
                        break;
                    }
                    b |= (uint32_t)RB_FOUND0 << k;
                    const int32_t* buf = R.pool + val;
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
                c[1] = p1;
                c[2] = p1;
                c[3] = p2;
                continue;
            }
            // numeric kinds (plans with a view have no FT / string field)
            int n_val = f.count;
            if (f.number == GLX_NUM_ALT) n_val = 1;
            else if (f.number == GLX_NUM_ALLELES) n_val = 2;
            else if (f.number == GLX_NUM_GENOTYPE) n_val = 3;
            bool found = false;
            int32_t val = 0;
            for (int ci = 0; ci < f.n_cols && !found && ok; ci++) {
                int rvv;
                column(f.cols[ci], rvv, val);
                if (rvv == -2) ok = false;
                else if (rvv >= 0) {
                    if (rvv != n_val) ok = false;
                    found = true;
                }
            }
            if (ok && !found && f.kind == GLX_FK_AD) {  // AD falls back to the reference depth as a 1-value field
                n_val = 1;
                val = ref_val;  // ref_rv == 1 was checked above
                found = true;
            }
            if (!ok) break;
            if (found) b |= (uint32_t)RB_FOUND0 << k;
            if (f.kind == GLX_FK_PL) continue;  // a lone band never yields a liftable PL vector; validated only
            // the first value (n_val == 1: col_val itself, else the head of its pool vector)
            const int32_t* vec = R.pool + val;
            if (f.number == GLX_NUM_BASIC) {
                for (int j = 0; j < f.count; j++) c[j] = found ? (n_val == 1 ? val : vec[j]) : dflt;
            } else {
                const int32_t x0 = found ? (n_val == 1 ? val : vec[0]) : dflt;
                c[0] = x0;
                c[1] = (plan.vec_rule[k] && x0 == GLXD_INT_MISSING && dflt == GLXD_INT_MISSING) ? GLXD_INT_VEND : dflt;
                c[2] = dflt;
                c[3] = dflt;
            }
        }
        if (!ok) {
            for (uint32_t j = GLX_BLOB_HDR; j < W; j++) blob[j] = 0;
            break;
        }
        mrd = (int)(unsigned)ref_val;
        bits = b;
    } while (0);
    // ---- a variant record: its view ----
    if ((bits & RB_SLOW) && !(flags & GLX_RF_IS_REF)) {
        bool fits = true;
        uint32_t code[GLX_VIEW_COLS];
#pragma unroll
        for (int j = 0; j < GLX_VIEW_COLS; j++) {
            code[j] = j < plan.n_vc ? view_len_code((uint32_t)clen[j]) : 0u;
            if (code[j] == 0xFFCu) fits = false;
        }
        if (fits) {
            mrd = (int)gtw;
            blob[6] = (int32_t)a_off;
            blob[7] = (int32_t)((uint32_t)(flags & 0x7f) | (uint32_t)na << 8 | code[0] << 16 | code[1] << 28);
            blob[8] = (int32_t)(code[1] >> 4 | code[2] << 8 | code[3] << 20);
            blob[9] = (int32_t)(code[4] | code[5] << 12);
#pragma unroll
            for (int j = 0; j < GLX_VIEW_COLS; j++) blob[10 + j] = j < plan.n_vc ? cval[j] : 0;
            bits |= RB_VIEW;
        }
    }
    blob[0] = rbeg;
    blob[1] = rend;
    blob[2] = rmaxend;
    blob[3] = (int32_t)bits;
    blob[4] = mrd;
    blob[5] = beg_next;
}

// one record -> its blob (thread per record in glx_resolve_records_kernel)
__device__ inline void build_record_blob(const DCfg& cfg, const FastPlan& plan, const DRecs& R, uint64_t r, bool last_of_sample, int32_t* blob) {
    if (plan.vview) {
        build_record_blob_cached(cfg, plan, R, r, last_of_sample, blob);
        return;
    }
    const uint32_t W = blob_words(plan.n_rv);
    for (uint32_t j = GLX_BLOB_HDR; j < W; j++) blob[j] = 0;
    int mrd = -1;
    uint32_t bits = resolve_ref_band(cfg, plan, R, r, blob + GLX_BLOB_HDR, 1, mrd);
    if ((bits & RB_SLOW) && plan.vview && !(R.flags[r] & GLX_RF_IS_REF)) {  // a variant record: its view
        for (uint32_t j = GLX_BLOB_HDR; j < W; j++) blob[j] = 0;
        uint32_t pk[4] = {(uint32_t)(R.flags[r] & 0x7f) | (uint32_t)R.n_allele[r] << 8, 0u, 0u, 0u};
        bool fits = true;
        for (uint32_t j = 0; j < plan.n_vc; j++) {
            const size_t idx = (size_t)plan.vc[j] * R.n_records + r;
            const uint32_t code = view_len_code(R.col_len[idx]);
            if (code == 0xFFCu) fits = false;
            const uint32_t bit = 16 + 12 * j;
            pk[bit >> 5] |= code << (bit & 31);
            if ((bit & 31) > 20) pk[(bit >> 5) + 1] |= code >> (32 - (bit & 31));
            blob[10 + j] = R.col_val[idx];
        }
        if (fits) {
            mrd = (int)R.gt[r];
            blob[6] = (int32_t)R.allele_off[r];
            blob[7] = (int32_t)pk[0];
            blob[8] = (int32_t)pk[1];
            blob[9] = (int32_t)pk[2];
            bits |= RB_VIEW;
        } else
            for (uint32_t j = GLX_BLOB_HDR; j < W; j++) blob[j] = 0;
    }
    blob[0] = R.beg[r];
    blob[1] = R.end[r];
    blob[2] = R.maxend[r];
    blob[3] = (int32_t)bits;
    blob[4] = mrd;
    blob[5] = last_of_sample ? 0x7fffffff : R.beg[r + 1];
}

// Classify the lane's cell at a site; unless CT_SLOW, decide GT / RNC: `called` = 0/0 (else ./.), rnc = RNC_* of both slots.
__device__ __forceinline__ int fast_classify(const DCfg& cfg, const int32_t* __restrict__ blobs, uint32_t W, uint32_t n_rv, LaneCursor& L, int sbeg, int send,
                                             int qbeg, int qend, bool slow_site, bool back, int32_t* rv, int rvs, bool& called, int& rnc) {
    called = false;
    rnc = RNC_M;
    cursor_seek(blobs, W, n_rv, L, qbeg, back, rv, rvs);
    if (slow_site) return CT_SLOW;
    if (L.c >= L.n || L.beg_c >= qend) return CT_MISSING;  // nothing overlaps the query range
    if (L.beg_n < qend) return CT_SLOW;                    // a second record may overlap it
    if (L.bits & RB_SLOW) return CT_SLOW;
    if (!(L.end_c > sbeg && L.beg_c < send)) return CT_MISSING;  // band only inside the unification hull: not kept
    if (!cfg.allow_partial && !(L.beg_c <= sbeg && L.end_c >= send)) {
        rnc = RNC_P;
        return CT_PARTIAL;
    }
    if (L.bits & RB_NONCALLED) rnc = RNC_I;
    else if ((unsigned long long)(long long)L.mrd >= (unsigned long long)cfg.required_dp) {
        called = true;
        rnc = RNC_NA;
    } else
        rnc = RNC_D;
    return CT_REF;
}

// ---------------------------------------------------------------------------------------------------------
// Closed form for a cell whose query range holds SEVERAL records, all of them well-formed reference bands (a site at
// a band boundary, or a multi-base site spanning bands): with no variant record, genotype_site reduces to folds over
// the kept bands — span merge, min_ref_depth, "any band non-called", and per field the combination
// (min / max / missing-if-several) of the bands' values — all available from the bands' blobs. PL2's first-record /
// smaller-margin rule is replayed on the raw PL triples. Runs thread-per-cell in the worklist kernel; returns false
// (nothing written) when some record is not a well-formed band or the config has a field this form does not cover.
__device__ inline bool genotype_cell_bands(const DCfg& cfg, const FastPlan& plan, const DRecs& R, const DSites& S, const DOut& O, const int32_t* __restrict__ blobs,
                                           uint32_t s, uint32_t sample, uint64_t first, uint64_t hi) {
    if (!plan.enabled || S.monoallelic[s]) return false;
    const uint32_t W = blob_words(plan.n_rv);
    const int sbeg = S.beg[s], send = S.end[s], qbeg = S.qbeg[s], qend = S.qend[s];
    const int A = S.n_allele[s];
    const uint32_t N = S.n_samples;
    for (uint32_t k = 0; k < cfg.n_fields; k++)
        if (cfg.f[k].kind == GLX_FK_FT || cfg.f[k].kind == GLX_FK_STRING) return false;
    // pass 1: every record of the query range must be a resolved band; fold the calling state over the kept ones
    int n_kept = 0, mrd = -1;
    bool spanned = false, have_cur = false, noncalled = false;
    int cur_b = 0, cur_e = 0;
    for (uint64_t r = first; r < hi; r++) {
        const int32_t* b = blobs + r * W;
        const int rb = b[0], re = b[1];
        if (rb >= qend) break;
        if (!(re > qbeg)) continue;
        const uint32_t bits = (uint32_t)b[3];
        if (bits & RB_SLOW) return false;
        if (!(re > sbeg && rb < send)) continue;  // a band is kept only if it overlaps site.pos
        n_kept++;
        if (!have_cur) {
            cur_b = rb;
            cur_e = re;
            have_cur = true;
        } else if (rb <= cur_e) {
            cur_e = max(cur_e, re);
        } else {
            if (cur_b <= sbeg && cur_e >= send) spanned = true;
            cur_b = rb;
            cur_e = re;
        }
        const int d = b[4];
        mrd = (mrd < 0) ? d : min(mrd, d);
        if (bits & RB_NONCALLED) noncalled = true;
    }
    if (have_cur && cur_b <= sbeg && cur_e >= send) spanned = true;
    int rnc = RNC_M;
    bool called = false, lift = false;
    if (n_kept == 0) {
    } else if (!cfg.allow_partial && !spanned) {
        rnc = RNC_P;
    } else {
        lift = true;
        if (noncalled) rnc = RNC_I;
        else if ((unsigned long long)(long long)mrd >= (unsigned long long)cfg.required_dp) {
            called = true;
            rnc = RNC_NA;
        } else
            rnc = RNC_D;
    }
    // pass 2: fields
    for (uint32_t k = 0; k < cfg.n_fields; k++) {
        const DFld& f = cfg.f[k];
        const int count = S.fld_cnt[k][s];
        int32_t* out = O.fld[k] + S.fld_off[k][s] + (uint64_t)sample * count;
        const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
        const bool vec_rule = f.type == GLX_TYPE_INT && f.number != GLX_NUM_BASIC && count > 1;
        if (!lift) {
            const int32_t c = f.kind == GLX_FK_PL2 ? 0 : missing;
            for (int j = 0; j < count; j++) out[j] = c;
            if (vec_rule && c == GLXD_INT_MISSING) out[1] = GLXD_INT_VEND;
            continue;
        }
        if (f.kind == GLX_FK_PL) {  // slot 0 gets one value per band with a PL: never a liftable vector
            for (int j = 0; j < count; j++) out[j] = GLXD_INT_MISSING;
            if (vec_rule) out[1] = GLXD_INT_VEND;
            continue;
        }
        if (f.kind == GLX_FK_PL2) {
            int32_t q0 = GLXD_INT_MISSING, q1 = GLXD_INT_MISSING, q2 = GLXD_INT_MISSING;
            bool self_censor = false;
            for (uint64_t r = first; r < hi; r++) {
                const int32_t* b = blobs + r * W;
                if (b[0] >= qend) break;
                if (!(b[1] > qbeg) || !(b[1] > sbeg && b[0] < send) || !((uint32_t)b[3] & ((uint32_t)RB_FOUND0 << k))) continue;
                const int32_t* buf = nullptr;
                if (col_get(R, cfg.col_pl, r, buf) != 3) return false;  // cannot happen for a resolved band
                int32_t v0 = q0;
                if (v0 == 0) {
                    if (buf[0] == 0) {
                        int32_t margin = GLXD_INT_MISSING;
                        for (int j = 1; j < 3; j++)
                            if (buf[j] != GLXD_INT_MISSING) margin = (margin != GLXD_INT_MISSING) ? min(margin, buf[j]) : buf[j];
                        if (margin != GLXD_INT_MISSING) {
                            int32_t old_margin = GLXD_INT_MISSING;  // over the slots of the unified vector: (0,a) = q1, the rest = q2
                            if (q1 != GLXD_INT_MISSING) old_margin = q1;
                            if (q2 != GLXD_INT_MISSING) old_margin = (old_margin != GLXD_INT_MISSING) ? min(old_margin, q2) : q2;
                            if (old_margin != GLXD_INT_MISSING && margin < old_margin) v0 = GLXD_INT_MISSING;
                        }
                    }
                } else if (v0 != GLXD_INT_MISSING) {
                    self_censor = true;
                }
                if (v0 == GLXD_INT_MISSING) {
                    q0 = buf[0];
                    q1 = buf[1];
                    q2 = buf[2];
                }
            }
            const bool zero = q0 == 0 || q1 == 0 || q2 == 0;
            const bool other = (q0 != 0 && q0 != GLXD_INT_MISSING) || (q1 != 0 && q1 != GLXD_INT_MISSING) || (q2 != 0 && q2 != GLXD_INT_MISSING);
            if (self_censor || !(zero && other)) q0 = q1 = q2 = 0;
            else {
                if (q0 == GLXD_INT_MISSING) q0 = 990;
                if (q1 == GLXD_INT_MISSING) q1 = 990;
                if (q2 == GLXD_INT_MISSING) q2 = 990;
            }
            int a = 0, bb = 0;
            for (int j = 0; j < count; j++) {
                out[j] = a == 0 ? q0 : (bb == 0 ? q1 : q2);
                if (++bb > a) {
                    a++;
                    bb = 0;
                }
            }
            continue;
        }
        // numeric kinds: BASIC-numbered fields combine every slot, allele/genotype-numbered ones only slot 0 (a band's map is [0,-1])
        const int n_slots = f.number == GLX_NUM_BASIC ? count : 1;
        const int32_t dflt = f.default_zero ? 0 : missing;
        for (int j = 0; j < n_slots; j++) {
            int n_found = 0;
            int32_t acc = dflt;
            for (uint64_t r = first; r < hi; r++) {
                const int32_t* b = blobs + r * W;
                if (b[0] >= qend) break;
                if (!(b[1] > qbeg) || !(b[1] > sbeg && b[0] < send) || !((uint32_t)b[3] & ((uint32_t)RB_FOUND0 << k))) continue;
                const int32_t v = b[GLX_BLOB_HDR + plan.rv0[k] + j];
                if (n_found == 0) acc = v;
                else if (f.combi == GLX_COMBI_MIN || f.combi == GLX_COMBI_MAX) {
                    bool take;
                    if (f.type == GLX_TYPE_FLOAT) take = f.combi == GLX_COMBI_MIN ? (__int_as_float(v) < __int_as_float(acc)) : (__int_as_float(acc) < __int_as_float(v));
                    else take = f.combi == GLX_COMBI_MIN ? (v < acc) : (acc < v);
                    if (take) acc = v;
                }
                n_found++;
            }
            if (n_found > 1 && f.combi != GLX_COMBI_MIN && f.combi != GLX_COMBI_MAX) acc = missing;
            out[j] = acc;
        }
        for (int j = n_slots; j < count; j++) out[j] = dflt;
        if (vec_rule) {
            bool all_missing = true;
            for (int j = 0; j < count; j++)
                if (out[j] != GLXD_INT_MISSING) all_missing = false;
            if (all_missing) out[1] = GLXD_INT_VEND;
        }
    }
    const size_t cell = ((size_t)s * N + sample) * 2;
    const signed char ga = called ? 0 : -1;
    const unsigned char rc = (unsigned char)kRncChar[rnc];
    *reinterpret_cast<char2*>(O.gt + cell) = make_char2(ga, ga);
    *reinterpret_cast<uchar2*>(O.rnc + cell) = make_uchar2(rc, rc);
    if (called) atomicOr(O.allele_used + s, 1u);
    if (rnc == RNC_I) {
        uint32_t* wd = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(O.site_flags + s) & ~(uintptr_t)3);
        atomicOr(wd, 1u << (8 * (reinterpret_cast<uintptr_t>(O.site_flags + s) & 3)));
    }
    (void)A;
    return true;
}

// write field k's vector of a finished cell (o = the cell's first slot)
__device__ __forceinline__ void fast_emit_field(const FastPlan& plan, int k, bool ref, int count, int32_t* __restrict__ o, const int32_t* rv, int rvs) {
    const int32_t cen = plan.cen[k];
    const int mode = plan.mode[k];
    const int32_t* c = rv + (size_t)plan.rv0[k] * rvs;
    if (mode == FM_VEC) {
        for (int j = 0; j < count; j++) o[j] = ref ? c[(size_t)j * rvs] : cen;
        return;
    }
    int32_t x0 = cen, x1 = plan.vec_rule[k] ? GLXD_INT_VEND : cen, xr = cen, xt = cen;
    if (ref && mode == FM_PATTERN) {
        x0 = c[0];
        x1 = c[(size_t)rvs];
        xr = c[(size_t)2 * rvs];
        xt = c[(size_t)3 * rvs];
    }
    o[0] = x0;
    if (count > 1) o[1] = x1;
    int next_row = 3, a = 2;  // slot of genotype (0,a)
    for (int j = 2; j < count; j++) {
        int32_t v = xt;
        if (j == next_row) {
            v = xr;
            a++;
            next_row += a;
        }
        o[j] = v;
    }
}

}  // namespace glxd
