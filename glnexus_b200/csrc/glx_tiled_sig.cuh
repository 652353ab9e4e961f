// glx_tiled_sig.cuh — the tiled fast path specialised at compile time on the FIELD SIGNATURE of the config
// (how many lifted fields, and for each: FM_* mode, RetainedFieldNumber, cached-value count). With the signature a
// template parameter every per-field loop unrolls, a lane's cached band values live in registers, BASIC-numbered
// fields get their output offset arithmetically, and a cursor advance is four 128-bit loads. Semantics are those of
// glx_cell_fast.cuh (same blobs, same classification, same patterns); configs whose signature has no instantiation
// run the dynamic kernel there. Reference lines: see glx_cell_fast.cuh.
#pragma once
#include <type_traits>
#include "glx_cell_fast.cuh"
#include "glx_cell_variant.cuh"

namespace glxd {

// signature = 4 bits per field in three words
struct FieldSig {
    uint32_t nf, modes_lo, modes_hi, nums_lo, nums_hi, nrvs_lo, nrvs_hi;
};
#define GLX_SIG4(x, k) (((x) >> (4 * (k))) & 15u)

inline bool sig_of_plan(const DCfg& c, const FastPlan& p, uint32_t& nf, uint64_t& modes, uint64_t& nums, uint64_t& nrvs) {
    nf = c.n_fields;
    modes = nums = nrvs = 0;
    if (!p.enabled || nf > GLX_MAX_FIELDS) return false;
    for (uint32_t k = 0; k < nf; k++) {
        if (p.nrv[k] > 15) return false;
        modes |= (uint64_t)p.mode[k] << (4 * k);
        nums |= (uint64_t)c.f[k].number << (4 * k);
        nrvs |= (uint64_t)p.nrv[k] << (4 * k);
    }
    return true;
}

template <uint32_t NF, uint64_t MODES, uint64_t NUMS, uint64_t NRVS>
struct TiledSig {
    static constexpr uint32_t nf = NF;
    static __host__ __device__ constexpr uint32_t mode(uint32_t k) { return (uint32_t)((MODES >> (4 * k)) & 15u); }
    static __host__ __device__ constexpr uint32_t num(uint32_t k) { return (uint32_t)((NUMS >> (4 * k)) & 15u); }
    static __host__ __device__ constexpr uint32_t nrv(uint32_t k) { return (uint32_t)((NRVS >> (4 * k)) & 15u); }
    static __host__ __device__ constexpr uint32_t rv0(uint32_t k) { return k == 0 ? 0u : rv0(k - 1) + nrv(k - 1); }
    static constexpr uint32_t n_rv = rv0(NF);
    static constexpr uint32_t W = ((GLX_BLOB_HDR + n_rv + 3u) & ~3u) < GLX_BLOB_MIN ? GLX_BLOB_MIN : ((GLX_BLOB_HDR + n_rv + 3u) & ~3u);
    static constexpr uint32_t NRV_ARR = n_rv ? n_rv : 1;
};

// build_record_blob_cached with the field table unrolled over the signature: the blob is a register array with
// compile-time word positions, so the resolve kernel can store its row straight from registers (no shared-memory staging,
// no block barrier). Same result as build_record_blob for a plan with views whose signature is Sig.
template <class Sig>
__device__ __forceinline__ void build_record_blob_sig(const DCfg& cfg, const FastPlan& plan, const DRecs& R, uint64_t r, bool last_of_sample, int32_t (&w)[Sig::W]) {
    static_assert(Sig::W == 16, "the variant view occupies words 6..15 of a 16-word row");
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
#pragma unroll
    for (uint32_t j = GLX_BLOB_HDR; j < Sig::W; j++) w[j] = 0;
    int mrd = -1;
    uint32_t bits = RB_SLOW;
    // ---- a well-formed reference band? ----
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
        glx_static_for<Sig::nf>([&](auto kc) {
            constexpr uint32_t k = decltype(kc)::value;
            if (!ok) return;
            const DFld& f = cfg.f[k];
            constexpr uint32_t c0 = GLX_BLOB_HDR + Sig::rv0(k);
            const int32_t missing = f.type == GLX_TYPE_FLOAT ? (int32_t)GLX_FLOAT_MISSING_BITS : GLXD_INT_MISSING;
            const int32_t dflt = f.default_zero ? 0 : missing;
            if (Sig::mode(k) == FM_PATTERN && f.kind == GLX_FK_PL2) {
                int rvv;
                int32_t val;
                column(cfg.col_pl, rvv, val);
                int32_t p0 = 0, p1 = 0, p2 = 0;
                if (rvv > 0) {
                    if (rvv != 3) {
                        ok = false;
                        return;
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
                if (Sig::nrv(k) >= 4) {
                    w[c0] = p0;
                    w[c0 + 1] = p1;
                    w[c0 + 2] = p1;
                    w[c0 + 3] = p2;
                }
                return;
            }
            int n_val = f.count;
            if (Sig::num(k) == GLX_NUM_ALT) n_val = 1;
            else if (Sig::num(k) == GLX_NUM_ALLELES) n_val = 2;
            else if (Sig::num(k) == GLX_NUM_GENOTYPE) n_val = 3;
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
                val = ref_val;
                found = true;
            }
            if (!ok) return;
            if (found) b |= (uint32_t)RB_FOUND0 << k;
            if (Sig::mode(k) == FM_CONST) return;  // PL of a lone band: validated only
            const int32_t* vec = R.pool + val;
            if (Sig::mode(k) == FM_VEC) {
#pragma unroll
                for (uint32_t j = 0; j < Sig::nrv(k); j++) w[c0 + j] = found ? (n_val == 1 ? val : vec[j]) : dflt;
            } else if (Sig::nrv(k) >= 4) {
                const int32_t x0 = found ? (n_val == 1 ? val : vec[0]) : dflt;
                w[c0] = x0;
                w[c0 + 1] = (plan.vec_rule[k] && x0 == GLXD_INT_MISSING && dflt == GLXD_INT_MISSING) ? GLXD_INT_VEND : dflt;
                w[c0 + 2] = dflt;
                w[c0 + 3] = dflt;
            }
        });
        if (!ok) {
#pragma unroll
            for (uint32_t j = GLX_BLOB_HDR; j < Sig::W; j++) w[j] = 0;
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
            w[6] = (int32_t)a_off;
            w[7] = (int32_t)((uint32_t)(flags & 0x7f) | (uint32_t)na << 8 | code[0] << 16 | code[1] << 28);
            w[8] = (int32_t)(code[1] >> 4 | code[2] << 8 | code[3] << 20);
            w[9] = (int32_t)(code[4] | code[5] << 12);
#pragma unroll
            for (int j = 0; j < GLX_VIEW_COLS; j++) w[10 + j] = j < plan.n_vc ? cval[j] : 0;
            bits |= RB_VIEW;
        }
    }
    w[0] = rbeg;
    w[1] = rend;
    w[2] = rmaxend;
    w[3] = (int32_t)bits;
    w[4] = mrd;
    w[5] = beg_next;
}

// A lane reads its sample's blobs in order, one 64-byte row per cursor step, and waits for each: ask L2 to bring the
// following rows along (256-byte prefetch granule), so that the next steps find theirs in L2 instead of HBM.
__device__ __forceinline__ int4 ldg_blob(const int4* p) {
#if defined(__CUDA_ARCH__) && !defined(GLX_ABLATE_BLOB_PREFETCH)
    int4 v;
    asm volatile("ld.global.nc.L2::256B.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// The lane's NEXT blob row, fetched ahead of need. A lane walks its sample's rows in order and, without this, waits a full
// L2 / HBM round trip at every cursor step (the profile's top stall: profiles/r1_ncu_full.md). BlobPrefetch keeps one
// 64-byte row per lane in flight: a TMA bulk copy (cp.async.bulk, SASS UBLKCP) global -> the lane's shared-memory slot,
// completion counted on the lane's own mbarrier, issued when the current row is taken and consumed — usually many sites
// later — with a shared-memory read. NoPrefetch is the plain form (host emulation, dynamic kernel).
#ifndef GLX_ROW_PREFETCH
#define GLX_ROW_PREFETCH 2  // 1: TMA bulk copies + per-lane mbarriers, 2: per-thread cp.async
#endif
struct NoPrefetch {
    template <uint32_t W>
    __host__ __device__ __forceinline__ bool take(uint32_t, int32_t (&)[W]) { return false; }
    __host__ __device__ __forceinline__ void issue(const int32_t*, uint32_t, int) {}
    __host__ __device__ __forceinline__ void drain() {}
};
#if defined(__CUDACC__)
template <uint32_t W>
struct BlobPrefetch {
    static constexpr uint32_t NONE = 0xffffffffu;
    uint32_t slot, bar;  // shared-space addresses: the lane's row slot (W words, 16-byte aligned) and its mbarrier
    uint32_t row;        // row index (within the sample) in flight or landed in the slot, NONE if nothing is
    uint32_t phase;
    int qmax;            // largest qbeg of the tile: a row whose maxend exceeds it is the lane's last in this tile
    __device__ __forceinline__ void init(void* slot_p, void* bar_p, int qmax_) {
        slot = (uint32_t)__cvta_generic_to_shared(slot_p);
        bar = (uint32_t)__cvta_generic_to_shared(bar_p);
        row = NONE;
        phase = 0;
        qmax = qmax_;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __device__ __forceinline__ void wait() {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "GLX_PF_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra GLX_PF_DONE;\n"
            "bra GLX_PF_WAIT;\n"
            "GLX_PF_DONE:\n"
            "}" ::"r"(bar),
            "r"(phase)
            : "memory");
        phase ^= 1u;
    }
    __device__ __forceinline__ bool take(uint32_t want, int32_t (&w)[W]) {
        if (row == NONE) return false;
        wait();
        const bool hit = row == want;
        row = NONE;
        if (!hit) return false;
#pragma unroll
        for (uint32_t q = 0; q < W / 4; q++)
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * q]), "=r"(w[4 * q + 1]), "=r"(w[4 * q + 2]), "=r"(w[4 * q + 3]) : "r"(slot + 16u * q) : "memory");
        return true;
    }
    // fetch row r (at src) unless the row just taken already reaches past the tile
    __device__ __forceinline__ void issue(const int32_t* src, uint32_t r, int maxend) {
        if (maxend > qmax) return;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(W * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot), "l"(src), "n"(W * 4), "r"(bar) : "memory");
        row = r;
    }
    __device__ __forceinline__ void drain() {  // nothing may be in flight when the CTA's shared memory goes away
        if (row != NONE) wait();
        row = NONE;
    }
};
// the same with per-thread cp.async (LDGSTS) copies: no mbarrier, no per-lane serialisation of the issue
template <uint32_t W>
struct BlobPrefetchLdgsts {
    static constexpr uint32_t NONE = 0xffffffffu;
    uint32_t slot, row;
    int qmax;
    __device__ __forceinline__ void init(void* slot_p, void*, int qmax_) {
        slot = (uint32_t)__cvta_generic_to_shared(slot_p);
        row = NONE;
        qmax = qmax_;
    }
    __device__ __forceinline__ bool take(uint32_t want, int32_t (&w)[W]) {
        if (row == NONE) return false;
        asm volatile("cp.async.wait_all;" ::: "memory");
        const bool hit = row == want;
        row = NONE;
        if (!hit) return false;
#pragma unroll
        for (uint32_t q = 0; q < W / 4; q++)
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * q]), "=r"(w[4 * q + 1]), "=r"(w[4 * q + 2]), "=r"(w[4 * q + 3]) : "r"(slot + 16u * q) : "memory");
        return true;
    }
    __device__ __forceinline__ void issue(const int32_t* src, uint32_t r, int maxend) {
        if (maxend > qmax) return;
#pragma unroll
        for (uint32_t q = 0; q < W / 4; q++) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + 16u * q), "l"(src + 4 * q) : "memory");
        row = r;
    }
    __device__ __forceinline__ void drain() {
        if (row != NONE) asm volatile("cp.async.wait_all;" ::: "memory");
        row = NONE;
    }
};
#endif

template <class Sig>
struct TiledLane {
    uint64_t base;
    uint32_t n, c;
    int beg_c, end_c, maxend_c, beg_n;
    uint32_t bits;
    int mrd;
    int32_t rv[Sig::NRV_ARR];

    // load record c's blob straight into registers (four 128-bit loads for the preset signatures)
    template <class PF>
    __device__ __forceinline__ void load(const int32_t* __restrict__ blobs, PF& pf) {
        if (c < n) {
            const int4* b = reinterpret_cast<const int4*>(blobs + (base + c) * Sig::W);
            int32_t w[Sig::W];
            if (!pf.take(c, w)) {
#pragma unroll
                for (uint32_t q = 0; q < Sig::W / 4; q++) {
                    const int4 v = ldg_blob(b + q);
                    w[4 * q] = v.x;
                    w[4 * q + 1] = v.y;
                    w[4 * q + 2] = v.z;
                    w[4 * q + 3] = v.w;
                }
#if defined(__CUDA_ARCH__) && defined(GLX_L2_AHEAD)
                if (c + GLX_L2_AHEAD < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + (Sig::W / 4) * GLX_L2_AHEAD));
#endif
            }
            // (skipping a prefetch is always allowed: the test on the row's last words only makes the copy into the slot wait
            // until every shared-memory read of the row taken from that slot has returned)
            if (c + 1 < n && (w[Sig::W / 4 + 3] ^ w[Sig::W / 2 + 3] ^ w[Sig::W - 1]) != (int32_t)0x5a17c3e9) pf.issue(blobs + (base + c + 1) * Sig::W, c + 1, w[2]);
            beg_c = w[0];
            end_c = w[1];
            maxend_c = w[2];
            bits = (uint32_t)w[3];
            mrd = w[4];
            beg_n = w[5];
#pragma unroll
            for (uint32_t j = 0; j < Sig::n_rv; j++) rv[j] = w[GLX_BLOB_HDR + j];
        } else {
            beg_c = beg_n = maxend_c = end_c = 0x7fffffff;
            bits = RB_SLOW;
            mrd = -1;
        }
    }
    __device__ __forceinline__ void load(const int32_t* __restrict__ blobs) {
        NoPrefetch pf;
        load(blobs, pf);
    }
    template <class PF>
    __device__ __forceinline__ void init(const DRecs& R, const int32_t* __restrict__ blobs, uint32_t sample, uint32_t tile_cursor, PF& pf) {
        base = R.ds_rec_off[sample];
        const uint64_t hi = R.ds_rec_off[sample + 1];
        n = (uint32_t)(hi - base);
        c = tile_cursor < n ? tile_cursor : n;  // tile cursor table (fill_tile_cursors)
        load(blobs, pf);
    }
    __device__ __forceinline__ void init(const DRecs& R, const int32_t* __restrict__ blobs, uint32_t sample, uint32_t tile_cursor) {
        NoPrefetch pf;
        init(R, blobs, sample, tile_cursor, pf);
    }
    __device__ __forceinline__ void idle() {
        base = 0;
        n = c = 0;
        beg_c = beg_n = maxend_c = end_c = 0x7fffffff;
        bits = RB_SLOW;
        mrd = -1;
#pragma unroll
        for (uint32_t j = 0; j < Sig::NRV_ARR; j++) rv[j] = 0;
    }
    template <class PF>
    __device__ __forceinline__ void seek(const int32_t* __restrict__ blobs, int qbeg, bool back, PF& pf) {
        if (back) {
            bool moved = false;
            while (c > 0 && blobs[(base + c - 1) * Sig::W + 2] > qbeg) {
                c--;
                moved = true;
            }
            if (moved) load(blobs, pf);
        }
        while (c < n && maxend_c <= qbeg) {
            c++;
            load(blobs, pf);
        }
    }
    // same decision table as fast_classify, written without early exits (a handful of predicate ops instead of a
    // branch ladder that the warp would walk divergently)
    __device__ __forceinline__ int classify(const DCfg& cfg, const int32_t* __restrict__ blobs, int sbeg, int send, int qbeg, int qend, bool slow_site, bool back,
                                            bool& called, int& rnc) {
        NoPrefetch pf;
        return classify(cfg, blobs, sbeg, send, qbeg, qend, slow_site, back, called, rnc, pf);
    }
    template <class PF>
    __device__ __forceinline__ int classify(const DCfg& cfg, const int32_t* __restrict__ blobs, int sbeg, int send, int qbeg, int qend, bool slow_site, bool back,
                                            bool& called, int& rnc, PF& pf) {
        seek(blobs, qbeg, back, pf);
        const bool none = beg_c >= qend;  // also true past the sample's last record (sentinel)
        const bool slow = slow_site | (!none & ((beg_n < qend) | ((bits & RB_SLOW) != 0)));
        const bool overl = (end_c > sbeg) & (beg_c < send);
        const bool spans = (beg_c <= sbeg) & (end_c >= send);
        const bool kept = !none & !slow & overl;
        const bool partial = kept & !cfg.allow_partial & !spans;
        const bool ref = kept & !partial;
        const bool noncalled = (bits & RB_NONCALLED) != 0;
        called = ref & !noncalled & ((mrd < 0) | ((uint32_t)mrd >= cfg.required_dp));
        rnc = partial ? RNC_P : (!ref ? RNC_M : (noncalled ? RNC_I : (called ? RNC_NA : RNC_D)));
        return slow ? CT_SLOW : (ref ? CT_REF : (partial ? CT_PARTIAL : CT_MISSING));
    }
    // field K's vector of a finished cell. A = unified alleles of the site; off_dyn(k) = fld_off[k][s] for
    // allele/genotype-numbered fields (BASIC-numbered fields sit at s*N*count).
    // `done`: this lane finished its cell (otherwise it only takes part in the warp-wide transposed stores; the
    // worklist kernels rewrite its slots later). `stage`: the warp's 32 x GLX_STAGE_MAX staging buffer.
    template <uint32_t K, class OffFn>
    __device__ __forceinline__ void emit_one(const FastPlan& plan, const DOut& O, bool done, bool ref, uint32_t s, uint32_t N, uint32_t sample, uint32_t A,
                                             OffFn off_dyn, int32_t* stage, uint32_t lane, uint32_t nvalid) const {
        const int32_t cen = plan.cen[K];
        constexpr uint32_t r0 = Sig::rv0(K);
        if constexpr (Sig::mode(K) == FM_VEC) {
            if (!done) return;
            constexpr uint32_t cnt = Sig::nrv(K);
#ifdef GLX_ABLATE_WRAP
            int32_t* __restrict__ o = O.fld[K] + ((((unsigned long long)s * N + sample) * cnt) & ((1ull << 21) - 1));
#else
            int32_t* __restrict__ o = O.fld[K] + ((unsigned long long)s * N + sample) * cnt;
#endif
            int32_t v[cnt ? cnt : 1];
#pragma unroll
            for (uint32_t j = 0; j < cnt; j++) v[j] = ref ? rv[r0 + j] : cen;
            if constexpr (cnt == 4) *reinterpret_cast<int4*>(o) = make_int4(v[0], v[1], v[2], v[3]);
            else if constexpr (cnt == 2) *reinterpret_cast<int2*>(o) = make_int2(v[0], v[1]);
            else {
#pragma unroll
                for (uint32_t j = 0; j < cnt; j++) o[j] = v[j];
            }
        } else {
            const uint32_t count = Sig::num(K) == GLX_NUM_ALLELES ? A : (Sig::num(K) == GLX_NUM_ALT ? A - 1 : A * (A + 1) / 2);
            const unsigned long long off = off_dyn(K);
            int32_t x0 = cen, x1 = plan.vec_rule[K] ? GLXD_INT_VEND : cen, xr = cen, xt = cen;
            if constexpr (Sig::mode(K) == FM_PATTERN) {
                if (ref) {
                    x0 = rv[r0];
                    x1 = rv[r0 + 1];
                    xr = rv[r0 + 2];
                    xt = rv[r0 + 3];
                }
            }
#ifdef __CUDACC__
            if (count >= 3 && count <= GLX_STAGE_MAX) {
                // A lane's vector of 3..32 ints is not a whole number of sectors: written lane by lane every
                // store instruction would touch a third of each sector. Transpose through shared memory so that the warp
                // writes its 32*count ints as `count` fully coalesced rows.
                int32_t* st = stage + lane * count;
                st[0] = x0;
                st[1] = x1;
                st[2] = xt;
                int next_row = 3, a = 2;
                for (uint32_t j = 3; j < count; j++) {
                    int32_t v = xt;
                    if ((int)j == next_row) {
                        v = xr;
                        a++;
                        next_row += a;
                    }
                    st[j] = v;
                }
                __syncwarp();
#ifdef GLX_ABLATE_WRAP
                int32_t* __restrict__ ow = O.fld[K] + ((off + (unsigned long long)(sample - lane) * count) & ((1ull << 21) - 1));
#else
                int32_t* __restrict__ ow = O.fld[K] + off + (unsigned long long)(sample - lane) * count;
#endif
                const uint32_t total = nvalid * count;
                if (count == 3) {  // biallelic site, the common case: three rows, no loop
#pragma unroll
                    for (uint32_t u = 0; u < 3; u++)
                        if (lane + 32 * u < total) ow[lane + 32 * u] = stage[lane + 32 * u];
                } else
                    for (uint32_t e = lane; e < total; e += 32) ow[e] = stage[e];
                __syncwarp();
                return;
            }
#endif
            if (!done) return;
#ifdef GLX_ABLATE_WRAP
            int32_t* __restrict__ o = O.fld[K] + ((off + (unsigned long long)sample * count) & ((1ull << 21) - 1));
#else
            int32_t* __restrict__ o = O.fld[K] + off + (unsigned long long)sample * count;
#endif
            if (count == 2 && !(off & 1)) {
                *reinterpret_cast<int2*>(o) = make_int2(x0, x1);
            } else {
                o[0] = x0;
                if (count > 1) o[1] = x1;
                if (count > 2) o[2] = xt;
                int next_row = 3, a = 2;
                for (uint32_t j = 3; j < count; j++) {
                    int32_t v = xt;
                    if ((int)j == next_row) {
                        v = xr;
                        a++;
                        next_row += a;
                    }
                    o[j] = v;
                }
            }
        }
    }
    template <uint32_t K, class OffFn>
    __device__ __forceinline__ void emit_from(const FastPlan& plan, const DOut& O, bool done, bool ref, uint32_t s, uint32_t N, uint32_t sample, uint32_t A,
                                              OffFn off_dyn, int32_t* stage, uint32_t lane, uint32_t nvalid) const {
        if constexpr (K < Sig::nf) {
            emit_one<K>(plan, O, done, ref, s, N, sample, A, off_dyn, stage, lane, nvalid);
            emit_from<K + 1>(plan, O, done, ref, s, N, sample, A, off_dyn, stage, lane, nvalid);
        }
    }
    // all lifted vectors of the lane's cell at site s; must be called by every lane of the warp
    template <class OffFn>
    __device__ __forceinline__ void emit(const FastPlan& plan, const DOut& O, bool done, bool ref, uint32_t s, uint32_t N, uint32_t sample, uint32_t A,
                                         OffFn off_dyn, int32_t* stage, uint32_t lane, uint32_t nvalid) const {
        emit_from<0>(plan, O, done, ref, s, N, sample, A, off_dyn, stage, lane, nvalid);
    }
};

#ifdef __CUDACC__
// TS = sites per tile. FUSE = 4: when the tile is done the CTA also finishes its own lone-variant-record cells (genotype_cell_variant<4>),
// while the tile's output lines and record blobs are still in L2. FUSE = 0: they are left to glx_cells_variant_kernel.
#ifndef GLX_FUSED_MIN_BLOCKS
#define GLX_FUSED_MIN_BLOCKS 7  // 72 registers: the walk over the tile wants the occupancy more than the fused closed form wants registers
#endif
template <class Sig, int FUSE, int TS>
__global__ void __launch_bounds__(GLX_TILE_THREADS, FUSE ? GLX_FUSED_MIN_BLOCKS : 1) glx_cells_tiled_sig_kernel(const DCfg cfg, const FastPlan plan, const DRecs R, const DSites S, const DOut O,
                                                                               const int32_t* __restrict__ blobs, const uint32_t* __restrict__ tile_cursors,
                                                                               uint2* __restrict__ worklist,
                                                                               uint32_t* __restrict__ n_work, uint32_t n_sample_blocks, uint32_t cta0) {
    static_assert(TS <= 64, "per-lane site masks are 64-bit");
    // per site of the tile, one record: {beg, end, qbeg, qend | meta | - | fld_off of the allele/genotype-numbered fields}; meta: bits 0-7
    // n_allele, bit 8 general path only, bit 9 qbeg below the previous site's. The walk reads it through one running
    // 32-bit shared address (explicit ld.shared: the compiler's uniform-datapath addressing recomputed the window base every site).
    constexpr uint32_t SITE_W = (6u + 2u * (Sig::nf ? Sig::nf : 1) + 3u) & ~3u;  // words, a multiple of 16 bytes
    __shared__ __align__(16) uint32_t s_site[TS][SITE_W];
    __shared__ uint32_t s_red[4][GLX_TILE_THREADS / 32];  // per-warp called/lost site masks (lo, hi words)
    __shared__ uint32_t s_nslow;  // cells this CTA handed to the worklist kernels (its worklist region is CTA-private)
    __shared__ int32_t s_stage[GLX_TILE_THREADS / 32][32 * GLX_STAGE_MAX];  // per-warp transposition buffer for narrow vectors
#ifndef GLX_ABLATE_ROW_PREFETCH
    __shared__ __align__(16) int32_t s_next[GLX_TILE_THREADS][Sig::W];  // per lane: its next blob row (BlobPrefetch)
    __shared__ __align__(8) unsigned long long s_bar[GLX_ROW_PREFETCH == 1 ? GLX_TILE_THREADS : 1];
    __shared__ int s_qmax;
    if (threadIdx.x == 0) s_qmax = (int)0x80000000;
    __syncthreads();
#endif

    const uint32_t cta = blockIdx.x + cta0;  // launches cover chunks of the tile grid so that the worklist kernels of one chunk overlap the next
    const uint32_t sb = cta % n_sample_blocks, tile = cta / n_sample_blocks;
    const uint32_t site0 = tile * TS;
    uint2* __restrict__ my_work = worklist + (size_t)cta * (TS * GLX_TILE_THREADS);
    if (threadIdx.x == 0) s_nslow = 0;
    const uint32_t nsite = min((uint32_t)TS, S.n_sites - site0);
    const uint32_t N = S.n_samples;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t sample = sb * GLX_TILE_THREADS + tid;
    const bool active = sample < N;
    const uint32_t warp_sample0 = sample - lane;
    const uint32_t nvalid = warp_sample0 >= N ? 0u : min(32u, N - warp_sample0);

    for (uint32_t i = tid; i < nsite; i += GLX_TILE_THREADS) {
        const uint32_t s = site0 + i;
        const int qb = S.qbeg[s];
        *reinterpret_cast<int4*>(&s_site[i][0]) = make_int4(S.beg[s], S.end[s], qb, S.qend[s]);
#ifndef GLX_ABLATE_ROW_PREFETCH
        atomicMax(&s_qmax, qb);
#endif
        s_site[i][4] = (uint32_t)S.n_allele[s] | (S.monoallelic[s] ? 256u : 0u) | ((i > 0 && qb < S.qbeg[s - 1]) ? 512u : 0u);
#pragma unroll
        for (uint32_t k = 0; k < Sig::nf; k++)
            if (Sig::mode(k) != FM_VEC) *reinterpret_cast<unsigned long long*>(&s_site[i][6 + 2 * k]) = S.fld_off[k][s];
    }
    __syncthreads();
    uint32_t site_a = (uint32_t)__cvta_generic_to_shared(&s_site[0][0]);

    TiledLane<Sig> L;
#ifndef GLX_ABLATE_ROW_PREFETCH
#if GLX_ROW_PREFETCH == 1
    BlobPrefetch<Sig::W> pf;
    pf.init(&s_next[tid][0], &s_bar[tid], s_qmax);
#else
    BlobPrefetchLdgsts<Sig::W> pf;
    pf.init(&s_next[tid][0], nullptr, s_qmax);
#endif
#else
    NoPrefetch pf;
#endif
    if (active) L.init(R, blobs, sample, tile_cursors[(size_t)tile * N + sample], pf);
    else L.idle();
    // sites (bit i of the tile) where this lane called 0/0, resp. produced a "lost call" RNC; reduced once per tile
    typedef typename std::conditional<(TS <= 32), uint32_t, unsigned long long>::type site_mask_t;
    site_mask_t called_bits = 0, lost_bits = 0;
    char2* __restrict__ gt_p = reinterpret_cast<char2*>(O.gt) + ((size_t)site0 * N + sample);
    uchar2* __restrict__ rnc_p = reinterpret_cast<uchar2*>(O.rnc) + ((size_t)site0 * N + sample);

    for (uint32_t i = 0; i < nsite; i++, gt_p += N, rnc_p += N, site_a += SITE_W * 4u) {
        const uint32_t s = site0 + i;
        int4 pos;
        uint32_t meta;
        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(pos.x), "=r"(pos.y), "=r"(pos.z), "=r"(pos.w) : "r"(site_a));
        asm volatile("ld.shared.u32 %0, [%1+16];" : "=r"(meta) : "r"(site_a));
        int type = CT_MISSING, rnc = RNC_M;
        bool called = false;
        if (active) type = L.classify(cfg, blobs, pos.x, pos.y, pos.z, pos.w, meta & 256u, meta & 512u, called, rnc, pf);
        const bool slow = active && type == CT_SLOW;
        // worklist entry: each slow lane takes its own slot of the CTA's region (order within the region is immaterial)
        if (slow)
            my_work[atomicAdd(&s_nslow, 1u)] =
                make_uint2(s * N + sample, L.c | ((!(meta & 256u) && L.c < L.n && L.beg_c < pos.w && L.beg_n >= pos.w) ? GLX_WORK_SINGLE : 0u));
        const bool done = active && !slow;
        if (done) {
            called_bits |= (site_mask_t)called << i;
            lost_bits |= (site_mask_t)(rnc == RNC_I) << i;
            const signed char a = called ? 0 : -1;
            const unsigned char rc = (unsigned char)kRncChar[rnc];
            *gt_p = make_char2(a, a);
            *rnc_p = make_uchar2(rc, rc);
        }
        L.emit(plan, O, done, type == CT_REF, s, N, sample, meta & 255u,
               [&](uint32_t k) {
                   unsigned long long off;
                   asm volatile("ld.shared.u64 %0, [%1];" : "=l"(off) : "r"(site_a + 24u + 8u * k));
                   return off;
               },
               s_stage[tid >> 5], lane, nvalid);
    }
    pf.drain();
    // ---- per-site reductions for the whole tile: allele 0 used / some call lost ----
    {
        const uint32_t w = tid >> 5;
        const uint32_t c_lo = __reduce_or_sync(0xffffffffu, (uint32_t)called_bits), l_lo = __reduce_or_sync(0xffffffffu, (uint32_t)lost_bits);
        uint32_t c_hi = 0, l_hi = 0;
        if constexpr (TS > 32) {
            c_hi = __reduce_or_sync(0xffffffffu, (uint32_t)((unsigned long long)called_bits >> 32));
            l_hi = __reduce_or_sync(0xffffffffu, (uint32_t)((unsigned long long)lost_bits >> 32));
        }
        if (lane == 0) {
            s_red[0][w] = c_lo;
            s_red[1][w] = c_hi;
            s_red[2][w] = l_lo;
            s_red[3][w] = l_hi;
        }
    }
    __syncthreads();
    if (FUSE == 0 && tid == 0) n_work[cta] = s_nslow;
    if constexpr (FUSE > 0) {
        __shared__ uint32_t s_ndone;
        if (tid == 0) s_ndone = 0;
        __syncthreads();
        uint32_t my_done = 0;
        const uint32_t n_items = s_nslow;
        for (uint32_t it = tid; it < n_items; it += GLX_TILE_THREADS) {
            const uint2 e = my_work[it];
            if (!(e.y & GLX_WORK_SINGLE)) continue;
            const uint32_t ws = e.x / N, wsample = e.x - ws * N;
            const uint64_t first = R.ds_rec_off[wsample] + (e.y & GLX_WORK_CURSOR);
            // The CTA itself only runs the biallelic closed form (straight-line, registers only): nine cells in ten. A cell it
            // declines would cost the whole warp the general closed form (2 k instructions) for a lane or two, so it goes to the
            // launch's declined list, which glx_cells_variant_list_kernel works off with every lane busy (a full list: the entry
            // stays in the region for glx_cells_single_kernel).
            if (genotype_cell_biallelic<Sig>(cfg, plan, R, S, O, blobs, ws, wsample, first)) {
                my_work[it].y = e.y | GLX_WORK_DONE;
                my_done++;
            } else if (plan.decl_list) {
                const uint32_t pos = atomicAdd(plan.decl_n, 1u);
                if (pos < plan.decl_cap) plan.decl_list[pos] = make_uint2(cta, it);
            }
        }
        if (my_done) atomicAdd(&s_ndone, my_done);
        __syncthreads();
        if (tid == 0) n_work[cta] = s_nslow | (s_ndone << 16);  // high half: cells this CTA finished itself
    }
    for (uint32_t i = tid; i < nsite; i += GLX_TILE_THREADS) {
        uint32_t cm = 0, lm = 0;
#pragma unroll
        for (uint32_t w = 0; w < GLX_TILE_THREADS / 32; w++) {
            cm |= s_red[i >> 5][w];
            lm |= s_red[2 + (i >> 5)][w];
        }
        if (cm >> (i & 31) & 1) atomicOr(O.allele_used + site0 + i, 1u);  // fast cells only ever call allele 0
        if (lm >> (i & 31) & 1) {
            uint8_t* fp = O.site_flags + site0 + i;
            uint32_t* wd = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(fp) & ~(uintptr_t)3);
            atomicOr(wd, 1u << (8 * (reinterpret_cast<uintptr_t>(fp) & 3)));
        }
    }
}

#endif

// instantiated signatures (mode, number, cached values per field; field 0 in the low nibble)
//   DeepVariant*: DP(VEC,basic,1) AD(PATTERN,alleles,4) GQ(VEC,basic,1) PL2(PATTERN,genotype,4)
//   gatk*:        DP(VEC,basic,1) AD(PATTERN,alleles,4) SB(VEC,basic,4) GQ(VEC,basic,1) PL(CONST,genotype,0)
typedef TiledSig<4, 0x1010ull, 0x2030ull, 0x4141ull> SigDeepVariant;
typedef TiledSig<5, 0x20010ull, 0x20030ull, 0x01441ull> SigGatk;

}  // namespace glxd
