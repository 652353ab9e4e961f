// glx_tiled_sig.cuh — the tiled fast path specialised at compile time on the FIELD SIGNATURE of the config
// (how many lifted fields, and for each: FM_* mode, RetainedFieldNumber, cached-value count). With the signature a
// template parameter every per-field loop unrolls, a lane's cached band values live in registers, BASIC-numbered
// fields get their output offset arithmetically, and a cursor advance is four 128-bit loads. Semantics are those of
// glx_cell_fast.cuh (same blobs, same classification, same patterns); configs whose signature has no instantiation
// run the dynamic kernel there. Reference lines: see glx_cell_fast.cuh.
#pragma once
#include "glx_cell_fast.cuh"

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
    static constexpr uint32_t W = (GLX_BLOB_HDR + n_rv + 3u) & ~3u;
    static constexpr uint32_t NRV_ARR = n_rv ? n_rv : 1;
};

template <class Sig>
struct TiledLane {
    uint64_t base;
    uint32_t n, c;
    int beg_c, end_c, maxend_c, beg_n;
    uint32_t bits;
    int mrd;
    int32_t rv[Sig::NRV_ARR];

    // The lane's next records are staged in a private shared-memory ring by asynchronous copies (cp.async, no register
    // staging): records c .. c+RING-1 are always in flight or landed, so a cursor advance reads shared memory and
    // never waits on HBM, even when a 1-bp variant record makes a lane advance on consecutive sites.
    static constexpr uint32_t RING = 4, Q = Sig::W / 4;  // ring depth (records), 16-byte chunks per blob

    __device__ __forceinline__ void request(const int32_t* __restrict__ blobs, int4* ring, uint32_t rs, uint32_t idx) const {
        if (idx < n) {
            const int4* src = reinterpret_cast<const int4*>(blobs + (base + idx) * Sig::W);
#pragma unroll
            for (uint32_t q = 0; q < Q; q++) cp_async16(ring + (size_t)((idx & (RING - 1)) * Q + q) * rs, src + q);
        }
        cp_async_commit();  // one group per record index, empty past the end, so group counting stays uniform
    }
    __device__ __forceinline__ void consume(const int4* ring, uint32_t rs) {
        if (c < n) {
            cp_async_wait_but<RING - 1>();
            int32_t w[Sig::W];
#pragma unroll
            for (uint32_t q = 0; q < Q; q++) {
                const int4 v = ring[(size_t)((c & (RING - 1)) * Q + q) * rs];
                w[4 * q] = v.x;
                w[4 * q + 1] = v.y;
                w[4 * q + 2] = v.z;
                w[4 * q + 3] = v.w;
            }
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
    // (re)start the ring at record c
    __device__ __forceinline__ void load(const int32_t* __restrict__ blobs, int4* ring, uint32_t rs) {
        cp_async_wait_but<0>();
#pragma unroll
        for (uint32_t d = 0; d < RING; d++) request(blobs, ring, rs, c + d);
        consume(ring, rs);
    }
    __device__ __forceinline__ void advance(const int32_t* __restrict__ blobs, int4* ring, uint32_t rs) {
        c++;
        request(blobs, ring, rs, c + RING - 1);
        consume(ring, rs);
    }
    __device__ __forceinline__ void init(const DRecs& R, const int32_t* __restrict__ blobs, int4* ring, uint32_t rs, uint32_t sample, int qbeg) {
        base = R.ds_rec_off[sample];
        const uint64_t hi = R.ds_rec_off[sample + 1];
        n = (uint32_t)(hi - base);
        c = (uint32_t)(first_candidate(R.maxend, base, hi, qbeg) - base);
        load(blobs, ring, rs);
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
    __device__ __forceinline__ void seek(const int32_t* __restrict__ blobs, int4* ring, uint32_t rs, int qbeg, bool back) {
        if (back) {
            bool moved = false;
            while (c > 0 && blobs[(base + c - 1) * Sig::W + 2] > qbeg) {
                c--;
                moved = true;
            }
            if (moved) load(blobs, ring, rs);
        }
        while (c < n && maxend_c <= qbeg) advance(blobs, ring, rs);
    }
    // same decision table as fast_classify
    __device__ __forceinline__ int classify(const DCfg& cfg, const int32_t* __restrict__ blobs, int4* ring, uint32_t rs, int sbeg, int send, int qbeg, int qend,
                                            bool slow_site, bool back, bool& called, int& rnc) {
        called = false;
        rnc = RNC_M;
        seek(blobs, ring, rs, qbeg, back);
        if (slow_site) return CT_SLOW;
        if (c >= n || beg_c >= qend) return CT_MISSING;
        if (beg_n < qend) return CT_SLOW;
        if (bits & RB_SLOW) return CT_SLOW;
        if (!(end_c > sbeg && beg_c < send)) return CT_MISSING;
        if (!cfg.allow_partial && !(beg_c <= sbeg && end_c >= send)) {
            rnc = RNC_P;
            return CT_PARTIAL;
        }
        if (bits & RB_NONCALLED) rnc = RNC_I;
        else if ((unsigned long long)(long long)mrd >= (unsigned long long)cfg.required_dp) {
            called = true;
            rnc = RNC_NA;
        } else
            rnc = RNC_D;
        return CT_REF;
    }
    // field K's vector of a finished cell. A = unified alleles of the site; off_dyn(k) = fld_off[k][s] for
    // allele/genotype-numbered fields (BASIC-numbered fields sit at s*N*count).
    template <uint32_t K, class OffFn>
    __device__ __forceinline__ void emit_one(const FastPlan& plan, const DOut& O, bool ref, uint32_t s, uint32_t N, uint32_t sample, uint32_t A, OffFn off_dyn) const {
        const int32_t cen = plan.cen[K];
        constexpr uint32_t r0 = Sig::rv0(K);
        if constexpr (Sig::mode(K) == FM_VEC) {
            constexpr uint32_t cnt = Sig::nrv(K);
            int32_t* __restrict__ o = O.fld[K] + ((unsigned long long)s * N + sample) * cnt;
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
            int32_t* __restrict__ o = O.fld[K] + off + (unsigned long long)sample * count;
            int32_t x0 = cen, x1 = plan.vec_rule[K] ? GLXD_INT_VEND : cen, xr = cen, xt = cen;
            if constexpr (Sig::mode(K) == FM_PATTERN) {
                if (ref) {
                    x0 = rv[r0];
                    x1 = rv[r0 + 1];
                    xr = rv[r0 + 2];
                    xt = rv[r0 + 3];
                }
            }
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
    __device__ __forceinline__ void emit_from(const FastPlan& plan, const DOut& O, bool ref, uint32_t s, uint32_t N, uint32_t sample, uint32_t A, OffFn off_dyn) const {
        if constexpr (K < Sig::nf) {
            emit_one<K>(plan, O, ref, s, N, sample, A, off_dyn);
            emit_from<K + 1>(plan, O, ref, s, N, sample, A, off_dyn);
        }
    }
    template <class OffFn>
    __device__ __forceinline__ void emit(const FastPlan& plan, const DOut& O, bool ref, uint32_t s, uint32_t N, uint32_t sample, uint32_t A, OffFn off_dyn) const {
        emit_from<0>(plan, O, ref, s, N, sample, A, off_dyn);
    }
};

#ifdef __CUDACC__
template <class Sig>
__global__ void __launch_bounds__(GLX_TILE_THREADS) glx_cells_tiled_sig_kernel(const DCfg cfg, const FastPlan plan, const DRecs R, const DSites S, const DOut O,
                                                                               const int32_t* __restrict__ blobs, uint2* __restrict__ worklist,
                                                                               uint32_t* __restrict__ n_work, uint32_t n_sample_blocks) {
    __shared__ int4 s_pos[GLX_TILE_SITES];  // beg, end, qbeg, qend
    __shared__ uint32_t s_meta[GLX_TILE_SITES];  // bits 0-7 n_allele, bit 8 general path only, bit 9 qbeg below the previous site's
    __shared__ unsigned long long s_off[Sig::nf ? Sig::nf : 1][GLX_TILE_SITES];
    __shared__ uint32_t s_used[GLX_TILE_SITES], s_lost[GLX_TILE_SITES];
    __shared__ uint32_t s_nslow;  // cells this CTA handed to the general kernel (its worklist region is CTA-private)
    __shared__ int4 s_ring[TiledLane<Sig>::RING * TiledLane<Sig>::Q * GLX_TILE_THREADS];  // per-lane record ring, chunk-major

    const uint32_t sb = blockIdx.x % n_sample_blocks, tile = blockIdx.x / n_sample_blocks;
    const uint32_t site0 = tile * GLX_TILE_SITES;
    uint2* __restrict__ my_work = worklist + (size_t)blockIdx.x * (GLX_TILE_SITES * GLX_TILE_THREADS);
    if (threadIdx.x == 0) s_nslow = 0;
    const uint32_t nsite = min((uint32_t)GLX_TILE_SITES, S.n_sites - site0);
    const uint32_t N = S.n_samples;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t sample = sb * GLX_TILE_THREADS + tid;
    const bool active = sample < N;

    for (uint32_t i = tid; i < nsite; i += GLX_TILE_THREADS) {
        const uint32_t s = site0 + i;
        const int qb = S.qbeg[s];
        s_pos[i] = make_int4(S.beg[s], S.end[s], qb, S.qend[s]);
        s_meta[i] = (uint32_t)S.n_allele[s] | (S.monoallelic[s] ? 256u : 0u) | ((i > 0 && qb < S.qbeg[s - 1]) ? 512u : 0u);
        s_used[i] = 0;
        s_lost[i] = 0;
#pragma unroll
        for (uint32_t k = 0; k < Sig::nf; k++)
            if (Sig::mode(k) != FM_VEC) s_off[k][i] = S.fld_off[k][s];
    }
    __syncthreads();

    TiledLane<Sig> L;
    int4* ring = s_ring + tid;
    if (active) L.init(R, blobs, ring, GLX_TILE_THREADS, sample, s_pos[0].z);
    else L.idle();

    for (uint32_t i = 0; i < nsite; i++) {
        const uint32_t s = site0 + i;
        const int4 pos = s_pos[i];
        const uint32_t meta = s_meta[i];
        int type = CT_MISSING, rnc = RNC_M;
        bool called = false;
        if (active) type = L.classify(cfg, blobs, ring, GLX_TILE_THREADS, pos.x, pos.y, pos.z, pos.w, meta & 256u, meta & 512u, called, rnc);
        const bool slow = active && type == CT_SLOW;
        const bool done = active && !slow;
        const uint32_t slow_mask = __ballot_sync(0xffffffffu, slow);
        if (slow_mask) {
            uint32_t wbase = 0;
            const int leader = __ffs(slow_mask) - 1;
            if ((int)lane == leader) wbase = atomicAdd(&s_nslow, (uint32_t)__popc(slow_mask));
            wbase = __shfl_sync(0xffffffffu, wbase, leader);
            if (slow) my_work[wbase + __popc(slow_mask & ((1u << lane) - 1))] = make_uint2(s * N + sample, L.c | ((!(meta & 256u) && L.c < L.n && L.beg_c < pos.w && L.beg_n >= pos.w) ? GLX_WORK_SINGLE : 0u));
        }
        const uint32_t called_mask = __ballot_sync(0xffffffffu, done && called);
        const uint32_t lost_mask = __ballot_sync(0xffffffffu, done && rnc == RNC_I);
        if (lane == 0) {
            if (called_mask) atomicOr(&s_used[i], 1u);
            if (lost_mask) atomicOr(&s_lost[i], 1u);
        }
        if (!done) continue;
        const size_t cell = (size_t)s * N + sample;
        const signed char a = called ? 0 : -1;
        const unsigned char rc = (unsigned char)kRncChar[rnc];
        reinterpret_cast<char2*>(O.gt)[cell] = make_char2(a, a);
        reinterpret_cast<uchar2*>(O.rnc)[cell] = make_uchar2(rc, rc);
        L.emit(plan, O, type == CT_REF, s, N, sample, meta & 255u, [&](uint32_t k) { return s_off[k][i]; });
    }
    __syncthreads();
    if (tid == 0) n_work[blockIdx.x] = s_nslow;
    for (uint32_t i = tid; i < nsite; i += GLX_TILE_THREADS) {
        if (s_used[i]) atomicOr(O.allele_used + site0 + i, s_used[i]);
        if (s_lost[i]) {
            uint8_t* fp = O.site_flags + site0 + i;
            uint32_t* w = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(fp) & ~(uintptr_t)3);
            atomicOr(w, 1u << (8 * (reinterpret_cast<uintptr_t>(fp) & 3)));
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
