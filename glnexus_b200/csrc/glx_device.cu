// glx_device.cu — CUDA side of the C ABI in include/glx.h: device context, batch upload, kernels, download.
//
// Replaces, for one contig-range batch, the per-site task loop of Service::genotype_sites
// (src/service.cc:344-381) and everything genotype_site does per (site, sample) cell
// (src/genotyper.cc:682-867). sm_100a only; there is no host fallback: every entry point fails with
// GLX_E_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/glx.h"
#include "glx_cell_fast.cuh"
#include "glx_cell_general.cuh"
#include "glx_cell_variant.cuh"
#include "glx_dev_types.cuh"
#include "glx_tiled_sig.cuh"
#include "glx_wire.cuh"

using namespace glxd;

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------

// Once per upload: the device-private flag bit of each dataset's last record, and the packed site / unification rows
// (DSites::site_rows, unif_rows; same layout as build_site_rows) from the uploaded columns.
__global__ void __launch_bounds__(256) glx_prepare_inputs_kernel(const DRecs R, const DSites S, uint32_t n_datasets, uint64_t n_unif, uint8_t* __restrict__ flags,
                                                                 int4* __restrict__ site_rows, int4* __restrict__ unif_rows) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_datasets && R.ds_rec_off[i + 1] > R.ds_rec_off[i]) flags[R.ds_rec_off[i + 1] - 1] |= GLX_RF_DEV_LAST;
    if (i < S.n_sites) {
        site_rows[2 * i] = make_int4(S.beg[i], S.end[i], S.qbeg[i], S.qend[i]);
        site_rows[2 * i + 1] = make_int4((int)S.n_allele[i] | (S.monoallelic[i] ? 256 : 0), (int)S.allele_off[i], (int)S.unif_off[i], (int)S.unif_off[i + 1]);
    } else if (i == S.n_sites) {
        site_rows[2 * i] = site_rows[2 * i + 1] = make_int4(0, 0, 0, 0);
    }
    if (i < n_unif) {
        unif_rows[2 * i] = make_int4(S.unif_beg[i], S.unif_end[i], (int)S.unif_len[i], (int)S.unif_to[i]);
        unif_rows[2 * i + 1] = make_int4((int)(uint32_t)S.unif_prefix[i], (int)(uint32_t)(S.unif_prefix[i] >> 32), (int)S.unif_str[i], 0);
    } else if (i == n_unif) {
        unif_rows[2 * i] = unif_rows[2 * i + 1] = make_int4(0, 0, 0, 0);
    }
}

// Thread per record: resolve each record into its blob (glx_cell_fast.cuh). Part of every glx_launch. Blobs are built
// in shared memory and written out by the whole block as one contiguous, fully coalesced span. The same threads fill
// the tile cursor table (fill_tile_cursors states what is written): record r claims the tiles [t_lo, t_hi) with
// t_hi = first tile whose tq >= maxend[r] and t_lo = the previous record's t_hi, so each thread runs one search and
// takes t_lo from its neighbour; the per-block hints of glx_block_hints_kernel confine the searches (when the whole block
// lies in one sample, maxend is non-decreasing across it, so t_hi lies between this block's hint and the next one's).
#define GLX_RESOLVE_THREADS 256
__device__ __forceinline__ uint32_t first_tile_at_or_after(const int32_t* __restrict__ tq, uint32_t lo, uint32_t hi, int v) {
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tq[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
__global__ void __launch_bounds__(GLX_RESOLVE_THREADS) glx_resolve_records_kernel(const DCfg cfg, const FastPlan plan, const DRecs R, uint32_t n_datasets,
                                                                                  int32_t* __restrict__ blobs, const int32_t* __restrict__ tile_q,
                                                                                  uint32_t n_tiles, uint32_t* __restrict__ tile_cursors,
                                                                                  const uint2* __restrict__ hints) {
    extern __shared__ int32_t s_blob[];  // [GLX_RESOLVE_THREADS][W + 1]: rows padded by one word so that column writes spread over the banks
    __shared__ uint32_t s_thi[GLX_RESOLVE_THREADS];
    const uint32_t W = blob_words(plan.n_rv), WP = W + 1;
    const uint32_t tid = threadIdx.x;
    const uint64_t r0 = (uint64_t)blockIdx.x * GLX_RESOLVE_THREADS;
    const uint64_t r = r0 + tid;
    const uint64_t n_here = min((uint64_t)GLX_RESOLVE_THREADS, R.n_records - r0);
    if (r < R.n_records) build_record_blob(cfg, plan, R, r, R.flags[r] & GLX_RF_DEV_LAST, s_blob + tid * WP);  // glx_upload marks each sample's last record
    __syncthreads();
#ifndef GLX_ABLATE_CURSORS
    uint32_t sample = 0, t_hi = 0;
    uint64_t base = 0;
    if (r < R.n_records) {
        // (sample, t_hi) of this block's and the next block's first record (glx_block_hints_kernel) bound both searches
        const uint2 h0 = hints[blockIdx.x], h1 = hints[blockIdx.x + 1];
        uint32_t lo = h0.x, hi = h1.x + 1;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (R.ds_rec_off[mid] <= r) lo = mid;
            else hi = mid;
        }
        sample = lo;
        base = R.ds_rec_off[lo];
        const bool one_sample = h0.x == h1.x;
        t_hi = first_tile_at_or_after(tile_q, one_sample ? h0.y : 0u, one_sample ? h1.y : n_tiles, R.maxend[r]);
        s_thi[tid] = t_hi;
    }
    __syncthreads();
    if (r < R.n_records) {
        uint32_t t_lo = 0;
        if (r > base) t_lo = tid > 0 ? s_thi[tid - 1] : first_tile_at_or_after(tile_q, 0, t_hi, R.maxend[r - 1]);
        for (uint32_t tile = t_lo; tile < t_hi; tile++) tile_cursors[(size_t)tile * n_datasets + sample] = (uint32_t)(r - base);
    }
#endif
    // write-out: each warp stores whole rows, a 16-byte quad per lane
    int32_t* __restrict__ dst = blobs + r0 * W;
    const uint32_t qpr = W / 4, lane = tid & 31, rows_per_pass = 32 / qpr;
    const uint32_t lrow = lane / qpr, lquad = lane - lrow * qpr;
    if (lrow < rows_per_pass)
        for (uint32_t row = (tid >> 5) * rows_per_pass + lrow; row < n_here; row += (GLX_RESOLVE_THREADS / 32) * rows_per_pass) {
            const int32_t* src = s_blob + row * WP + lquad * 4;
            *reinterpret_cast<int4*>(dst + (size_t)row * W + lquad * 4) = make_int4(src[0], src[1], src[2], src[3]);
        }
}

// Once per upload: for every block of GLX_RESOLVE_THREADS records, the sample of its first record and that record's
// t_hi (first tile whose tq >= its maxend); entry n_blocks closes the table with (last sample, n_tiles). The resolve
// kernel's per-record searches then run inside [hint[b], hint[b + 1]] instead of over all samples / all tiles.
// suffix minimum of the tiles' first query positions (one block; tiles are few: sites / 32..64)
__global__ void __launch_bounds__(1024) glx_tile_q_kernel(const int32_t* __restrict__ qbeg, uint32_t n_tiles, uint32_t tile_sites, int32_t* __restrict__ tile_q) {
    __shared__ int32_t s_min[1024];
    const uint32_t t = threadIdx.x, chunk = (n_tiles + 1023u) / 1024u;
    const uint32_t lo = min(n_tiles, t * chunk), hi = min(n_tiles, lo + chunk);
    int32_t run = INT32_MAX;
    for (uint32_t i = hi; i-- > lo;) {
        run = min(run, qbeg[(size_t)i * tile_sites]);
        tile_q[i] = run;
    }
    s_min[t] = run;
    __syncthreads();
    int32_t later = INT32_MAX;
    for (uint32_t u = t + 1; u < 1024; u++) later = min(later, s_min[u]);
    for (uint32_t i = lo; i < hi; i++) tile_q[i] = min(tile_q[i], later);
}

__global__ void __launch_bounds__(256) glx_block_hints_kernel(const DRecs R, uint32_t n_datasets, const int32_t* __restrict__ tile_q, uint32_t n_tiles,
                                                              uint2* __restrict__ hints, uint32_t n_blocks) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_blocks) return;
    if (b == n_blocks) {
        hints[b] = make_uint2(n_datasets ? n_datasets - 1 : 0u, n_tiles);
        return;
    }
    const uint64_t r = (uint64_t)b * GLX_RESOLVE_THREADS;
    uint32_t lo = 0, hi = n_datasets;  // invariant: ds_rec_off[lo] <= r < ds_rec_off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (R.ds_rec_off[mid] <= r) lo = mid;
        else hi = mid;
    }
    hints[b] = make_uint2(lo, first_tile_at_or_after(tile_q, 0, n_tiles, R.maxend[r]));
}

// The same, for a config whose field signature is instantiated (the preset signatures) and whose plan has record views:
// the row is built in registers by the unrolled builder and stored as four 16-byte vectors; the tile-cursor searches are
// confined and exchanged inside the warp (shuffles), so the kernel has no shared memory and no block barrier.
template <class Sig>
__global__ void __launch_bounds__(GLX_RESOLVE_THREADS) glx_resolve_sig_kernel(const DCfg cfg, const FastPlan plan, const DRecs R, uint32_t n_datasets,
                                                                              int32_t* __restrict__ blobs, const int32_t* __restrict__ tile_q, uint32_t n_tiles,
                                                                              uint32_t* __restrict__ tile_cursors, const uint2* __restrict__ hints) {
    const uint64_t r = (uint64_t)blockIdx.x * GLX_RESOLVE_THREADS + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool valid = r < R.n_records;
    const uint64_t rr = valid ? r : R.n_records - 1;  // idle lanes shadow the last record (they take part in the shuffles)
    int32_t w[Sig::W];
    const uint8_t fl = R.flags[rr];
    build_record_blob_sig<Sig>(cfg, plan, R, rr, fl & GLX_RF_DEV_LAST, w);
    if (valid) {
        int4* dst = reinterpret_cast<int4*>(blobs + r * Sig::W);
#pragma unroll
        for (uint32_t q = 0; q < Sig::W / 4; q++) dst[q] = make_int4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
#ifndef GLX_ABLATE_CURSORS
    // sample of the record and its t_hi (first tile whose tq >= maxend[r]): searched inside the block's hint range.
    // Inside one sample maxend is non-decreasing, so t_hi lies between the hints of this block and the next.
    const uint2 h0 = hints[blockIdx.x], h1 = hints[blockIdx.x + 1];
    uint32_t lo = h0.x, hi = h1.x + 1;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (R.ds_rec_off[mid] <= rr) lo = mid;
        else hi = mid;
    }
    const uint32_t sample = lo;
    const uint64_t base = R.ds_rec_off[sample];
    const bool one_sample = h0.x == h1.x;
    const uint32_t t_hi = first_tile_at_or_after(tile_q, one_sample ? h0.y : 0u, one_sample ? h1.y : n_tiles, w[2]);
    uint32_t t_lo = __shfl_up_sync(0xffffffffu, t_hi, 1);  // the previous record's t_hi
    if (valid) {
        if (r == base) t_lo = 0;
        else if (lane == 0)  // the previous record is another warp's (for thread 0: another block's, below this block's hint)
            t_lo = first_tile_at_or_after(tile_q, (one_sample && threadIdx.x) ? h0.y : 0u, t_hi, R.maxend[r - 1]);
        for (uint32_t tile = t_lo; tile < t_hi; tile++) tile_cursors[(size_t)tile * n_datasets + sample] = (uint32_t)(r - base);
    }
#endif
}

// Tiled fast path. CTA = 128 lanes = 128 consecutive samples; it walks GLX_TILE_SITES consecutive sites. Lane state:
// a cursor into its sample's records and the resolved values of the band under the cursor (shared memory, one column
// per lane). Per site each lane classifies its cell (fast_classify) and writes a finished cell's GT, RNC and lifted
// vectors; consecutive lanes are consecutive samples, so every store instruction of a warp covers one contiguous
// span of the sample-major output. Unfinished cells go to the worklist of the general kernel with their cursor.
__global__ void __launch_bounds__(GLX_TILE_THREADS) glx_cells_tiled_kernel(const DCfg cfg, const FastPlan plan, const DRecs R, const DSites S, const DOut O,
                                                                           const int32_t* __restrict__ blobs, const uint32_t* __restrict__ tile_cursors,
                                                                           uint2* __restrict__ worklist, uint32_t* __restrict__ n_work, uint32_t n_sample_blocks,
                                                                           uint32_t cta0) {
    __shared__ int s_beg[GLX_TILE_SITES], s_end[GLX_TILE_SITES], s_qbeg[GLX_TILE_SITES], s_qend[GLX_TILE_SITES];
    __shared__ uint8_t s_flags[GLX_TILE_SITES];  // bit0: general path only (monoallelic), bit1: qbeg below the previous site's
    __shared__ uint16_t s_cnt[GLX_MAX_FIELDS][GLX_TILE_SITES];
    __shared__ unsigned long long s_off[GLX_MAX_FIELDS][GLX_TILE_SITES];
    __shared__ uint32_t s_used[GLX_TILE_SITES], s_lost[GLX_TILE_SITES];
    __shared__ uint32_t s_nslow;  // cells this CTA handed to the general kernel (its worklist region is CTA-private)
    extern __shared__ int32_t s_rv[];  // [plan.n_rv][GLX_TILE_THREADS]

    const uint32_t cta = blockIdx.x + cta0;
    const uint32_t sb = cta % n_sample_blocks, tile = cta / n_sample_blocks;
    const uint32_t site0 = tile * GLX_TILE_SITES;
    uint2* __restrict__ my_work = worklist + (size_t)cta * (GLX_TILE_SITES * GLX_TILE_THREADS);
    if (threadIdx.x == 0) s_nslow = 0;
    const uint32_t nsite = min((uint32_t)GLX_TILE_SITES, S.n_sites - site0);
    const uint32_t N = S.n_samples;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t sample = sb * GLX_TILE_THREADS + tid;
    const bool active = sample < N;
    const uint32_t n_fields = cfg.n_fields;

    for (uint32_t i = tid; i < nsite; i += GLX_TILE_THREADS) {
        const uint32_t s = site0 + i;
        s_beg[i] = S.beg[s];
        s_end[i] = S.end[s];
        const int qb = S.qbeg[s];
        s_qbeg[i] = qb;
        s_qend[i] = S.qend[s];
        s_flags[i] = (S.monoallelic[s] ? 1 : 0) | ((i > 0 && qb < S.qbeg[s - 1]) ? 2 : 0);
        s_used[i] = 0;
        s_lost[i] = 0;
        for (uint32_t k = 0; k < n_fields; k++) {
            s_cnt[k][i] = S.fld_cnt[k][s];
            s_off[k][i] = S.fld_off[k][s];
        }
    }
    __syncthreads();

    const uint32_t n_rv = plan.n_rv, W = blob_words(n_rv);
    int32_t* rv = s_rv + tid;
    LaneCursor L;
    L.base = 0;
    L.n = L.c = 0;
    L.beg_c = L.end_c = L.maxend_c = L.beg_n = 0x7fffffff;
    L.bits = 1;
    L.mrd = -1;
    if (active) cursor_init(R, blobs, W, n_rv, L, sample, tile_cursors[(size_t)tile * N + sample], rv, GLX_TILE_THREADS);

    for (uint32_t i = 0; i < nsite; i++) {
        const uint32_t s = site0 + i;
        int type = CT_MISSING, rnc = RNC_M;
        bool called = false;
        if (active) {
            const uint32_t fl = s_flags[i];
            type = fast_classify(cfg, blobs, W, n_rv, L, s_beg[i], s_end[i], s_qbeg[i], s_qend[i], fl & 1, fl & 2, rv, GLX_TILE_THREADS, called, rnc);
        }
        const bool slow = active && type == CT_SLOW;
        const bool done = active && !slow;
        // ---- hand unfinished cells to the general kernel (warp-aggregated append) ----
        const uint32_t slow_mask = __ballot_sync(0xffffffffu, slow);
        if (slow_mask) {
            uint32_t base = 0;
            const int leader = __ffs(slow_mask) - 1;
            if ((int)lane == leader) base = atomicAdd(&s_nslow, (uint32_t)__popc(slow_mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (slow) my_work[base + __popc(slow_mask & ((1u << lane) - 1))] =
                    make_uint2(s * N + sample, L.c | ((!(s_flags[i] & 1) && L.c < L.n && L.beg_c < s_qend[i] && L.beg_n >= s_qend[i]) ? GLX_WORK_SINGLE : 0u));
        }
        const uint32_t called_mask = __ballot_sync(0xffffffffu, done && called);
        const uint32_t lost_mask = __ballot_sync(0xffffffffu, done && rnc == RNC_I);
        if (lane == 0) {
            if (called_mask) atomicOr(&s_used[i], 1u);  // fast cells only ever call allele 0
            if (lost_mask) atomicOr(&s_lost[i], 1u);
        }
        if (!done) continue;
        // ---- GT / RNC ----
        const size_t cell = (size_t)s * N + sample;
        const signed char a = called ? 0 : -1;
        const unsigned char rc = (unsigned char)kRncChar[rnc];
        reinterpret_cast<char2*>(O.gt)[cell] = make_char2(a, a);
        reinterpret_cast<uchar2*>(O.rnc)[cell] = make_uchar2(rc, rc);
        // ---- lifted FORMAT fields ----
        const bool ref = type == CT_REF;
#pragma unroll
        for (uint32_t k = 0; k < GLX_MAX_FIELDS; k++) {  // unrolled: plan/O indices become static (kernel-parameter constant bank)
            if (k < n_fields) {
                const int count = s_cnt[k][i];
                fast_emit_field(plan, k, ref, count, O.fld[k] + s_off[k][i] + (unsigned long long)sample * count, rv, GLX_TILE_THREADS);
            }
        }
    }
    __syncthreads();
    if (tid == 0) n_work[cta] = s_nslow;
    for (uint32_t i = tid; i < nsite; i += GLX_TILE_THREADS) {
        if (s_used[i]) atomicOr(O.allele_used + site0 + i, s_used[i]);
        if (s_lost[i]) {
            uint8_t* fp = O.site_flags + site0 + i;
            uint32_t* w = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(fp) & ~(uintptr_t)3);
            atomicOr(w, 1u << (8 * (reinterpret_cast<uintptr_t>(fp) & 3)));
        }
    }
}

// Worklist, pass 1: thread per worklist cell flagged "exactly one record in the query range". The closed form for a lone
// variant record (genotype_cell_single) needs no scratch and few registers, so this kernel runs at high occupancy; a
// cell it declines has its flag cleared in place and is left to the general kernel.
__global__ void __launch_bounds__(128) glx_cells_single_kernel(const DCfg cfg, const DRecs R, const DSites S, const DOut O, uint2* __restrict__ worklist,
                                                               const uint32_t* __restrict__ n_work, uint32_t reg0, uint32_t n_regions, uint32_t region_cells) {
    const uint32_t N = S.n_samples;
    for (uint32_t reg = reg0 + blockIdx.x; reg < n_regions; reg += gridDim.x) {
        const uint32_t nw = n_work[reg];
        const uint32_t n_items = nw & GLX_NWORK_MASK;
        if ((nw >> 16) == n_items) continue;  // nothing left: the tiled kernel finished every entry of this region itself
        uint2* __restrict__ w = worklist + (size_t)reg * region_cells;
        for (uint32_t it = threadIdx.x; it < n_items; it += blockDim.x) {
            const uint2 e = w[it];
            if ((e.y & (GLX_WORK_SINGLE | GLX_WORK_DONE)) != GLX_WORK_SINGLE) continue;
            const uint32_t s = e.x / N, sample = e.x % N;
            const uint64_t first = R.ds_rec_off[sample] + (e.y & GLX_WORK_CURSOR);
            if (!genotype_cell_single(cfg, R, S, O, s, sample, first)) w[it].y = e.y & GLX_WORK_CURSOR;
        }
    }
}

// Worklist, pass 0: the same cells as pass 1, through the record's view (glx_cell_variant.cuh): all loads of the closed
// form issued up front. Finished cells get GLX_WORK_DONE; a record without a view is left to pass 1.
#ifndef GLX_VARIANT_MIN_BLOCKS
#define GLX_VARIANT_MIN_BLOCKS 5  // 96 registers, no spills (6: 80 registers with 72 B of spills, measured 2 % slower on the 8-allele instantiation)
#endif
template <int CAPA>
__global__ void __launch_bounds__(128, GLX_VARIANT_MIN_BLOCKS) glx_cells_variant_kernel(const DCfg cfg, const FastPlan plan, const DRecs R, const DSites S, const DOut O,
                                                                uint2* __restrict__ worklist, const uint32_t* __restrict__ n_work, uint32_t reg0, uint32_t n_regions,
                                                                uint32_t region_cells, const int32_t* __restrict__ blobs) {
    extern __shared__ int32_t s_cols[];  // [VariantSmem<CAPA>::SLOTS][blockDim.x]: one column per thread
    const uint32_t N = S.n_samples;
    for (uint32_t reg = reg0 + blockIdx.x; reg < n_regions; reg += gridDim.x) {
        const uint32_t nw = n_work[reg];
        const uint32_t n_items = nw & GLX_NWORK_MASK;
        if ((nw >> 16) == n_items) continue;  // nothing left: the tiled kernel finished every entry of this region itself
        uint2* __restrict__ w = worklist + (size_t)reg * region_cells;
        for (uint32_t it = threadIdx.x; it < n_items; it += blockDim.x) {
            const uint2 e = w[it];
            if ((e.y & (GLX_WORK_SINGLE | GLX_WORK_DONE)) != GLX_WORK_SINGLE) continue;
            const uint32_t s = e.x / N, sample = e.x % N;
            const uint64_t first = R.ds_rec_off[sample] + (e.y & GLX_WORK_CURSOR);
            if (genotype_cell_variant<CAPA>(cfg, plan, R, S, O, blobs, s, sample, first, s_cols + threadIdx.x, blockDim.x)) w[it].y = e.y | GLX_WORK_DONE;
        }
    }
}

// Worklist, pass 0 for the fused tiled kernel's declined list: one thread per listed entry (region, index), any region, the
// general closed form. A finished entry gets GLX_WORK_DONE and is added to its region's finished count (high half of n_work),
// so that the region passes skip regions with nothing left.
template <int CAPA>
__global__ void __launch_bounds__(128, GLX_VARIANT_MIN_BLOCKS) glx_cells_variant_list_kernel(const DCfg cfg, const FastPlan plan, const DRecs R, const DSites S, const DOut O,
                                                                uint2* __restrict__ worklist, uint32_t* __restrict__ n_work, uint32_t region_cells,
                                                                const int32_t* __restrict__ blobs) {
    extern __shared__ int32_t s_cols[];
    const uint32_t N = S.n_samples;
    const uint32_t offered = *plan.decl_n, n = offered < plan.decl_cap ? offered : plan.decl_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 at = plan.decl_list[i];
        uint2* w = worklist + (size_t)at.x * region_cells + at.y;
        const uint2 e = *w;
        if ((e.y & (GLX_WORK_SINGLE | GLX_WORK_DONE)) != GLX_WORK_SINGLE) continue;
        const uint32_t s = e.x / N, sample = e.x % N;
        const uint64_t first = R.ds_rec_off[sample] + (e.y & GLX_WORK_CURSOR);
        if (genotype_cell_variant<CAPA>(cfg, plan, R, S, O, blobs, s, sample, first, s_cols + threadIdx.x, blockDim.x)) {
            w->y = e.y | GLX_WORK_DONE;
            atomicAdd(n_work + at.x, 1u << 16);
            atomicAdd(plan.decl_n + GLX_MAX_CHUNKS, 1u);  // cells this kernel finished (glx_fused_count leaves them out)
        }
    }
}

// Worklist, pass 2: the unflagged cells (several records in the query range) try the closed form for a run of
// reference bands (genotype_cell_bands, computed from the record blobs); finished cells get GLX_WORK_DONE.
__global__ void __launch_bounds__(128) glx_cells_bands_kernel(const DCfg cfg, const FastPlan plan, const DRecs R, const DSites S, const DOut O,
                                                              uint2* __restrict__ worklist, const uint32_t* __restrict__ n_work, uint32_t reg0, uint32_t n_regions,
                                                              uint32_t region_cells, const int32_t* __restrict__ blobs) {
    const uint32_t N = S.n_samples;
    for (uint32_t reg = reg0 + blockIdx.x; reg < n_regions; reg += gridDim.x) {
        const uint32_t nw = n_work[reg];
        const uint32_t n_items = nw & GLX_NWORK_MASK;
        if ((nw >> 16) == n_items) continue;  // nothing left: the tiled kernel finished every entry of this region itself
        uint2* __restrict__ w = worklist + (size_t)reg * region_cells;
        for (uint32_t it = threadIdx.x; it < n_items; it += blockDim.x) {
            const uint2 e = w[it];
            if (e.y & (GLX_WORK_SINGLE | GLX_WORK_DONE)) continue;
            const uint32_t s = e.x / N, sample = e.x % N;
            const uint64_t lo = R.ds_rec_off[sample];
            if (genotype_cell_bands(cfg, plan, R, S, O, blobs, s, sample, lo + (e.y & GLX_WORK_CURSOR), R.ds_rec_off[sample + 1])) w[it].y = e.y | GLX_WORK_DONE;
        }
    }
}

__device__ __forceinline__ void pf_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// General path: one thread per cell. With a worklist, block b serves the CTA-private regions b, b+gridDim.x, ... that
// the tiled kernel filled (cell id + the lane's record cursor; region r holds n_work[r] entries) and takes the cells
// the single-record pass left; without one it takes every cell of the batch and finds the cursor by binary search.
__global__ void __launch_bounds__(128) glx_cells_general_kernel(const DCfg cfg, const DRecs R, const DSites S, const DOut O, const uint2* __restrict__ worklist,
                                                                const uint32_t* __restrict__ n_work, uint32_t reg0, uint32_t n_regions, uint32_t region_cells,
                                                                uint64_t n_all, uint8_t* __restrict__ cnt_scratch, uint32_t max_slots, uint32_t n_cols) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t* cnt = cnt_scratch + tid;
    const uint32_t N = S.n_samples;
    if (worklist) {
        for (uint32_t reg = reg0 + blockIdx.x; reg < n_regions; reg += gridDim.x) {
            const uint32_t nw = n_work[reg];
            const uint32_t n_items = nw & GLX_NWORK_MASK;
            if ((nw >> 16) == n_items) continue;  // nothing left: the tiled kernel finished every entry of this region itself
            const uint2* __restrict__ w = worklist + (size_t)reg * region_cells;
            for (uint32_t it = threadIdx.x; it < n_items; it += blockDim.x) {
                const uint2 e = w[it];
                if (e.y & (GLX_WORK_SINGLE | GLX_WORK_DONE)) continue;  // finished by glx_cells_single_kernel
                const uint32_t s = e.x / N, sample = e.x % N;
                const uint64_t lo = R.ds_rec_off[sample], hi = R.ds_rec_off[sample + 1];
#ifndef GLX_ABLATE_GENERAL_PREFETCH
                // The general form walks the cell's records one dependent load at a time; ask L2 for their lines first.
                for (uint64_t r = lo + (e.y & GLX_WORK_CURSOR), k = 0; r < hi && k < 4; r++, k++) {
                    pf_l2(R.beg + r); pf_l2(R.end + r); pf_l2(R.flags + r); pf_l2(R.n_allele + r); pf_l2(R.gt + r); pf_l2(R.allele_off + r);
                    pf_l2(R.qual + r); pf_l2(R.filt + r);
                    for (uint32_t c = 0; c < n_cols; c++) {
                        pf_l2(R.col_len + (size_t)c * R.n_records + r);
                        pf_l2(R.col_val + (size_t)c * R.n_records + r);
                    }
                }
#endif
                genotype_cell_general(cfg, R, S, O, s, sample, lo + (e.y & GLX_WORK_CURSOR), hi, cnt, stride);
            }
        }
        return;
    }
    for (uint64_t cell = tid; cell < n_all; cell += stride) {
        const uint32_t s = (uint32_t)(cell / N), sample = (uint32_t)(cell % N);
        const uint64_t lo = R.ds_rec_off[sample], hi = R.ds_rec_off[sample + 1];
        genotype_cell_general(cfg, R, S, O, s, sample, first_candidate(R.maxend, lo, hi, S.qbeg[s]), hi, cnt, stride);
    }
}

// Batches with a multi-sample dataset (or datasets out of output order): one thread per (site, output sample); the thread
// presents its sample of the dataset as a single-sample view (MRecs) to the general cell logic, whose MULTI branches restate what
// the reference decides per dataset across its samples. Rare inputs (the reference itself calls its multi-sample handling
// incomplete, src/genotyper.cc:362,595), so no fast path: correctness only.
__global__ void __launch_bounds__(128) glx_cells_multi_kernel(const DCfg cfg, const DRecs R0, const int32_t* __restrict__ gt_pool, const uint2* __restrict__ out_map,
                                                              const DSites S, const DOut O, uint64_t n_all, uint8_t* __restrict__ cnt_scratch) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t* cnt = cnt_scratch + tid;
    const uint32_t N = S.n_samples;
    MRecs R;
    static_cast<DRecs&>(R) = R0;
    R.gt_pool = gt_pool;
    for (uint64_t cell = tid; cell < n_all; cell += stride) {
        const uint32_t s = (uint32_t)(cell / N), sample = (uint32_t)(cell % N);
        const uint2 m = out_map[sample];
        uint64_t lo = 0, hi = 0;
        R.ns = 1;
        R.si = 0;
        uint32_t ds = sample;
        if (m.x != 0xFFFFFFFFu) {
            ds = m.x;
            lo = R.ds_rec_off[m.x];
            hi = R.ds_rec_off[m.x + 1];
            R.si = m.y & 0xFFFFu;
            R.ns = m.y >> 16;
        }
        const unsigned long long err_cell = (unsigned long long)s * N + (ds < N ? ds : N - 1);
        genotype_cell_general(cfg, R, S, O, s, sample, lo < hi ? first_candidate(R.maxend, lo, hi, S.qbeg[s]) : lo, hi, cnt, stride, err_cell);
    }
}

// int32 column -> int16 or int8 with BCF's sentinels of that width. *overflow: bit 0 some value does not fit int16,
// bit 1 some value does not fit int8 (both recorded whatever the output width: the next batch's width follows them).
template <class T>
__global__ void __launch_bounds__(256) glx_narrow_kernel(const int32_t* __restrict__ in, T* __restrict__ out, uint64_t n, uint32_t* __restrict__ overflow) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t bad = 0;
    auto conv = [&](int32_t v) -> uint32_t {  // the value's bits in T
        if (v == GLXD_INT_MISSING) return sizeof(T) == 1 ? 0x80u : 0x8000u;
        if (v == GLXD_INT_VEND) return sizeof(T) == 1 ? 0x81u : 0x8001u;
        if (v < -32760 || v > 32767) bad |= 1u;  // -32768..-32761 are reserved in BCF
        if (v < -120 || v > 127) bad |= 2u;      // -128..-121 likewise
        return sizeof(T) == 1 ? (uint32_t)(uint8_t)v : (uint32_t)(uint16_t)v;
    };
    const uint64_t n4 = n / 4;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int4 v = reinterpret_cast<const int4*>(in)[i];
        const uint32_t a = conv(v.x), b = conv(v.y), c = conv(v.z), d = conv(v.w);
        if (sizeof(T) == 1) reinterpret_cast<uint32_t*>(out)[i] = a | b << 8 | c << 16 | d << 24;
        else reinterpret_cast<uint2*>(out)[i] = make_uint2(a | b << 16, c | d << 16);
    }
    for (uint64_t i = n4 * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (T)conv(in[i]);
    const uint32_t any = __reduce_or_sync(0xffffffffu, bad);
    if (any && (threadIdx.x & 31) == 0) atomicOr(overflow, any);
}

// Known-answer aid (tests): the device's diploid index math, out[j(j+1)/2 + i] = alleles_gt(i, j) | alleles_gt(j, i) << 16
// for all i <= j < n (src/diploid.cc:12-50, test/diploid.cc:9-39).
__global__ void glx_selftest_diploid_kernel(uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t j = blockIdx.x, i = threadIdx.x;
    if (j < n && i <= j) out[j * (j + 1) / 2 + i] = (uint32_t)alleles_gt((int)i, (int)j) | (uint32_t)alleles_gt((int)j, (int)i) << 16;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#include "glx_dev_state.cuh"

// Pinned status / overflow words of a batch: a free slot of the context's pools, held until glx_free_batch; with 64
// batches alive the batch gets its own allocation (slow: cudaHostAlloc synchronises the device).
int glx_take_err_slot(glx_ctx* ctx, glx_dev_batch* d) {
    if (d->host_err) return GLX_OK;
    for (int i = 0; i < 64; i++)
        if (!(ctx->err_slots_used >> i & 1)) {
            ctx->err_slots_used |= 1ull << i;
            d->err_slot = i;
            d->host_err = ctx->host_err_pool + 8 * i;  // {error word, then the BCF encoder's totals: stream bytes, header bytes, BGZF bytes, BGZF blocks, records}
            d->host_overflow = ctx->host_ovf_pool + (size_t)i * GLX_MAX_FIELDS;
            return GLX_OK;
        }
    void* p = nullptr;
    if (cudaHostAlloc(&p, 64 + GLX_MAX_FIELDS * 4, cudaHostAllocDefault) != cudaSuccess) return GLX_E_CUDA;
    d->err_slot = -1;
    d->host_err = static_cast<unsigned long long*>(p);
    d->host_overflow = reinterpret_cast<uint32_t*>(static_cast<char*>(p) + 64);
    return GLX_OK;
}

extern "C" {

void* glx_pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void glx_pinned_free(void* p) { cudaFreeHost(p); }

int glx_create(glx_ctx** out, int device) {
    if (!out) return GLX_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return GLX_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return GLX_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GLX_E_CUDA;
    if (prop.major != 10) return GLX_E_CUDA;  // sm_100a binary only
    auto* ctx = new glx_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    memset(ctx->width_hint, 2, sizeof ctx->width_hint);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return GLX_E_CUDA;
    }
    for (auto& e : ctx->chunk_ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            delete ctx;
            return GLX_E_CUDA;
        }
    if (cudaHostAlloc((void**)&ctx->host_err_pool, 64 * 64, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->host_ovf_pool, 64 * GLX_MAX_FIELDS * 4, cudaHostAllocDefault) != cudaSuccess) {
        delete ctx;
        return GLX_E_CUDA;
    }
    *out = ctx;
    return GLX_OK;
}

void glx_destroy(glx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto& c : ctx->cache)
        if (c.p) cudaFree(c.p);
    for (auto& e : ctx->chunk_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->crc_pw) cudaFree(ctx->crc_pw);
    if (ctx->host_err_pool) cudaFreeHost(ctx->host_err_pool);
    if (ctx->host_ovf_pool) cudaFreeHost(ctx->host_ovf_pool);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* glx_last_error(const glx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int glx_synchronize(glx_ctx* ctx) {
    if (!ctx) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return GLX_OK;
}

void* glx_stream(glx_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

void glx_free_batch(glx_ctx* ctx, glx_dev_batch* d) {
    if (!d) return;
    if (ctx) cudaSetDevice(ctx->device);
    for (auto& ev : d->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : d->up_ev)
        if (ev) cudaEventDestroy(ev);
    if (d->host_err) {
        if (d->err_slot >= 0 && ctx) ctx->err_slots_used &= ~(1ull << d->err_slot);
        else if (d->err_slot < 0) cudaFreeHost(d->host_err);
    }

    if (d->bcf.len_ev) cudaEventDestroy(d->bcf.len_ev);
    if (d->bcf.meta_ev) cudaEventDestroy(d->bcf.meta_ev);
    if (ctx) {
        ctx->give(d->bcf.meta_base, d->bcf.meta_bytes);
        ctx->give(d->bcf.work_base, d->bcf.work_bytes);
        ctx->give(d->bcf.stream, d->bcf.stream_bytes);
        ctx->give(d->bcf.packed, d->bcf.packed_bytes);
        ctx->give(d->in_base, d->in_alloc);
        ctx->give(d->out_base, d->out_alloc);
        ctx->give(d->cnt_scratch, d->cnt_scratch_bytes);
        ctx->give(d->worklist, d->worklist_bytes);
        ctx->give(d->blobs, d->blob_bytes);
        ctx->give(d->narrow, d->narrow_bytes);
    } else {
        if (d->bcf.meta_base) cudaFree(d->bcf.meta_base);
        if (d->bcf.work_base) cudaFree(d->bcf.work_base);
        if (d->bcf.stream) cudaFree(d->bcf.stream);
        if (d->bcf.packed) cudaFree(d->bcf.packed);
        if (d->narrow) cudaFree(d->narrow);
        if (d->worklist) cudaFree(d->worklist);
        if (d->blobs) cudaFree(d->blobs);
        if (d->in_base) cudaFree(d->in_base);
        if (d->out_base) cudaFree(d->out_base);
        if (d->cnt_scratch) cudaFree(d->cnt_scratch);
    }
    delete d;
}

int glx_upload(glx_ctx* ctx, const glx_config* cfg, const glx_records* recs, const glx_sites* sites, glx_dev_batch** out) {
    return glx_upload_on(ctx, cfg, recs, sites, nullptr, 1, out);
}

}  // extern "C"

static inline int32_t* WO_pool_ptr(const DRecs& R) { return const_cast<int32_t*>(R.pool); }

// recs: the kernel-facing record store on the host (classic upload), or wire: its compact wire form (expanded on the device)
static int upload_impl(glx_ctx* ctx, const glx_config* cfg, const glx_records* recs, const glx_wire* wire, const glx_sites* sites, void* stream_, int sync,
                       glx_dev_batch** out) {
    if (!ctx) return GLX_E_INVALID;
    if (!cfg || (!recs && !wire) || !sites || !out) {
        ctx->err = "glx_upload: null argument";
        return GLX_E_INVALID;
    }
    glx_records wrec;  // the counts of a wire batch, in the shape the code below reads
    if (wire) {
        if (wire->magic != GLX_WIRE_MAGIC || wire->n_cols != cfg->n_cols) {
            ctx->err = "glx_upload_wire: not a wire buffer of this config";
            return GLX_E_INVALID;
        }
        memset(&wrec, 0, sizeof wrec);
        wrec.n_datasets = wrec.n_samples_out = wire->n_datasets;
        wrec.n_cols = wire->n_cols;
        wrec.n_records = wire->n_records;
        wrec.n_pool = wire->n_pool;
        wrec.n_al = wire->n_al;
        wrec.n_al_pool = wire->n_al_pool;
        recs = &wrec;
    }
    *out = nullptr;
    if (cfg->abi_version != GLX_ABI_VERSION) {
        ctx->err = "glx_upload: ABI version mismatch";
        return GLX_E_INVALID;
    }
    // preconditions of the device path: single-sample datasets in output-sample order
    if (recs->n_samples_out != sites->n_samples || (wire && recs->n_datasets != sites->n_samples)) {
        ctx->err = "glx_upload: the record store's output samples are not the sites' samples";
        return GLX_E_UNSUPPORTED;
    }
    // The fast paths want one single-sample dataset per output sample, in output order. Anything else (multi-sample datasets,
    // reordered or absent samples) takes the general multi-sample kernel, provided every sample of every dataset is in the sample
    // set (the reference's FORMAT helpers throw on a partly selected dataset, src/genotyper_utils.h:75) and none is there twice.
    bool multi = false;
    std::vector<uint2> out_map;
    if (!wire) {
        multi = recs->n_datasets != sites->n_samples;
        for (uint32_t d = 0; !multi && d < recs->n_datasets; d++)
            multi = recs->ds_n_sample[d] != 1 || recs->sample_out[recs->ds_sample_off[d]] != (int32_t)d;
    }
    if (multi) {
        out_map.assign(sites->n_samples, make_uint2(0xFFFFFFFFu, 0u));
        for (uint32_t d = 0; d < recs->n_datasets; d++) {
            const uint32_t ns = recs->ds_n_sample[d];
            if (ns == 0 || ns > 0xFFFFu || recs->ds_sample_off[d + 1] - recs->ds_sample_off[d] != ns) {
                ctx->err = "glx_upload: malformed dataset sample table";
                return GLX_E_INVALID;
            }
            for (uint32_t k = 0; k < ns; k++) {
                const int32_t o = recs->sample_out[recs->ds_sample_off[d] + k];
                if (o < 0 || (uint32_t)o >= sites->n_samples || out_map[o].x != 0xFFFFFFFFu) {
                    ctx->err = "glx_upload: a dataset sample outside the sample set (or selected twice) is outside the device path";
                    return GLX_E_UNSUPPORTED;
                }
                out_map[o] = make_uint2(d, k | ns << 16);
            }
        }
    }
    const uint64_t S = sites->n_sites, N = sites->n_samples;
    if (S * N >= (1ull << 32)) {
        ctx->err = "glx_upload: more than 2^32 cells in one batch; split the site range";
        return GLX_E_UNSUPPORTED;
    }
    if (cfg->n_fields > GLX_MAX_FIELDS || cfg->n_cols > GLX_MAX_COLS) {
        ctx->err = "glx_upload: config exceeds GLX_MAX_FIELDS / GLX_MAX_COLS";
        return GLX_E_INVALID;
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t up = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    auto* d = new glx_dev_batch();
    d->cfg = make_dcfg(*cfg);
    d->n_sites = (uint32_t)S;
    d->n_samples = (uint32_t)N;
    d->n_fields = cfg->n_fields;
    d->n_cols = cfg->n_cols;
    d->n_cells = S * N;
    d->force_general = ctx->opt.force_general != 0 || multi;
    d->multi = multi;
    d->out_map_host.swap(out_map);
    // chunked launches share the context's aux stream and events: only for batches on the context's own stream
    d->n_chunks = (stream_ && (cudaStream_t)stream_ != ctx->stream) ? 1u : (uint32_t)ctx->opt.chunks;

    const uint64_t Rn = recs->n_records, D = recs->n_datasets, C = cfg->n_cols;
    const uint64_t nA = sites->n_site_alleles, nU = sites->n_unif;
    // ---- size the input arena ----
    struct Item {
        const void* src;
        size_t bytes;
        void** dst;
    };
    std::vector<Item> items;
    auto add = [&](const void* src, size_t bytes, auto** dst) { items.push_back(Item{src, bytes, (void**)dst}); };
    // failure after work was queued: wait for it before the arenas return to the cache (another stream may take them)
    auto bail = [&](cudaError_t err, const char* what) {
        cudaStreamSynchronize(up);
        glx_free_batch(ctx, d);
        return cuda_fail(ctx, err, what);
    };
    DRecs& R = d->R;
    DSites& Sx = d->S;
    memset(&R, 0, sizeof R);
    memset(&Sx, 0, sizeof Sx);
    R.n_records = Rn;
    add(recs->ds_rec_off, (D + 1) * 8, &R.ds_rec_off);
    add(recs->beg, Rn * 4, &R.beg);
    add(recs->end, Rn * 4, &R.end);
    add(recs->maxend, Rn * 4, &R.maxend);
    add(recs->qual, Rn * 4, &R.qual);
    add(recs->n_allele, Rn, &R.n_allele);
    uint64_t n_variant_records = wire ? wire->n_variant : 0;
    for (uint64_t rr = 0; !wire && rr < Rn; rr++) n_variant_records += !(recs->flags[rr] & GLX_RF_IS_REF);
    add(recs->flags, Rn, &R.flags);  // glx_prepare_inputs_kernel adds the device-private "last record of its dataset" bit
    add(recs->filt, Rn * 4, &R.filt);
    add(recs->gt, Rn * 4, &R.gt);
    add(recs->allele_off, Rn * 4, &R.allele_off);
    add(recs->col_val, C * Rn * 4, &R.col_val);
    add(recs->col_len, C * Rn * 2, &R.col_len);
    add(recs->pool, recs->n_pool * 4, &R.pool);
    add(recs->al_len, recs->n_al * 4, &R.al_len);
    add(recs->al_flags, recs->n_al, &R.al_flags);
    add(recs->al_prefix, recs->n_al * 8, &R.al_prefix);
    add(recs->al_str, recs->n_al * 4, &R.al_str);
    add(recs->al_pool, recs->n_al_pool, &R.al_pool);
    if (multi) {
        add(recs->gt_pool, recs->n_gt_pool * 4, &d->gt_pool);
        add(d->out_map_host.data(), N * sizeof(uint2), &d->out_map);
    }
    Sx.n_sites = (uint32_t)S;
    Sx.n_samples = (uint32_t)N;
    add(sites->beg, S * 4, &Sx.beg);
    add(sites->end, S * 4, &Sx.end);
    add(sites->qbeg, S * 4, &Sx.qbeg);
    add(sites->qend, S * 4, &Sx.qend);
    add(sites->n_allele, S, &Sx.n_allele);
    add(sites->monoallelic, S, &Sx.monoallelic);
    add(sites->allele_off, (S + 1) * 4, &Sx.allele_off);
    add(sites->allele_is_indel, nA, &Sx.allele_is_indel);
    add(sites->hom_log_prior, nA * 4, &Sx.hom_log_prior);
    add(sites->lost_log_prior, S * 4, &Sx.lost_log_prior);
    add(sites->unif_off, (S + 1) * 4, &Sx.unif_off);
    add(sites->unif_beg, nU * 4, &Sx.unif_beg);
    add(sites->unif_end, nU * 4, &Sx.unif_end);
    add(sites->unif_to, nU, &Sx.unif_to);
    add(sites->unif_len, nU * 4, &Sx.unif_len);
    add(sites->unif_prefix, nU * 8, &Sx.unif_prefix);
    add(sites->unif_str, nU * 4, &Sx.unif_str);
    add(sites->unif_pool, sites->n_unif_pool, &Sx.unif_pool);
    // packed on the device from the columns above (every source of the upload stays caller memory: a pageable staging
    // copy in the middle of the stream would make the call wait for the copies queued before it)
    add((const void*)nullptr, (S * 2 + 2) * sizeof(int4), &Sx.site_rows);
    add((const void*)nullptr, (nU * 2 + 2) * sizeof(int4), &Sx.unif_rows);
    for (uint32_t k = 0; k < cfg->n_fields; k++) {
        add(sites->fld_cnt[k], S * 2, &Sx.fld_cnt[k]);
        add(sites->fld_off[k], (S + 1) * 8, &Sx.fld_off[k]);
        d->fld_total[k] = S ? sites->fld_off[k][S] : 0;
    }
    char* wire_dev = nullptr;
    if (wire) {  // the wire buffer rides in the same arena; the record-store arrays above have no host source and are expanded from it
        add((const void*)wire, (size_t)wire->n_bytes, &wire_dev);
        const size_t nb = (Rn + GLX_WIRE_BLOCK * GLX_WIRE_ITEMS - 1) / (GLX_WIRE_BLOCK * GLX_WIRE_ITEMS) + 1;
        const size_t nab = (recs->n_al + GLX_WIRE_BLOCK * GLX_WIRE_ITEMS - 1) / (GLX_WIRE_BLOCK * GLX_WIRE_ITEMS) + 1;
        add((const void*)nullptr, nb * sizeof(uint3x), &d->wire_sums);
        add((const void*)nullptr, nab * sizeof(uint3x), &d->wire_al_sums);
        if (wire->flags & GLX_WF_SHAPES) add((const void*)nullptr, (size_t)Rn * (2 + C), &d->wire_shape_scratch);  // GT pairs + column lengths, expanded
    }
    size_t total = 0;
    size_t copied = 0;
    for (auto& it : items) {
        total += ((it.bytes + 255) & ~(size_t)255) + 256;
        if (it.src) copied += it.bytes;
    }
    d->in_bytes = copied;  // host-to-device bytes of this upload
    d->in_alloc = total ? total : 256;
    cudaError_t e = ctx->take((void**)&d->in_base, &d->in_alloc);
    if (e != cudaSuccess) return bail(e, "cudaMalloc(input arena)");
    Arena ar{d->in_base, total, 0};
    for (auto& it : items) {
        char* p = ar.take<char>(it.bytes ? it.bytes : 1);
        *it.dst = p;
        if (it.bytes && it.src) {
            e = cudaMemcpyAsync(p, it.src, it.bytes, cudaMemcpyHostToDevice, up);
            if (e != cudaSuccess) return bail(e, "cudaMemcpyAsync(H2D)");
        }
    }
    if (wire) {  // expand the wire form into the arrays the kernels read
        DWire W;
        memset(&W, 0, sizeof W);
        W.n_records = Rn;
        W.n_variant = wire->n_variant;
        W.n_al = wire->n_al;
        W.n_al_pool = wire->n_al_pool;
        W.n_pool = wire->n_pool;
        W.n_exc = wire->n_maxend_exc;
        W.n_cols = wire->n_cols;
        W.flags = wire->flags;
        W.base = reinterpret_cast<const uint8_t*>(wire_dev);
        uint64_t co = wire->off[GLX_WS_COL_VAL];
        for (uint32_t c = 0; c < wire->n_cols; c++) {
            W.col_w[c] = wire->col_w[c];
            W.col_val_off[c] = co;
            co += ((Rn * wire->col_w[c]) + 15) & ~(uint64_t)15;
        }
        for (int i = 0; i < GLX_WS_COUNT; i++) W.off[i] = wire->off[i];
        // columns that travel as they are
        R.ds_rec_off = reinterpret_cast<const uint64_t*>(wire_dev + wire->off[GLX_WS_DS_REC_OFF]);
        R.beg = reinterpret_cast<const int32_t*>(wire_dev + wire->off[GLX_WS_BEG]);
        if (wire->flags & GLX_WF_SHAPES) {
            // (flags, n_allele, GT pair, column lengths) come as one shape code per record: expand them first — flags and n_allele into
            // the arrays the arena holds for them anyway, the GT pairs and column lengths into scratch the record pass reads from
            if (!(wire->flags & GLX_WF_GT8) || !(wire->flags & GLX_WF_COLLEN8) || wire->n_shapes > 255) return bail(cudaErrorInvalidValue, "glx_upload_wire: malformed shape sections");
            uint8_t* scratch = reinterpret_cast<uint8_t*>(d->wire_shape_scratch);
            W.x_flags = R.flags;
            W.x_nal = R.n_allele;
            W.x_gt8 = scratch;
            W.x_collen8 = scratch + 2 * Rn;
            W.n_shapes = wire->n_shapes;
            W.n_shape_exc = wire->n_shape_exc;
            if (Rn)
                glx_wire_shapes_kernel<<<(unsigned)std::min<uint64_t>((Rn + 255) / 256, (uint64_t)ctx->sm_count * 8), 256, 0, up>>>(
                    W, const_cast<uint8_t*>(R.flags), const_cast<uint8_t*>(R.n_allele), scratch, scratch + 2 * Rn);
        } else {
            R.flags = reinterpret_cast<const uint8_t*>(wire_dev + wire->off[GLX_WS_FLAGS]);
            R.n_allele = reinterpret_cast<const uint8_t*>(wire_dev + wire->off[GLX_WS_NALLELE]);
        }
        R.al_pool = wire_dev + wire->off[GLX_WS_AL_POOL];
        const unsigned nb = (unsigned)((Rn + GLX_WIRE_BLOCK * GLX_WIRE_ITEMS - 1) / (GLX_WIRE_BLOCK * GLX_WIRE_ITEMS));
        const unsigned nab = (unsigned)((recs->n_al + GLX_WIRE_BLOCK * GLX_WIRE_ITEMS - 1) / (GLX_WIRE_BLOCK * GLX_WIRE_ITEMS));
        uint3x* sums = reinterpret_cast<uint3x*>(d->wire_sums);
        uint3x* al_sums = reinterpret_cast<uint3x*>(d->wire_al_sums);
        if (nab) {
            glx_wire_al_sums_kernel<<<nab, GLX_WIRE_BLOCK, 0, up>>>(W, al_sums);
            glx_wire_scan_sums_kernel<<<1, GLX_WIRE_BLOCK, 0, up>>>(al_sums, nab);
            glx_wire_alleles_kernel<<<nab, GLX_WIRE_BLOCK, 0, up>>>(W, al_sums, const_cast<uint32_t*>(R.al_len), const_cast<uint32_t*>(R.al_str),
                                                                      const_cast<uint64_t*>(R.al_prefix));
        }
        if (nb) {
            glx_wire_sums_kernel<<<nb, GLX_WIRE_BLOCK, 0, up>>>(W, sums);
            glx_wire_scan_sums_kernel<<<1, GLX_WIRE_BLOCK, 0, up>>>(sums, nb);
            WireOut WO;
            WO.end = const_cast<int32_t*>(R.end);
            WO.maxend = const_cast<int32_t*>(R.maxend);
            WO.col_val = const_cast<int32_t*>(R.col_val);
            WO.pool = const_cast<int32_t*>(R.pool);
            WO.qual = const_cast<float*>(R.qual);
            WO.filt = const_cast<uint32_t*>(R.filt);
            WO.gt = const_cast<uint32_t*>(R.gt);
            WO.allele_off = const_cast<uint32_t*>(R.allele_off);
            WO.col_len = const_cast<uint16_t*>(R.col_len);
            WO.al_flags = const_cast<uint8_t*>(R.al_flags);
            WO.al_len = R.al_len;
            WO.al_str = R.al_str;
            glx_wire_records_kernel<<<nb, GLX_WIRE_BLOCK, 0, up>>>(W, sums, WO);
            if (W.n_exc) glx_wire_maxend_kernel<<<(unsigned)((W.n_exc + 255) / 256), 256, 0, up>>>(W, WO.maxend);
        }
        if (W.n_pool) glx_wire_pool_kernel<<<(unsigned)std::min<uint64_t>((W.n_pool + 255) / 256, (uint64_t)ctx->sm_count * 16), 256, 0, up>>>(W, WO_pool_ptr(R));
        e = cudaGetLastError();
        if (e != cudaSuccess) return bail(e, "wire expansion kernels");
    }
    {
        const uint64_t n = std::max<uint64_t>(std::max<uint64_t>(S + 1, nU + 1), D);
        if (ctx->opt.profile_upload) {
            for (int i = 0; i < 2; i++) cudaEventCreate(&d->up_ev[i]);
            cudaEventRecord(d->up_ev[0], up);
        }
        glx_prepare_inputs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, up>>>(R, Sx, (uint32_t)D, nU, const_cast<uint8_t*>(R.flags), const_cast<int4*>(Sx.site_rows),
                                                                               const_cast<int4*>(Sx.unif_rows));
        e = cudaGetLastError();
        if (ctx->opt.profile_upload) cudaEventRecord(d->up_ev[1], up);
        if (e != cudaSuccess) return bail(e, "glx_prepare_inputs_kernel");
    }
    // max output slots of any cell (scratch sizing of the general path)
    uint32_t max_slots = 1;
    for (uint64_t s = 0; s < S; s++) {
        uint32_t t = 0;
        for (uint32_t k = 0; k < cfg->n_fields; k++) t += sites->fld_cnt[k][s];
        if (t > max_slots) max_slots = t;
    }
    d->max_slots = max_slots;
    for (uint64_t s = 0; s < S; s++) d->max_site_alleles = std::max<uint32_t>(d->max_site_alleles, sites->n_allele[s]);

    // ---- output arena ----
    size_t ob = 0;
    auto osz = [&](size_t b) { ob += ((b + 255) & ~(size_t)255) + 256; };
    osz(S * N * 2);
    osz(S * N * 2);
    osz(S * 4);
    osz(S + 4);
    osz(8);
    for (uint32_t k = 0; k < cfg->n_fields; k++) osz(d->fld_total[k] * 4);
    d->out_bytes = ob;
    d->out_alloc = ob;
    e = ctx->take((void**)&d->out_base, &d->out_alloc);
    if (e != cudaSuccess) return bail(e, "cudaMalloc(output arena)");
    Arena oa{d->out_base, ob, 0};
    memset(&d->O, 0, sizeof d->O);
    d->O.gt = oa.take<int8_t>(S * N * 2);
    d->O.rnc = oa.take<uint8_t>(S * N * 2);
    d->O.allele_used = oa.take<uint32_t>(S);
    d->O.site_flags = oa.take<uint8_t>(S + 4);
    d->O.error = oa.take<unsigned long long>(1);
    for (uint32_t k = 0; k < cfg->n_fields; k++) d->O.fld[k] = oa.take<int32_t>(d->fld_total[k]);

    // ---- general-path scratch ----
    const int threads = 128;
    uint64_t want_blocks = (d->n_cells + threads - 1) / threads;
#ifndef GLX_GENERAL_BLOCKS_PER_SM
#define GLX_GENERAL_BLOCKS_PER_SM 96
#endif
    // (the worklist pass walks regions block by block: with as many blocks as regions every region's few cells start at once instead
    // of queueing behind the block's other regions — the kernel is a chain of dependent loads per cell; the scratch bound keeps it small)
    uint64_t cap = (uint64_t)ctx->sm_count * GLX_GENERAL_BLOCKS_PER_SM;
    while (cap > (uint64_t)ctx->sm_count * 16 && cap * threads * max_slots > (128ull << 20)) cap /= 2;
    d->general_blocks = (int)(want_blocks < cap ? (want_blocks ? want_blocks : 1) : cap);
    d->cnt_scratch_bytes = (size_t)d->general_blocks * threads * max_slots;
    e = ctx->take((void**)&d->cnt_scratch, &d->cnt_scratch_bytes);
    if (e != cudaSuccess) return bail(e, "cudaMalloc(scratch)");
    d->plan = make_fast_plan(d->cfg);
    if (d->force_general) d->plan.enabled = 0;
    {
        uint32_t nf;
        uint64_t modes, nums, nrvs;
        d->tiled_kernel = glx_cells_tiled_kernel;
        d->tiled_smem = (size_t)d->plan.n_rv * GLX_TILE_THREADS * 4;
        if (!ctx->opt.force_dynamic_tiled && sig_of_plan(d->cfg, d->plan, nf, modes, nums, nrvs)) {
            const bool no_rsig = ctx->opt.no_resolve_sig != 0;
            // Fusing pays when the output the resident CTAs have in flight (~1,000 half-height tiles) is of the order of
            // L2: up to ~48 output bytes per cell (DeepVariant-style fields at mostly biallelic sites are 32; gatk-style
            // cohorts with up to 7 alleles are ~100 and measured 6 % slower fused). Cells of sites with more than 4
            // alleles are declined by the fused form and go to glx_cells_variant_kernel<8>.
            uint64_t out_bytes = d->n_cells * 4;
            for (uint32_t k = 0; k < cfg->n_fields; k++) out_bytes += d->fld_total[k] * 4;
            const bool fuse = d->plan.vview && out_bytes <= d->n_cells * 48 && !ctx->opt.no_fuse;
            // A CTA finishes its lone-variant-record cells when its tile is done; their output sectors are still in L2
            // only if the tiles in flight are small enough, so variant-dense batches (>= 1 variant record per 100
            // cells) walk half-height tiles. Sparse ones keep the full height (fewer cursor set-ups per site).
            const bool short_tiles = fuse && n_variant_records * 100 >= d->n_cells;
            auto match = [&](uint32_t NF, uint64_t M, uint64_t Nu, uint64_t Nr) { return nf == NF && modes == M && nums == Nu && nrvs == Nr; };
            if (match(4, 0x1010ull, 0x2030ull, 0x4141ull)) {
                d->tiled_kernel = glx_cells_tiled_sig_kernel<SigDeepVariant, 0, GLX_TILE_SITES>;
                if (d->plan.vview && !no_rsig) d->resolve_sig = glx_resolve_sig_kernel<SigDeepVariant>;
                if (fuse) {
                    d->tiled_kernel = short_tiles ? glx_cells_tiled_sig_kernel<SigDeepVariant, 4, GLX_TILE_SITES / 2> : glx_cells_tiled_sig_kernel<SigDeepVariant, 4, GLX_TILE_SITES>;
                    d->fused_variant = true;
                }
                d->tiled_smem = 0;
            } else if (match(5, 0x20010ull, 0x20030ull, 0x01441ull)) {
                d->tiled_kernel = glx_cells_tiled_sig_kernel<SigGatk, 0, GLX_TILE_SITES>;
                if (d->plan.vview && !no_rsig) d->resolve_sig = glx_resolve_sig_kernel<SigGatk>;
                if (fuse) {
                    d->tiled_kernel = short_tiles ? glx_cells_tiled_sig_kernel<SigGatk, 4, GLX_TILE_SITES / 2> : glx_cells_tiled_sig_kernel<SigGatk, 4, GLX_TILE_SITES>;
                    d->fused_variant = true;
                }
                d->tiled_smem = 0;
            }
        }
    }
    d->n_sample_blocks = (uint32_t)((N + d->samples_per_cta - 1) / d->samples_per_cta);
    if (d->fused_variant && n_variant_records * 100 >= d->n_cells) d->tile_sites = GLX_TILE_SITES / 2;
    d->n_site_tiles = (uint32_t)((S + d->tile_sites - 1) / d->tile_sites);
    if (d->plan.enabled && d->n_cells) {
        const size_t n_regions = (size_t)d->n_sample_blocks * d->n_site_tiles;
        const size_t region_cells = (size_t)d->tile_sites * d->samples_per_cta;
        // declined list of the fused tiled kernel: per chunk a counter and a share of room for 1/8 of the batch's cells
        d->decl_cap = d->fused_variant ? (uint32_t)std::min<uint64_t>(d->n_cells / 8 + 4096, 1u << 28) : 0u;
        const size_t decl_bytes = d->fused_variant ? (size_t)d->decl_cap * 8 + 2 * GLX_MAX_CHUNKS * 4 + 64 : 0;
        const size_t nwork_bytes = ((n_regions + 64) * 4 + 7) & ~(size_t)7;
        d->worklist_bytes = n_regions * region_cells * 8 + nwork_bytes + decl_bytes;
        e = ctx->take((void**)&d->worklist, &d->worklist_bytes);
        if (e != cudaSuccess) return bail(e, "cudaMalloc(worklist)");
        d->n_work = reinterpret_cast<uint32_t*>(d->worklist + n_regions * region_cells);
        d->decl_list = d->fused_variant ? reinterpret_cast<uint2*>(reinterpret_cast<char*>(d->n_work) + nwork_bytes) : nullptr;
        d->decl_n = d->fused_variant ? reinterpret_cast<uint32_t*>(d->decl_list + d->decl_cap) : nullptr;
        const size_t blob_part = ((Rn + 1) * blob_words(d->plan.n_rv) * 4 + 255) & ~(size_t)255;
        const size_t cur_part = (((size_t)d->n_site_tiles * N + 1) * 4 + 255) & ~(size_t)255;
        const size_t tq_part = (((size_t)d->n_site_tiles + 1) * 4 + 255) & ~(size_t)255;
        const size_t n_rblocks = (Rn + GLX_RESOLVE_THREADS - 1) / GLX_RESOLVE_THREADS;
        d->blob_bytes = blob_part + cur_part + tq_part + (n_rblocks + 2) * sizeof(uint2);
        e = ctx->take((void**)&d->blobs, &d->blob_bytes);
        if (e != cudaSuccess) return bail(e, "cudaMalloc(record blobs)");
        d->tile_cursors = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(d->blobs) + blob_part);
        d->tile_q = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(d->blobs) + blob_part + cur_part);
        // tile_q[t] = the smallest qbeg of any site from tile t on (non-decreasing, never later than a later tile's start), on the
        // device: a host-built array would be a pageable copy in the middle of the batch's stream (the host would wait for the
        // stream, and the copy would queue behind other batches' uploads)
        glx_tile_q_kernel<<<1, 1024, 0, up>>>(d->S.qbeg, d->n_site_tiles, d->tile_sites, d->tile_q);
        e = cudaGetLastError();
        if (e != cudaSuccess) return bail(e, "glx_tile_q_kernel");
        d->block_hints = reinterpret_cast<uint2*>(reinterpret_cast<char*>(d->blobs) + blob_part + cur_part + tq_part);
        if (ctx->opt.profile_upload) {
            for (int i = 2; i < 4; i++) cudaEventCreate(&d->up_ev[i]);
            cudaEventRecord(d->up_ev[2], up);
        }
        glx_block_hints_kernel<<<(unsigned)((n_rblocks + 1 + 255) / 256), 256, 0, up>>>(d->R, (uint32_t)D, d->tile_q, d->n_site_tiles, d->block_hints, (uint32_t)n_rblocks);
        e = cudaGetLastError();
        if (ctx->opt.profile_upload) cudaEventRecord(d->up_ev[3], up);
        if (e != cudaSuccess) return bail(e, "glx_block_hints_kernel");
    }
    d->n_datasets = (uint32_t)D;
    e = sync ? cudaStreamSynchronize(up) : cudaSuccess;  // after a synchronous upload the caller's host arrays may be released
    if (e != cudaSuccess) return bail(e, "cudaStreamSynchronize(upload)");
    *out = d;
    return GLX_OK;
}

extern "C" {

int glx_upload_on(glx_ctx* ctx, const glx_config* cfg, const glx_records* recs, const glx_sites* sites, void* stream, int sync, glx_dev_batch** dev) {
    if (!recs) {
        if (ctx) ctx->err = "glx_upload: null argument";
        return GLX_E_INVALID;
    }
    return upload_impl(ctx, cfg, recs, nullptr, sites, stream, sync, dev);
}

int glx_upload_wire_on(glx_ctx* ctx, const glx_config* cfg, const glx_wire* wire, const glx_sites* sites, void* stream, int sync, glx_dev_batch** dev) {
    if (!wire) {
        if (ctx) ctx->err = "glx_upload_wire: null argument";
        return GLX_E_INVALID;
    }
    return upload_impl(ctx, cfg, nullptr, wire, sites, stream, sync, dev);
}

// test aid: which record-store arrays on the device differ from the host's
int glx_debug_compare_records(glx_ctx* ctx, glx_dev_batch* d, const glx_records* recs, uint32_t* diff) {
    if (!ctx || !d || !recs || !diff) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    *diff = 0;
    const uint64_t Rn = recs->n_records, C = d->n_cols;
    if (Rn != d->R.n_records) {
        *diff = ~0u;
        return GLX_OK;
    }
    std::vector<char> tmp;
    auto fetch = [&](const void* dev, size_t bytes) -> const char* {
        tmp.assign(bytes + 8, 0);
        if (bytes && cudaMemcpy(tmp.data(), dev, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return nullptr;
        return tmp.data();
    };
    auto whole = [&](int bit, const void* dev, const void* host, size_t bytes) {
        const char* g = fetch(dev, bytes);
        if (!g || (bytes && memcmp(g, host, bytes) != 0)) *diff |= 1u << bit;
    };
    const bool var_only = true;
    whole(0, d->R.beg, recs->beg, Rn * 4);
    whole(1, d->R.end, recs->end, Rn * 4);
    whole(2, d->R.maxend, recs->maxend, Rn * 4);
    whole(4, d->R.n_allele, recs->n_allele, Rn);
    whole(7, d->R.gt, recs->gt, Rn * 4);
    whole(10, d->R.col_len, recs->col_len, C * Rn * 2);
    whole(11, d->R.pool, recs->pool, recs->n_pool * 4);
    whole(12, d->R.al_len, recs->al_len, recs->n_al * 4);
    whole(13, d->R.al_flags, recs->al_flags, recs->n_al);
    whole(14, d->R.al_prefix, recs->al_prefix, recs->n_al * 8);
    whole(15, d->R.al_str, recs->al_str, recs->n_al * 4);
    whole(16, d->R.al_pool, recs->al_pool, recs->n_al_pool);
    {  // qual (3), filt (6), allele_off (8): compared on non-reference records; flags (5) without the device-private bit
        const char* g = fetch(d->R.qual, Rn * 4);
        for (uint64_t r = 0; g && r < Rn; r++)
            if (!(recs->flags[r] & GLX_RF_IS_REF) && memcmp(g + 4 * r, &recs->qual[r], 4) != 0) *diff |= 1u << 3;
        if (!g) *diff |= 1u << 3;
        g = fetch(d->R.filt, Rn * 4);
        for (uint64_t r = 0; g && r < Rn; r++)
            if ((!(recs->flags[r] & GLX_RF_IS_REF) || !var_only) && memcmp(g + 4 * r, &recs->filt[r], 4) != 0) *diff |= 1u << 6;
        if (!g) *diff |= 1u << 6;
        g = fetch(d->R.allele_off, Rn * 4);
        for (uint64_t r = 0; g && r < Rn; r++)
            if (!(recs->flags[r] & GLX_RF_IS_REF) && memcmp(g + 4 * r, &recs->allele_off[r], 4) != 0) *diff |= 1u << 8;
        if (!g) *diff |= 1u << 8;
        g = fetch(d->R.flags, Rn);
        for (uint64_t r = 0; g && r < Rn; r++)
            if (((uint8_t)g[r] & 0x7f) != (recs->flags[r] & 0x7f)) *diff |= 1u << 5;
        if (!g) *diff |= 1u << 5;
        g = fetch(d->R.col_val, C * Rn * 4);
        for (uint64_t i = 0; g && i < C * Rn; i++)
            if (recs->col_len[i] < 0xFFF0 && memcmp(g + 4 * i, &recs->col_val[i], 4) != 0) *diff |= 1u << 9;
        if (!g) *diff |= 1u << 9;
    }
    return GLX_OK;
}

int glx_launch(glx_ctx* ctx, glx_dev_batch* d, void* stream_) {
    if (!ctx || !d) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    d->last_stream = st;
    const uint64_t S = d->n_sites;
    CU(cudaMemsetAsync(d->O.allele_used, 0, S * 4, st));
    CU(cudaMemsetAsync(d->O.site_flags, 0, S + 4, st));
    CU(cudaMemsetAsync(d->O.error, 0xff, 8, st));
    d->launches_per_call = 0;
    const bool tiled = d->plan.enabled && d->n_cells;
    if (d->profile) CU(cudaEventRecord(d->ev[0], st));
    if (tiled) {
        CU(cudaMemsetAsync(d->tile_cursors, 0xff, (size_t)d->n_site_tiles * d->n_samples * 4, st));
        if (d->R.n_records) {
            const unsigned rgrid = (unsigned)((d->R.n_records + GLX_RESOLVE_THREADS - 1) / GLX_RESOLVE_THREADS);
            if (d->resolve_sig)
                d->resolve_sig<<<rgrid, GLX_RESOLVE_THREADS, 0, st>>>(d->cfg, d->plan, d->R, d->n_datasets, d->blobs, d->tile_q, d->n_site_tiles, d->tile_cursors,
                                                                      d->block_hints);
            else
                glx_resolve_records_kernel<<<rgrid, GLX_RESOLVE_THREADS, (size_t)GLX_RESOLVE_THREADS * (blob_words(d->plan.n_rv) + 1) * 4, st>>>(
                    d->cfg, d->plan, d->R, d->n_datasets, d->blobs, d->tile_q, d->n_site_tiles, d->tile_cursors, d->block_hints);
            d->launches_per_call++;
            CU(cudaGetLastError());
        }
        if (d->profile) CU(cudaEventRecord(d->ev[1], st));
        // The tile grid is launched in chunks: the worklist kernels of chunk j (latency-bound gathers) run on the aux
        // stream while the tiled kernel of chunk j+1 (store-heavy) runs on the launching stream.
        const uint32_t n_regions = d->n_sample_blocks * d->n_site_tiles;
        uint32_t n_chunks = st == ctx->stream ? d->n_chunks : 1u;  // the aux stream and its events belong to the context stream
        if (n_chunks > n_regions) n_chunks = n_regions;
        if (d->decl_list) CU(cudaMemsetAsync(d->decl_n, 0, 2 * GLX_MAX_CHUNKS * 4, st));  // [j]: entries offered by chunk j; [GLX_MAX_CHUNKS + j]: finished by the list kernel
        for (uint32_t j = 0; j < n_chunks; j++) {
            const uint32_t c0 = (uint32_t)((uint64_t)n_regions * j / n_chunks), c1 = (uint32_t)((uint64_t)n_regions * (j + 1) / n_chunks);
            if (c1 == c0) continue;
            FastPlan plan_j = d->plan;
            if (d->decl_list) {  // this chunk's share of the declined list
                const uint32_t share = d->decl_cap / n_chunks;
                plan_j.decl_list = d->decl_list + (size_t)j * share;
                plan_j.decl_n = d->decl_n + j;
                plan_j.decl_cap = share;
            }
            d->tiled_kernel<<<c1 - c0, GLX_TILE_THREADS, d->tiled_smem, st>>>(d->cfg, plan_j, d->R, d->S, d->O, d->blobs, d->tile_cursors, d->worklist, d->n_work,
                                                                            d->n_sample_blocks, c0);
            d->launches_per_call++;
            CU(cudaGetLastError());
            cudaStream_t ws = n_chunks > 1 ? ctx->aux : st;
            if (n_chunks > 1) {
                CU(cudaEventRecord(ctx->chunk_ev[j], st));
                CU(cudaStreamWaitEvent(ws, ctx->chunk_ev[j], 0));
            }
            const uint32_t nreg = c1 - c0;
#ifndef GLX_WORKLIST_BLOCKS_PER_SM
#define GLX_WORKLIST_BLOCKS_PER_SM 96
#endif
            const uint32_t g1 = nreg < (uint32_t)ctx->sm_count * GLX_WORKLIST_BLOCKS_PER_SM ? nreg : (uint32_t)ctx->sm_count * GLX_WORKLIST_BLOCKS_PER_SM;
            if (d->profile && n_chunks == 1) CU(cudaEventRecord(d->ev[2], st));
            const uint32_t region_cells = d->tile_sites * d->samples_per_cta;
            if (d->decl_list) {
                glx_cells_variant_list_kernel<4><<<ctx->sm_count * GLX_VARIANT_MIN_BLOCKS, 128, VariantSmem<4>::SLOTS * 128 * 4, ws>>>(d->cfg, plan_j, d->R, d->S, d->O, d->worklist,
                                                                                                                                 d->n_work, region_cells, d->blobs);
                d->launches_per_call++;
            }
            glx_cells_bands_kernel<<<g1, 128, 0, ws>>>(d->cfg, d->plan, d->R, d->S, d->O, d->worklist, d->n_work, c0, c1, region_cells, d->blobs);
            d->launches_per_call++;
            if (d->plan.vview && (!d->fused_variant || d->max_site_alleles > 4)) {  // small sites keep the per-thread shared-memory columns short
                if (d->max_site_alleles <= 4)
                    glx_cells_variant_kernel<4><<<g1, 128, VariantSmem<4>::SLOTS * 128 * 4, ws>>>(d->cfg, d->plan, d->R, d->S, d->O, d->worklist, d->n_work, c0, c1,
                                                                                                   region_cells, d->blobs);
                else
                    glx_cells_variant_kernel<8><<<g1, 128, VariantSmem<8>::SLOTS * 128 * 4, ws>>>(d->cfg, d->plan, d->R, d->S, d->O, d->worklist, d->n_work, c0, c1,
                                                                                                   region_cells, d->blobs);
                d->launches_per_call++;
            }
            glx_cells_single_kernel<<<g1, 128, 0, ws>>>(d->cfg, d->R, d->S, d->O, d->worklist, d->n_work, c0, c1, region_cells);
            if (d->profile && n_chunks == 1) CU(cudaEventRecord(d->ev[3], st));
            const uint32_t g2 = nreg < (uint32_t)d->general_blocks ? nreg : (uint32_t)d->general_blocks;
            glx_cells_general_kernel<<<g2, 128, 0, ws>>>(d->cfg, d->R, d->S, d->O, d->worklist, d->n_work, c0, c1, region_cells, d->n_cells, d->cnt_scratch,
                                                         d->max_slots, d->n_cols);
            d->launches_per_call += 2;
            CU(cudaGetLastError());
        }
        if (d->profile && n_chunks > 1) {
            CU(cudaEventRecord(d->ev[2], st));
            CU(cudaEventRecord(d->ev[3], st));
        }
        if (n_chunks > 1) {  // join
            CU(cudaEventRecord(ctx->chunk_ev[GLX_MAX_CHUNKS], ctx->aux));
            CU(cudaStreamWaitEvent(st, ctx->chunk_ev[GLX_MAX_CHUNKS], 0));
        }
    } else {
        if (d->profile) {
            CU(cudaEventRecord(d->ev[1], st));
            CU(cudaEventRecord(d->ev[2], st));
            CU(cudaEventRecord(d->ev[3], st));
        }
        if (d->n_cells && d->multi) {
            glx_cells_multi_kernel<<<d->general_blocks, 128, 0, st>>>(d->cfg, d->R, d->gt_pool, d->out_map, d->S, d->O, d->n_cells, d->cnt_scratch);
            d->launches_per_call++;
            CU(cudaGetLastError());
        } else if (d->n_cells) {
            glx_cells_general_kernel<<<d->general_blocks, 128, 0, st>>>(d->cfg, d->R, d->S, d->O, nullptr, d->n_work, 0, 0, 0, d->n_cells, d->cnt_scratch, d->max_slots,
                                                                        d->n_cols);
            d->launches_per_call++;
            CU(cudaGetLastError());
        }
    }
    if (d->profile) CU(cudaEventRecord(d->ev[4], st));
    return GLX_OK;
}

int glx_check(glx_ctx* ctx, glx_dev_batch* d) {
    if (!ctx || !d) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    unsigned long long w = 0;
    cudaStream_t st = d->last_stream ? d->last_stream : ctx->stream;  // the stream the batch's kernels were launched on
    CU(cudaMemcpyAsync(&w, d->O.error, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (w == ~0ull) return GLX_OK;
    const int code = (int)(w & 0xff);
    const unsigned long long site = (w >> 8) / (d->n_samples ? d->n_samples : 1);
    char buf[160];
    snprintf(buf, sizeof buf, "genotyper: site %llu failed with status %d (first failing site of the batch)", site, -code);
    ctx->err = buf;
    return code == 101 ? GLX_E_UNSUPPORTED : -code;
}

int glx_download(glx_ctx* ctx, glx_dev_batch* d, glx_out* out) { return glx_download_on(ctx, d, out, nullptr, 1); }

int glx_download_on(glx_ctx* ctx, glx_dev_batch* d, glx_out* out, void* stream_, int sync) {
    if (!ctx || !d || !out) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    const uint64_t S = d->n_sites, N = d->n_samples;
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    if (glx_take_err_slot(ctx, d) != GLX_OK) return cuda_fail(ctx, cudaErrorMemoryAllocation, "pinned status word");
    d->dl_stream = st;
    if (S * N) {
        CU(cudaMemcpyAsync(out->gt, d->O.gt, S * N * 2, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->rnc, d->O.rnc, S * N * 2, cudaMemcpyDeviceToHost, st));
    }
    if (S) {
        CU(cudaMemcpyAsync(out->allele_used, d->O.allele_used, S * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->site_flags, d->O.site_flags, S, cudaMemcpyDeviceToHost, st));
    }
    for (uint32_t k = 0; k < d->n_fields; k++)
        if (d->fld_total[k]) CU(cudaMemcpyAsync(out->fld[k], d->O.fld[k], d->fld_total[k] * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(d->host_err, d->O.error, 8, cudaMemcpyDeviceToHost, st));
    if (sync) CU(cudaStreamSynchronize(st));
    return GLX_OK;
}

static bool field_is_narrowable(const glx_dev_batch* d, uint32_t k) {
    const DFld& f = d->cfg.f[k];
    return f.type == GLX_TYPE_INT && f.kind != GLX_FK_FT && f.kind != GLX_FK_STRING;
}

int glx_download_narrow_on(glx_ctx* ctx, glx_dev_batch* d, glx_out16* out, void* stream_, int sync) {
    if (!ctx || !d || !out) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    const uint64_t S = d->n_sites, N = d->n_samples;
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    if (glx_take_err_slot(ctx, d) != GLX_OK) return cuda_fail(ctx, cudaErrorMemoryAllocation, "pinned status word");
    d->dl_stream = st;
    if (!d->narrow) {
        size_t b = 256;
        for (uint32_t k = 0; k < d->n_fields; k++)
            if (field_is_narrowable(d, k)) b += ((d->fld_total[k] * 2 + 255) & ~(size_t)255);
        d->narrow_bytes = b;
        cudaError_t e = ctx->take((void**)&d->narrow, &d->narrow_bytes);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(narrow columns)");
    }
    d->narrow_overflow = reinterpret_cast<uint32_t*>(d->narrow);
    CU(cudaMemsetAsync(d->narrow_overflow, 0, GLX_MAX_FIELDS * 4, st));
    size_t off = 256;
    for (uint32_t k = 0; k < d->n_fields; k++) {
        if (!d->fld_total[k]) continue;
        if (!field_is_narrowable(d, k)) {
            CU(cudaMemcpyAsync(out->fld32[k], d->O.fld[k], d->fld_total[k] * 4, cudaMemcpyDeviceToHost, st));
            continue;
        }
        char* dst = reinterpret_cast<char*>(d->narrow) + off;
        off += (d->fld_total[k] * 2 + 255) & ~(size_t)255;
        const uint64_t n = d->fld_total[k];
        const unsigned grid = (unsigned)std::min<uint64_t>((n / 4 + 255) / 256 + 1, (uint64_t)ctx->sm_count * 16);
        const int w = (ctx->width_hint[k] == 1 && out->fld8[k]) ? 1 : 2;
        d->narrow_width[k] = (uint8_t)w;
        if (w == 1) glx_narrow_kernel<int8_t><<<grid, 256, 0, st>>>(d->O.fld[k], reinterpret_cast<int8_t*>(dst), n, d->narrow_overflow + k);
        else glx_narrow_kernel<int16_t><<<grid, 256, 0, st>>>(d->O.fld[k], reinterpret_cast<int16_t*>(dst), n, d->narrow_overflow + k);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(w == 1 ? (void*)out->fld8[k] : (void*)out->fld16[k], dst, n * w, cudaMemcpyDeviceToHost, st));
    }
    if (S * N) {
        CU(cudaMemcpyAsync(out->gt, d->O.gt, S * N * 2, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->rnc, d->O.rnc, S * N * 2, cudaMemcpyDeviceToHost, st));
    }
    if (S) {
        CU(cudaMemcpyAsync(out->allele_used, d->O.allele_used, S * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(out->site_flags, d->O.site_flags, S, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaMemcpyAsync(d->host_overflow, d->narrow_overflow, GLX_MAX_FIELDS * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(d->host_err, d->O.error, 8, cudaMemcpyDeviceToHost, st));
    if (sync) CU(cudaStreamSynchronize(st));
    return GLX_OK;
}

int glx_narrow_finish(glx_ctx* ctx, glx_dev_batch* d, glx_out16* out) {
    if (!ctx || !d || !out || !d->host_overflow) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    bool refetched = false;
    for (uint32_t k = 0; k < d->n_fields; k++) {
        if (!field_is_narrowable(d, k)) {
            out->width[k] = 4;
            continue;
        }
        const uint32_t fl = d->host_overflow[k];
        const int w = d->narrow_width[k] ? d->narrow_width[k] : 2;
        if (fl & (w == 1 ? 2u : 1u)) {  // some value does not fit the width sent: fetch the int32 column (rare) on the batch's own stream
            out->width[k] = 4;
            if (d->fld_total[k]) {
                CU(cudaMemcpyAsync(out->fld32[k], d->O.fld[k], d->fld_total[k] * 4, cudaMemcpyDeviceToHost, d->dl_stream ? d->dl_stream : ctx->stream));
                refetched = true;
            }
        } else
            out->width[k] = (uint8_t)w;
        if (d->fld_total[k]) ctx->width_hint[k] = (fl & 2u) ? 2 : 1;  // what the next batch of this range stream is sent in
    }
    if (refetched) CU(cudaStreamSynchronize(d->dl_stream ? d->dl_stream : ctx->stream));  // other streams keep running
    return GLX_OK;
}

// status of a batch whose (asynchronous) download has completed
int glx_status(glx_ctx* ctx, glx_dev_batch* d) {
    if (!ctx || !d || !d->host_err) return GLX_E_INVALID;
    const unsigned long long w = *d->host_err;
    if (w == ~0ull) return GLX_OK;
    const int code = (int)(w & 0xff);
    char buf[160];
    snprintf(buf, sizeof buf, "genotyper: site %llu failed with status %d (first failing site of the batch)", (w >> 8) / (d->n_samples ? d->n_samples : 1), -code);
    ctx->err = buf;
    return code == 101 ? GLX_E_UNSUPPORTED : -code;
}

void* glx_stream_create(glx_ctx* ctx) {
    if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    return (void*)s;
}
void glx_stream_destroy(glx_ctx* ctx, void* s) {
    if (ctx && s) {
        cudaSetDevice(ctx->device);
        cudaStreamDestroy((cudaStream_t)s);
    }
}
int glx_stream_sync(glx_ctx* ctx, void* s) {
    if (!ctx) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(s ? (cudaStream_t)s : ctx->stream));
    return GLX_OK;
}
void* glx_mark(glx_ctx* ctx, void* s) {
    if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    if (cudaEventRecord(e, s ? (cudaStream_t)s : ctx->stream) != cudaSuccess) {
        cudaEventDestroy(e);
        return nullptr;
    }
    return (void*)e;
}
float glx_mark_ms(glx_ctx* ctx, void* a, void* b) {
    float ms = -1.f;
    if (!ctx || !a || !b || cudaSetDevice(ctx->device) != cudaSuccess) return ms;
    if (cudaEventSynchronize((cudaEvent_t)b) != cudaSuccess || cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) return -1.f;
    return ms;
}
int glx_mark_wait(glx_ctx* ctx, void* m) {
    if (!ctx || !m) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize((cudaEvent_t)m));
    return GLX_OK;
}
void glx_mark_free(glx_ctx* ctx, void* m) {
    if (ctx && m) {
        cudaSetDevice(ctx->device);
        cudaEventDestroy((cudaEvent_t)m);
    }
}
int glx_host_register(void* p, size_t bytes) { return (!p || !bytes || cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess) ? GLX_OK : GLX_E_CUDA; }
void glx_host_unregister(void* p) {
    if (p) cudaHostUnregister(p);
}

int glx_selftest_diploid(glx_ctx* ctx, uint32_t n_allele, uint32_t* out) {
    if (!ctx || !out || n_allele == 0 || n_allele > GLX_MAX_ALLELES) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)n_allele * (n_allele + 1) / 2;
    uint32_t* dv = nullptr;
    CU(cudaMalloc(&dv, n * 4));
    glx_selftest_diploid_kernel<<<n_allele, GLX_MAX_ALLELES, 0, ctx->stream>>>(n_allele, dv);
    cudaError_t e = cudaMemcpyAsync(out, dv, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dv);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "glx_selftest_diploid");
    return GLX_OK;
}

int glx_set_option(glx_ctx* ctx, const char* name, int value) {
    if (!ctx || !name) return GLX_E_INVALID;
    const std::string n(name);
    if (n == "force_general") ctx->opt.force_general = value;
    else if (n == "force_dynamic_tiled") ctx->opt.force_dynamic_tiled = value;
    else if (n == "no_resolve_sig") ctx->opt.no_resolve_sig = value;
    else if (n == "no_fuse") ctx->opt.no_fuse = value;
    else if (n == "profile_upload") ctx->opt.profile_upload = value;
    else if (n == "chunks") {
        if (value < 1 || value > GLX_MAX_CHUNKS) return GLX_E_INVALID;
        ctx->opt.chunks = value;
    } else {
        ctx->err = "glx_set_option: unknown option " + n;
        return GLX_E_INVALID;
    }
    return GLX_OK;
}

// durations of the two upload-side kernels of a batch uploaded with option profile_upload set
int glx_upload_kernel_ms(glx_ctx* ctx, glx_dev_batch* d, float ms[2]) {
    if (!ctx || !d || !ms || !d->up_ev[0]) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    ms[0] = ms[1] = 0.f;
    CU(cudaEventSynchronize(d->up_ev[1]));
    CU(cudaEventElapsedTime(&ms[0], d->up_ev[0], d->up_ev[1]));
    if (d->up_ev[3]) {
        CU(cudaEventSynchronize(d->up_ev[3]));
        CU(cudaEventElapsedTime(&ms[1], d->up_ev[2], d->up_ev[3]));
    }
    return GLX_OK;
}

int glx_set_profile(glx_ctx* ctx, glx_dev_batch* d, int on) {
    if (!ctx || !d) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    d->profile = on != 0;
    if (d->profile)
        for (auto& ev : d->ev)
            if (!ev) CU(cudaEventCreate(&ev));
    return GLX_OK;
}

int glx_kernel_ms(glx_ctx* ctx, glx_dev_batch* d, float ms[4]) {
    if (!ctx || !d || !d->profile || !ms) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(d->ev[4]));
    for (int i = 0; i < 4; i++) CU(cudaEventElapsedTime(&ms[i], d->ev[i], d->ev[i + 1]));
    return GLX_OK;
}

// cells the tiled kernel handed to the worklist kernels in the last launch (synchronises the context stream)
int glx_work_count(glx_ctx* ctx, glx_dev_batch* d, uint64_t* n) {
    if (!ctx || !d || !n) return GLX_E_INVALID;
    *n = 0;
    if (!d->plan.enabled || !d->n_cells) return GLX_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t n_regions = (size_t)d->n_sample_blocks * d->n_site_tiles;
    std::vector<uint32_t> h(n_regions);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h.data(), d->n_work, n_regions * 4, cudaMemcpyDeviceToHost));
    for (uint32_t v : h) *n += v & GLX_NWORK_MASK;
    return GLX_OK;
}

// of those, the cells the tiled kernel finished itself (fused lone-variant-record closed form)
int glx_fused_count(glx_ctx* ctx, glx_dev_batch* d, uint64_t* n) {
    if (!ctx || !d || !n) return GLX_E_INVALID;
    *n = 0;
    if (!d->plan.enabled || !d->n_cells) return GLX_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t n_regions = (size_t)d->n_sample_blocks * d->n_site_tiles;
    std::vector<uint32_t> h(n_regions);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h.data(), d->n_work, n_regions * 4, cudaMemcpyDeviceToHost));
    for (uint32_t v : h) *n += v >> 16;
    if (d->decl_n) {  // the high halves also count what glx_cells_variant_list_kernel finished
        uint32_t c[2 * GLX_MAX_CHUNKS];
        CU(cudaMemcpy(c, d->decl_n, sizeof c, cudaMemcpyDeviceToHost));
        for (int j = 0; j < GLX_MAX_CHUNKS; j++) *n -= c[GLX_MAX_CHUNKS + j];
    }
    return GLX_OK;
}

uint64_t glx_out_bytes(const glx_dev_batch* d) {
    if (!d) return 0;
    uint64_t b = (uint64_t)d->n_cells * 4 + (uint64_t)d->n_sites * 5;
    for (uint32_t k = 0; k < d->n_fields; k++) b += d->fld_total[k] * 4;
    return b;
}
uint64_t glx_in_bytes(const glx_dev_batch* d) { return d ? d->in_bytes : 0; }
int glx_launch_count(const glx_dev_batch* d) { return d ? d->launches_per_call : 0; }

int glx_genotype_batch(glx_ctx* ctx, const glx_config* cfg, const glx_records* recs, const glx_sites* sites, glx_out* out) {
    glx_dev_batch* d = nullptr;
    int rc = glx_upload(ctx, cfg, recs, sites, &d);
    if (rc != GLX_OK) return rc;
    rc = glx_launch(ctx, d, nullptr);
    if (rc == GLX_OK) rc = glx_check(ctx, d);
    if (rc == GLX_OK) rc = glx_download(ctx, d, out);
    glx_free_batch(ctx, d);
    return rc;
}
}
