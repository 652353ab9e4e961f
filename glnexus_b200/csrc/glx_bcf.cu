// glx_bcf.cu — kernels and C ABI of the on-device BCF2 record encoder (stages in glx_bcf.cuh,
// BGZF/DEFLATE in glx_bgzf.cuh). sm_100a only; no host fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

// Phase clock of the BGZF kernel (thread 0 of every block; a dozen clock reads and atomics per 64 KB block): where a block's
// time goes, for profiles/ — glx_debug_bgzf_phases. [0, 10): the kernel's phases; [10, 16): parts of the gather (builds with
// -DGLX_FILL_PROFILE).
__device__ unsigned long long glx_bgzf_phase_cycles[16];
#define GLX_FILL_PROFILE_HERE 1

#include "glx_bgzf.cuh"
#include "glx_dev_state.cuh"

using namespace glxd;

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) glx_bcf_init_kernel(BcfEnc E) {
    const size_t n = (size_t)E.n_sites * E.n_fields;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        E.mm[2 * i] = INT32_MAX;
        E.mm[2 * i + 1] = INT32_MIN + 1;
    }
}

// grid = (sites, sample chunks): block (s, c) folds samples [c*chunk, (c+1)*chunk) of site s into the site's min/max table
__global__ void __launch_bounds__(256) glx_bcf_scan_kernel(BcfEnc E, uint32_t chunk) {
    __shared__ int32_t s_mm[2 * GLX_MAX_FIELDS];
    const uint32_t s = blockIdx.x, n0 = blockIdx.y * chunk;
    const uint32_t n1 = min(E.n_samples, n0 + chunk);
    bcf_scan_init(E, s_mm, threadIdx.x, blockDim.x);
    __syncthreads();
    bcf_scan_accum(E, s, n0, n1, s_mm, threadIdx.x, blockDim.x);
    __syncthreads();
    bcf_scan_flush(E, s, s_mm, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(128) glx_bcf_layout_kernel(BcfEnc E) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < E.n_sites) bcf_layout_site(E, s);
}

// one block: exclusive scans of rec_off[0..S] and hdr_off[0..S] in place (entry S = the totals), tile by tile with coalesced
// loads; totals[4] counts the sites that have a record
#define GLX_BCF_SCAN_THREADS 1024
__global__ void __launch_bounds__(GLX_BCF_SCAN_THREADS) glx_bcf_offsets_kernel(BcfEnc E, unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long s_r[32], s_h[32];
    __shared__ uint32_t s_n[32];
    __shared__ unsigned long long s_carry[3];
    const uint32_t n = E.n_sites, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry[0] = s_carry[1] = s_carry[2] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += GLX_BCF_SCAN_THREADS) {
        const uint32_t i = base + tid;
        const unsigned long long r = i < n ? E.rec_off[i] : 0ull, h = i < n ? E.hdr_off[i] : 0ull;
        unsigned long long ir = r, ih = h;
        uint32_t in_ = r != 0;
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const unsigned long long a = __shfl_up_sync(0xffffffffu, ir, d), b = __shfl_up_sync(0xffffffffu, ih, d);
            const uint32_t c = __shfl_up_sync(0xffffffffu, in_, d);
            if (lane >= d) {
                ir += a;
                ih += b;
                in_ += c;
            }
        }
        if (lane == 31) {
            s_r[wid] = ir;
            s_h[wid] = ih;
            s_n[wid] = in_;
        }
        __syncthreads();
        unsigned long long br = s_carry[0], bh = s_carry[1];
        for (uint32_t k = 0; k < wid; k++) {
            br += s_r[k];
            bh += s_h[k];
        }
        if (i < n) {
            E.rec_off[i] = br + ir - r;
            E.hdr_off[i] = bh + ih - h;
        }
        __syncthreads();
        if (tid == GLX_BCF_SCAN_THREADS - 1) {
            uint32_t cn = 0;
            for (uint32_t k = 0; k < 32; k++) cn += s_n[k];
            s_carry[0] = br + ir;
            s_carry[1] = bh + ih;
            s_carry[2] += cn;
        }
        __syncthreads();
    }
    if (tid == 0) {
        E.rec_off[n] = s_carry[0];
        E.hdr_off[n] = s_carry[1];
        totals[0] = s_carry[0];
        totals[1] = s_carry[1];
        totals[4] = s_carry[2];
    }
}

__global__ void __launch_bounds__(128) glx_bcf_headers_kernel(BcfEnc E) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < E.n_sites) bcf_header_site(E, s);
}

// first site of every span (thread per span), so that a span's block starts with one load instead of a binary search
__global__ void __launch_bounds__(256) glx_bcf_span_index_kernel(BcfEnc E, uint32_t* __restrict__ span_site, uint32_t n_spans) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_spans) return;
    const unsigned long long lo = (unsigned long long)b * GLX_BCF_SPAN;
    span_site[b] = lo < E.rec_off[E.n_sites] ? bcf_first_site(E, lo) : E.n_sites;
}

// One CTA per 65,280-byte span of the record stream: gather it in shared memory, then one bulk (TMA) store.
#define GLX_BCF_SPAN_THREADS 512
struct RawSmem {
    uint8_t buf[GLX_BCF_SPAN];
    BcfSpanTab tab;
};
__global__ void __launch_bounds__(GLX_BCF_SPAN_THREADS, 3) glx_bcf_span_raw_kernel(BcfEnc E, const uint32_t* __restrict__ span_site, uint8_t* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    RawSmem& M = *reinterpret_cast<RawSmem*>(s_raw);
    const unsigned long long total = E.rec_off[E.n_sites];
    const unsigned long long lo = (unsigned long long)blockIdx.x * GLX_BCF_SPAN;
    if (lo >= total) return;
    const unsigned long long hi = min(total, lo + GLX_BCF_SPAN);
    bcf_fill_span<BcfLinear>(E, lo, hi, span_site[blockIdx.x], M.buf, M.tab, threadIdx.x, blockDim.x);
    const uint32_t n = (uint32_t)(hi - lo), n16 = n & ~15u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the bulk-copy engine
    __syncthreads();
    if (threadIdx.x == 0 && n16) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(M.buf);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + lo), "r"(src), "r"(n16) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the source may be released once it has been read
    }
    for (uint32_t i = n16 + threadIdx.x; i < n; i += blockDim.x) out[lo + i] = M.buf[i];  // tail of the stream's last span
}

#define GLX_PHASE(i)                                                                    \
    do {                                                                                \
        if (tid == 0) {                                                                 \
            const long long now_ = clock64();                                           \
            atomicAdd(&glx_bgzf_phase_cycles[i], (unsigned long long)(now_ - t_phase)); \
            t_phase = now_;                                                             \
        }                                                                               \
    } while (0)

#define GLX_BGZF_THREADS 1024
struct BgzfSampleSmem {
    uint8_t in[(GLX_BGZF_IN_PAD + 15u) & ~15u];
    BgzfShared sh;
    BcfSpanTab tab;
};

// Sampling: GLX_BGZF_SAMPLES spans spread over the stream are gathered and tokenized; their symbol counts go to hist[320].
__global__ void __launch_bounds__(GLX_BGZF_THREADS, 1) glx_bgzf_sample_kernel(BcfEnc E, const uint32_t* __restrict__ span_site, uint32_t* __restrict__ hist) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    BgzfSampleSmem& M = *reinterpret_cast<BgzfSampleSmem*>(s_raw);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const unsigned long long total = E.rec_off[E.n_sites];
    const unsigned long long n_spans = (total + GLX_BCF_SPAN - 1) / GLX_BCF_SPAN;
    if (blockIdx.x >= n_spans) return;
    const unsigned long long b = n_spans <= gridDim.x ? blockIdx.x : (unsigned long long)blockIdx.x * n_spans / gridDim.x;
    const unsigned long long lo = b * GLX_BCF_SPAN, hi = min(total, lo + GLX_BCF_SPAN);
    const uint32_t n = (uint32_t)(hi - lo);
    for (uint32_t i = tid; i < GLX_BGZF_HIST; i += nt) M.sh.freq[i] = 0;
    bcf_fill_span<BcfPadded>(E, lo, hi, span_site[b], M.in, M.tab, tid, nt);
    bgzf_tokenize<true, false>(M.sh, M.tab, M.in, n, nullptr, tid, nt);
    __syncthreads();
    for (uint32_t i = tid; i < GLX_BGZF_HIST; i += nt)
        if (M.sh.freq[i]) atomicAdd(&hist[i], M.sh.freq[i]);
}

__global__ void __launch_bounds__(320) glx_bgzf_table_kernel(const uint32_t* __restrict__ hist, BgzfTable* __restrict__ tab) {
    __shared__ BgzfTable s_tab;
    __shared__ uint32_t s_scratch[64];
    bgzf_table_build(hist, s_tab, s_scratch, threadIdx.x, blockDim.x);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < sizeof(BgzfTable) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tab)[i] = reinterpret_cast<const uint32_t*>(&s_tab)[i];
}

// One CTA per span: gather the span into shared memory, CRC-32 and DEFLATE under the batch's code there, and write the codes
// straight into the block's slot in global memory (interior words of a worker's bit range with plain stores, the boundary
// words it shares with its neighbours with atomic ORs onto zeroed words), then the BGZF framing around them. Without an
// output stage the CTA needs ~112 KB of shared memory and 512 threads x 64 registers, so two spans are resident per SM:
// one's gather (waiting on HBM) and barriers overlap the other's tokenizer and emitter (ALU and shared memory).
#define GLX_BGZF_SPAN_THREADS 512
struct BgzfSpanSmem {
    uint8_t in[(GLX_BGZF_IN_PAD + 15u) & ~15u];  // BcfPadded layout
    BgzfShared sh;
    BcfSpanTab tab;
};
__global__ void __launch_bounds__(GLX_BGZF_SPAN_THREADS, 2) glx_bgzf_span_kernel(BcfEnc E, const uint32_t* __restrict__ span_site, const BgzfTable* __restrict__ gtab,
                                                                         const uint32_t* __restrict__ crc_pw, uint8_t* __restrict__ slots,
                                                                         uint32_t* __restrict__ sizes) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    BgzfSpanSmem& M = *reinterpret_cast<BgzfSpanSmem*>(s_raw);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const unsigned long long total = E.rec_off[E.n_sites];
    const unsigned long long lo = (unsigned long long)blockIdx.x * GLX_BCF_SPAN;
    if (lo >= total) {
        if (tid == 0) sizes[blockIdx.x] = 0;
        return;
    }
    const unsigned long long hi = min(total, lo + GLX_BCF_SPAN);
    const uint32_t n = (uint32_t)(hi - lo);
    uint8_t* const slot = slots + (size_t)blockIdx.x * GLX_BGZF_SLOT;
    uint32_t* const out_words = reinterpret_cast<uint32_t*>(slot + GLX_BGZF_LEAD + 18);  // 32 bytes into the (aligned) slot
    long long t_phase = clock64();
    for (uint32_t i = tid; i < sizeof(BgzfTable) / 4; i += nt) reinterpret_cast<uint32_t*>(&M.sh.tab)[i] = reinterpret_cast<const uint32_t*>(gtab)[i];
    if (tid == 0) M.sh.n_in = n;
    bgzf_crc_setup(M.sh, tid, nt);
    bcf_fill_span<BcfPadded>(E, lo, hi, span_site[blockIdx.x], M.in, M.tab, tid, nt);  // ends with a barrier
    GLX_PHASE(0);
    bgzf_tokenize<false, true>(M.sh, M.tab, M.in, n, crc_pw, tid, nt);
    __syncthreads();
    GLX_PHASE(4);
    if (tid == 0) bgzf_crc_finish(M.sh, crc_pw);
    {  // exclusive scan of the GLX_BGZF_WORKERS bit counts (+ header bits): K consecutive workers per thread, warp scans, warp totals
        constexpr uint32_t K = GLX_BGZF_WORKERS / GLX_BGZF_SPAN_THREADS;
        static_assert(K * GLX_BGZF_SPAN_THREADS == GLX_BGZF_WORKERS && GLX_BGZF_SPAN_THREADS / 32 <= 32, "scan shape");
        __shared__ uint32_t s_wsum[32];
        uint32_t mine[K], sum = 0;
#pragma unroll
        for (uint32_t k = 0; k < K; k++) {
            mine[k] = M.sh.wbits[tid * K + k];
            sum += mine[k];
        }
        uint32_t incl = sum;
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += a;
        }
        if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            uint32_t v = tid < GLX_BGZF_SPAN_THREADS / 32 ? s_wsum[tid] : 0u;
            for (uint32_t d = 1; d < 32; d <<= 1) {
                const uint32_t a = __shfl_up_sync(0xffffffffu, v, d);
                if (tid >= d) v += a;
            }
            s_wsum[tid] = v;
        }
        __syncthreads();
        uint32_t acc = ((tid >> 5) ? s_wsum[(tid >> 5) - 1] : 0u) + incl - sum + M.sh.tab.hdr_bits;
#pragma unroll
        for (uint32_t k = 0; k < K; k++) {
            M.sh.wbits[tid * K + k] = acc;
            acc += mine[k];
        }
        if (tid == GLX_BGZF_SPAN_THREADS - 1) M.sh.data_bits = acc;
    }
    __syncthreads();
    {  // zero the words the codes (and the end-of-block code) will be OR-ed into
        const uint32_t nz = ((M.sh.data_bits + 15u + 31u) >> 5) + 1u;
        for (uint32_t i = tid; i < nz; i += nt) out_words[i] = 0u;
    }
    __syncthreads();
    GLX_PHASE(7);
    bgzf_put_header(M.sh, out_words, tid, nt);
    bgzf_emit(M.sh, M.tab, M.in, n, out_words, tid, nt);
    __syncthreads();
    GLX_PHASE(8);
    if (tid == 0) bgzf_finish_deflate(M.sh, out_words);
    __syncthreads();
    if (M.sh.stored) {
        bgzf_store_raw(M.sh, M.in, slot, tid, nt);
        __syncthreads();
    }
    if (tid == 0) {
        bgzf_frame(M.sh, slot);
        sizes[blockIdx.x] = M.sh.block_len;
    }
    GLX_PHASE(9);
}

// exclusive scan of the block sizes (one block); totals[2] = BGZF bytes, totals[3] = blocks with data
__global__ void __launch_bounds__(1024) glx_bgzf_offsets_kernel(const uint32_t* __restrict__ sizes, uint32_t n, unsigned long long* __restrict__ offsets,
                                                                unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long s_sum[1024];
    __shared__ uint32_t s_cnt[1024];
    const uint32_t tid = threadIdx.x, per = (n + 1023) / 1024;
    const uint32_t i0 = min(n, tid * per), i1 = min(n, i0 + per);
    unsigned long long sum = 0;
    uint32_t cnt = 0;
    for (uint32_t i = i0; i < i1; i++) {
        sum += sizes[i];
        cnt += sizes[i] != 0;
    }
    s_sum[tid] = sum;
    s_cnt[tid] = cnt;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        const unsigned long long a = tid >= d ? s_sum[tid - d] : 0;
        const uint32_t c = tid >= d ? s_cnt[tid - d] : 0;
        __syncthreads();
        s_sum[tid] += a;
        s_cnt[tid] += c;
        __syncthreads();
    }
    unsigned long long base = tid ? s_sum[tid - 1] : 0;
    for (uint32_t i = i0; i < i1; i++) {
        offsets[i] = base;
        base += sizes[i];
    }
    if (tid == 1023) {
        totals[2] = s_sum[tid];
        totals[3] = s_cnt[tid];
    }
}

// block b: slot -> its place in the contiguous output
__global__ void __launch_bounds__(256) glx_bgzf_pack_kernel(const uint8_t* __restrict__ slots, const uint32_t* __restrict__ sizes,
                                                            const unsigned long long* __restrict__ offsets, uint8_t* __restrict__ packed) {
    const uint32_t len = sizes[blockIdx.x];
    if (!len) return;
    const uint8_t* src = slots + (size_t)blockIdx.x * GLX_BGZF_SLOT + GLX_BGZF_LEAD;
    uint8_t* dst = packed + offsets[blockIdx.x];
    // bytes up to the destination's next 4-byte boundary, then words (the source is misaligned by a constant), then the tail
    const uint32_t head = min(len, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3));
    for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
    const uint32_t nw = (len - head) / 4;
    const uint8_t* s0 = src + head;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(s0) & 3);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s0 - mis);
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
    for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) {
        const uint32_t a = sw[i];
        dw[i] = mis ? __funnelshift_r(a, sw[i + 1], 8 * mis) : a;
    }
    for (uint32_t i = head + nw * 4 + threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
}

__global__ void glx_publish_words_kernel(unsigned long long* __restrict__ host_words, const unsigned long long* __restrict__ src, uint32_t n) {
    if (threadIdx.x < n) host_words[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

static int glx_bgzf_launch(glx_ctx* ctx, const BcfEnc& E, const uint32_t* span_site, GlxBcfState& B, uint32_t n_spans, cudaStream_t st) {
    if (!(ctx->func_attrs & 1)) {  // per device: a process may drive several
        CU(cudaFuncSetAttribute(glx_bgzf_span_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BgzfSpanSmem)));
        CU(cudaFuncSetAttribute(glx_bgzf_span_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CU(cudaFuncSetAttribute(glx_bgzf_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BgzfSampleSmem)));
        ctx->func_attrs |= 1;
    }
    CU(cudaMemsetAsync(B.hist, 0, GLX_BGZF_HIST * 4, st));
    glx_bgzf_sample_kernel<<<std::min<uint32_t>(n_spans, GLX_BGZF_SAMPLES), GLX_BGZF_THREADS, sizeof(BgzfSampleSmem), st>>>(E, span_site, B.hist);
    glx_bgzf_table_kernel<<<1, 320, 0, st>>>(B.hist, B.table);
    glx_bgzf_span_kernel<<<n_spans, GLX_BGZF_SPAN_THREADS, sizeof(BgzfSpanSmem), st>>>(E, span_site, B.table, ctx->crc_pw, B.stream, B.blk_sizes);
    glx_bgzf_offsets_kernel<<<1, 1024, 0, st>>>(B.blk_sizes, n_spans, B.blk_offsets, B.totals);
    glx_bgzf_pack_kernel<<<n_spans, 256, 0, st>>>(B.stream, B.blk_sizes, B.blk_offsets, B.packed);
    CU(cudaGetLastError());
    return GLX_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
namespace {
struct Arena2 {
    char* base;
    size_t used = 0;
    template <class T>
    T* take(size_t n) {
        used = (used + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + used);
        used += n * sizeof(T);
        return p;
    }
};
size_t pad256(size_t b) { return ((b + 255) & ~(size_t)255) + 256; }
}  // namespace

extern "C" {

int glx_bcf_attach(glx_ctx* ctx, glx_dev_batch* d, const glx_sites* sites, const glx_bcf_meta* meta, const glx_config* cfg, void* stream_) {
    if (!ctx || !d || !sites || !meta || !cfg) return GLX_E_INVALID;
    if (sites->n_sites != d->n_sites || sites->n_samples != d->n_samples || meta->n_sites != d->n_sites || cfg->n_fields != d->n_fields) {
        ctx->err = "glx_bcf_attach: sites / meta do not belong to this batch";
        return GLX_E_INVALID;
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    GlxBcfState& B = d->bcf;
    const uint64_t S = d->n_sites, N = d->n_samples, F = d->n_fields;
    const uint64_t nA = sites->n_site_alleles;
    // ---- upper bounds of the header pool and of the record stream (widths <= 4 bytes, every allele kept) ----
    uint64_t ft_max = 1;
    for (uint32_t i = 0; i < meta->n_filters; i++) ft_max += meta->filter_off[i + 1] - meta->filter_off[i] + 1;
    uint64_t hdr_bound = 0, indiv_bound = 0;
    for (uint64_t s = 0; s < S; s++) {
        const uint32_t A = sites->n_allele[s], a0 = sites->allele_off[s];
        const uint64_t ctg = meta->contig_off[meta->rid[s] + 1] - meta->contig_off[meta->rid[s]];
        uint64_t sh = 24 + 6 + 2 + 2 * (2 + 6 + 4ull * A);
        for (uint32_t a = 0; a < A; a++) {
            sh += 6 + (meta->dna_off[a0 + a + 1] - meta->dna_off[a0 + a]);
            if (a) sh += ctg + 14 + (uint64_t)(meta->norm_end[a0 + a] - meta->norm_beg[a0 + a]) + (meta->norm_off[a0 + a + 1] - meta->norm_off[a0 + a]);
        }
        hdr_bound += 8 + sh;
        uint64_t iv = (F + 2) * GLX_BCF_HDR_MAX + 4 * N;
        for (uint32_t k = 0; k < F; k++) {
            const bool str = cfg->fields[k].kind == GLX_FK_FT || cfg->fields[k].kind == GLX_FK_STRING;
            iv += str ? N * ft_max : N * (uint64_t)sites->fld_cnt[k][s] * 4;
        }
        indiv_bound += iv;
    }
    B.hdr_bound = hdr_bound;
    B.stream_bound = hdr_bound + indiv_bound;
    // ---- meta arena ----
    struct Item {
        const void* src;
        size_t bytes;
        const void** dst;
    };
    BcfEnc& E = B.E;
    memset(&E, 0, sizeof E);
    bcf_enc_scalars(E, *cfg, *meta, (uint32_t)S, (uint32_t)N);
    std::vector<Item> items;
    auto add = [&](const void* src, size_t bytes, auto** dst) { items.push_back(Item{src, bytes, (const void**)dst}); };
    add(meta->rid, S * 4, &E.m.rid);
    add(meta->qual, S * 4, &E.m.qual);
    add(meta->dna_off, (nA + 1) * 4, &E.m.dna_off);
    add(meta->dna_pool, meta->n_dna_pool, &E.m.dna_pool);
    add(meta->norm_beg, nA * 4, &E.m.norm_beg);
    add(meta->norm_end, nA * 4, &E.m.norm_end);
    add(meta->norm_off, (nA + 1) * 4, &E.m.norm_off);
    add(meta->norm_pool, meta->n_norm_pool, &E.m.norm_pool);
    add(meta->aq, nA * 4, &E.m.aq);
    add(sites->allele_af, nA * 4, &E.m.af);
    add(meta->contig_off, ((size_t)meta->n_contigs + 1) * 4, &E.m.contig_off);
    add(meta->contig_pool, meta->n_contig_pool, &E.m.contig_pool);
    add(meta->filter_off, ((size_t)meta->n_filters + 1) * 4, &E.m.filter_off);
    add(meta->filter_pool, meta->n_filter_pool, &E.m.filter_pool);
    size_t mb = 256;
    for (auto& it : items) mb += pad256(it.bytes);
    if (!B.meta_base || B.meta_bytes < mb) {
        if (B.meta_base) ctx->give(B.meta_base, B.meta_bytes);
        B.meta_base = nullptr;
        B.meta_bytes = mb;
        cudaError_t e = ctx->take((void**)&B.meta_base, &B.meta_bytes);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(bcf meta)");
    }
    Arena2 ar{B.meta_base};
    for (auto& it : items) {
        char* p = ar.take<char>(it.bytes ? it.bytes : 1);
        *it.dst = p;
        if (it.bytes) CU(cudaMemcpyAsync(p, it.src, it.bytes, cudaMemcpyHostToDevice, st));
    }
    // ---- working arena ----
    const size_t n_spans_b = (size_t)((B.stream_bound + GLX_BCF_SPAN - 1) / GLX_BCF_SPAN) + 1;
    const size_t wb = 256 + pad256(S * F * 8 + 16) + pad256((S + 1) * sizeof(BcfSitePlan)) + 2 * pad256((S + 1) * 8) + pad256(hdr_bound + 16) + pad256(64) +
                      2 * pad256(n_spans_b * 4) + pad256(n_spans_b * 8) + pad256(GLX_BGZF_HIST * 4) + pad256(sizeof(BgzfTable));
    if (!B.work_base || B.work_bytes < wb) {
        if (B.work_base) ctx->give(B.work_base, B.work_bytes);
        B.work_base = nullptr;
        B.work_bytes = wb;
        cudaError_t e = ctx->take((void**)&B.work_base, &B.work_bytes);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(bcf work)");
    }
    Arena2 wa{B.work_base};
    E.mm = wa.take<int32_t>(S * F * 2 + 4);
    E.plan = wa.take<BcfSitePlan>(S + 1);
    E.rec_off = wa.take<unsigned long long>(S + 1);
    E.hdr_off = wa.take<unsigned long long>(S + 1);
    E.hdr_pool = wa.take<uint8_t>(hdr_bound + 16);
    B.totals = wa.take<unsigned long long>(8);
    B.blk_sizes = wa.take<uint32_t>(n_spans_b);
    B.blk_offsets = wa.take<unsigned long long>(n_spans_b);
    B.span_site = wa.take<uint32_t>(n_spans_b);
    B.hist = wa.take<uint32_t>(GLX_BGZF_HIST);
    B.table = wa.take<BgzfTable>(1);
    if (!ctx->crc_pw) {  // the 1,025 CRC powers and their multiplication tables, once per context
        std::vector<uint32_t> pw(GLX_BGZF_CRC_WORDS);
        bgzf_crc_powers(pw.data());
        CU(cudaMalloc((void**)&ctx->crc_pw, pw.size() * 4));
        CU(cudaMemcpy(ctx->crc_pw, pw.data(), pw.size() * 4, cudaMemcpyHostToDevice));
    }
    // the batch's own device arrays
    E.beg = d->S.beg;
    E.n_allele = d->S.n_allele;
    E.monoallelic = d->S.monoallelic;
    E.allele_off = d->S.allele_off;
    for (uint32_t k = 0; k < F; k++) {
        E.fld_cnt[k] = d->S.fld_cnt[k];
        E.fld_off[k] = d->S.fld_off[k];
        E.fld[k] = d->O.fld[k];
    }
    E.gt = d->O.gt;
    E.rnc = d->O.rnc;
    E.allele_used = d->O.allele_used;
    if (!B.len_ev) CU(cudaEventCreateWithFlags(&B.len_ev, cudaEventDisableTiming));
    if (!B.meta_ev) CU(cudaEventCreateWithFlags(&B.meta_ev, cudaEventDisableTiming));
    CU(cudaEventRecord(B.meta_ev, st));  // the encoder's stream waits for this: the attach may ride a copy-only stream
    if (glx_take_err_slot(ctx, d) != GLX_OK) return cuda_fail(ctx, cudaErrorMemoryAllocation, "pinned status word");
    B.attached = true;
    return GLX_OK;
}

int glx_encode_bcf_on(glx_ctx* ctx, glx_dev_batch* d, int mode, void* stream_) {
    if (!ctx || !d || !d->bcf.attached || (mode != 0 && mode != 1)) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    GlxBcfState& B = d->bcf;
    const BcfEnc& E = B.E;
    const uint32_t S = d->n_sites, N = d->n_samples;
    B.mode = mode;
    CU(cudaStreamWaitEvent(st, B.meta_ev, 0));
    const uint32_t n_spans = (uint32_t)((B.stream_bound + GLX_BCF_SPAN - 1) / GLX_BCF_SPAN);
    const size_t need = mode == 0 ? B.stream_bound + 256 : (size_t)n_spans * GLX_BGZF_SLOT + 256;
    if (!B.stream || B.stream_bytes < need) {
        if (B.stream) ctx->give(B.stream, B.stream_bytes);
        B.stream = nullptr;
        B.stream_bytes = need;
        cudaError_t e = ctx->take((void**)&B.stream, &B.stream_bytes);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(bcf stream)");
    }
    if (mode == 1) {
        const size_t pneed = (size_t)n_spans * GLX_BGZF_SLOT + 256;
        if (!B.packed || B.packed_bytes < pneed) {
            if (B.packed) ctx->give(B.packed, B.packed_bytes);
            B.packed = nullptr;
            B.packed_bytes = pneed;
            cudaError_t e = ctx->take((void**)&B.packed, &B.packed_bytes);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(bgzf blocks)");
        }
    }
    CU(cudaMemsetAsync(B.totals, 0, 64, st));
    if (S && N) {
        const uint32_t chunk = 2048;
        glx_bcf_init_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)S * d->n_fields + 255) / 256 + 1, (uint64_t)ctx->sm_count * 8), 256, 0, st>>>(E);
        glx_bcf_scan_kernel<<<dim3(S, (N + chunk - 1) / chunk), 256, 0, st>>>(E, chunk);
        glx_bcf_layout_kernel<<<(S + 127) / 128, 128, 0, st>>>(E);
    } else {  // no cells: no records
        CU(cudaMemsetAsync(E.rec_off, 0, ((size_t)S + 1) * 8, st));
        CU(cudaMemsetAsync(E.hdr_off, 0, ((size_t)S + 1) * 8, st));
    }
    if (S && N) {
        glx_bcf_offsets_kernel<<<1, GLX_BCF_SCAN_THREADS, 0, st>>>(E, B.totals);
        glx_bcf_headers_kernel<<<(S + 127) / 128, 128, 0, st>>>(E);
        glx_bcf_span_index_kernel<<<(n_spans + 255) / 256, 256, 0, st>>>(E, B.span_site, n_spans);
        CU(cudaGetLastError());
        if (mode == 0) {
            const size_t smem = sizeof(RawSmem);
            if (!(ctx->func_attrs & 2)) {
                CU(cudaFuncSetAttribute(glx_bcf_span_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ctx->func_attrs |= 2;
            }
            glx_bcf_span_raw_kernel<<<n_spans, GLX_BCF_SPAN_THREADS, smem, st>>>(E, B.span_site, B.stream);
        } else {
            int rc = glx_bgzf_launch(ctx, E, B.span_site, B, n_spans, st);
            if (rc != GLX_OK) return rc;
        }
        CU(cudaGetLastError());
    }
    // lengths -> the batch's pinned words {., raw stream bytes, BGZF bytes, BGZF blocks}, stored by a kernel (the words are mapped
    // pinned memory): a copy-engine D2H here would queue the stream's next operation behind whatever download that engine is
    // busy with — the stream stays on the compute engine from its first kernel to this event
    glx_publish_words_kernel<<<1, 32, 0, st>>>(reinterpret_cast<unsigned long long*>(d->host_err + 1), B.totals, 5);
    CU(cudaGetLastError());
    CU(cudaEventRecord(B.len_ev, st));
    return GLX_OK;
}

int glx_bcf_length(glx_ctx* ctx, glx_dev_batch* d, uint64_t* n_bytes, uint64_t* n_raw_bytes, uint32_t* n_blocks) {
    if (!ctx || !d || !d->bcf.attached || d->bcf.mode < 0) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(d->bcf.len_ev));
    if (n_raw_bytes) *n_raw_bytes = d->host_err[1];
    if (n_bytes) *n_bytes = d->bcf.mode == 0 ? d->host_err[1] : d->host_err[3];
    if (n_blocks) *n_blocks = d->bcf.mode == 0 ? 0u : (uint32_t)d->host_err[4];
    return GLX_OK;
}

int glx_bcf_ready(glx_ctx* ctx, glx_dev_batch* d) {
    if (!ctx || !d || !d->bcf.attached || d->bcf.mode < 0) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    const cudaError_t e = cudaEventQuery(d->bcf.len_ev);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    return cuda_fail(ctx, e, "cudaEventQuery");
}

// records in the encoded batch (sites that survived allele trimming); valid after glx_bcf_length
uint64_t glx_bcf_records(const glx_dev_batch* d) { return d && d->bcf.attached && d->host_err ? d->host_err[5] : 0; }

int glx_download_bcf_on(glx_ctx* ctx, glx_dev_batch* d, void* dst, uint64_t capacity, void* stream_, int sync) {
    if (!ctx || !d || !dst) return GLX_E_INVALID;
    uint64_t n = 0;
    int rc = glx_bcf_length(ctx, d, &n, nullptr, nullptr);
    if (rc != GLX_OK) return rc;
    if (n > capacity) {
        ctx->err = "glx_download_bcf_on: destination too small";
        return GLX_E_INVALID;
    }
    cudaStream_t st = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    d->dl_stream = st;
    CU(cudaStreamWaitEvent(st, d->bcf.len_ev, 0));  // `stream` may be another one than the encoder's (a download-only stream)
    const uint8_t* src = d->bcf.mode == 0 ? d->bcf.stream : d->bcf.packed;
    if (n) CU(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(d->host_err, d->O.error, 8, cudaMemcpyDeviceToHost, st));
    if (sync) CU(cudaStreamSynchronize(st));
    return GLX_OK;
}

// record offsets in the raw stream (tests): rec_off[0..S]
int glx_bcf_record_offsets(glx_ctx* ctx, glx_dev_batch* d, uint64_t* rec_off) {
    if (!ctx || !d || !d->bcf.attached || !rec_off) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(d->bcf.len_ev));
    CU(cudaMemcpy(rec_off, d->bcf.E.rec_off, ((size_t)d->n_sites + 1) * 8, cudaMemcpyDeviceToHost));
    return GLX_OK;
}

// cycles thread 0 of the BGZF blocks spent per phase since the last call (gather, segments, setup, crc, tokenize, crc tree,
// codes + header, count + scan, emit, finish + store); out[15] unused
int glx_debug_bgzf_phases(glx_ctx* ctx, uint64_t out[16]) {
    if (!ctx || !out) return GLX_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    unsigned long long h[16], z[16] = {0};
    CU(cudaMemcpyFromSymbol(h, glx_bgzf_phase_cycles, sizeof h));
    CU(cudaMemcpyToSymbol(glx_bgzf_phase_cycles, z, sizeof z));
    for (int i = 0; i < 16; i++) out[i] = h[i];
    return GLX_OK;
}

uint64_t glx_bcf_bound(const glx_dev_batch* d) { return d && d->bcf.attached ? d->bcf.stream_bound + ((d->bcf.stream_bound / GLX_BCF_SPAN) + 2) * 64 : 0; }
}
