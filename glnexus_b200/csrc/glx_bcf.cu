// glx_bcf.cu — kernels and C ABI of the on-device BCF2 record encoder (stages in glx_bcf.cuh,
// BGZF/DEFLATE in glx_bgzf.cuh). sm_100a only; no host fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

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

// one block: exclusive scans of rec_off[0..S] and hdr_off[0..S] in place (entry S = the totals)
#define GLX_BCF_SCAN_THREADS 1024
__global__ void __launch_bounds__(GLX_BCF_SCAN_THREADS) glx_bcf_offsets_kernel(BcfEnc E, unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long s_r[GLX_BCF_SCAN_THREADS], s_h[GLX_BCF_SCAN_THREADS];
    const uint32_t n = E.n_sites, tid = threadIdx.x;
    const uint32_t per = (n + GLX_BCF_SCAN_THREADS - 1) / GLX_BCF_SCAN_THREADS;
    const uint32_t i0 = min(n, tid * per), i1 = min(n, i0 + per);
    unsigned long long r = 0, h = 0;
    uint32_t nrec = 0;
    for (uint32_t i = i0; i < i1; i++) {
        r += E.rec_off[i];
        h += E.hdr_off[i];
        nrec += E.rec_off[i] != 0;
    }
    if (nrec) atomicAdd(totals + 4, (unsigned long long)nrec);  // sites that have a record (zeroed before the encode)
    s_r[tid] = r;
    s_h[tid] = h;
    __syncthreads();
    for (uint32_t d = 1; d < GLX_BCF_SCAN_THREADS; d <<= 1) {  // Hillis-Steele inclusive scan of the per-thread sums
        const unsigned long long ar = tid >= d ? s_r[tid - d] : 0, ah = tid >= d ? s_h[tid - d] : 0;
        __syncthreads();
        s_r[tid] += ar;
        s_h[tid] += ah;
        __syncthreads();
    }
    unsigned long long br = tid ? s_r[tid - 1] : 0, bh = tid ? s_h[tid - 1] : 0;
    for (uint32_t i = i0; i < i1; i++) {
        const unsigned long long lr = E.rec_off[i], lh = E.hdr_off[i];
        E.rec_off[i] = br;
        E.hdr_off[i] = bh;
        br += lr;
        bh += lh;
    }
    if (tid == GLX_BCF_SCAN_THREADS - 1) {
        E.rec_off[n] = s_r[tid];
        E.hdr_off[n] = s_h[tid];
        totals[0] = s_r[tid];
        totals[1] = s_h[tid];
    }
}

__global__ void __launch_bounds__(128) glx_bcf_headers_kernel(BcfEnc E) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < E.n_sites) bcf_header_site(E, s);
}

// One CTA per 65,280-byte span of the record stream: gather it in shared memory, then one bulk (TMA) store.
#define GLX_BCF_SPAN_THREADS 256
__global__ void __launch_bounds__(GLX_BCF_SPAN_THREADS) glx_bcf_span_raw_kernel(BcfEnc E, uint8_t* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t s_buf[];  // [GLX_BCF_SPAN] span bytes, then the slot table
    uint16_t* s_src = reinterpret_cast<uint16_t*>(s_buf + GLX_BCF_SPAN);
    const unsigned long long total = E.rec_off[E.n_sites];
    const unsigned long long lo = (unsigned long long)blockIdx.x * GLX_BCF_SPAN;
    if (lo >= total) return;
    const unsigned long long hi = min(total, lo + GLX_BCF_SPAN);
    bcf_fill_span<BcfLinear>(E, lo, hi, s_buf, s_src, threadIdx.x, blockDim.x);
    const uint32_t n = (uint32_t)(hi - lo), n16 = n & ~15u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the bulk-copy engine
    __syncthreads();
    if (threadIdx.x == 0 && n16) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(s_buf);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + lo), "r"(src), "r"(n16) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the source may be released once it has been read
    }
    for (uint32_t i = n16 + threadIdx.x; i < n; i += blockDim.x) out[lo + i] = s_buf[i];  // tail of the stream's last span
}

// One CTA per span: gather the span, CRC-32, DEFLATE and BGZF framing in shared memory, one bulk store into the block's slot.
#define GLX_BGZF_THREADS 1024
#define GLX_BGZF_OUT_PAD ((GLX_BGZF_OUT + 15u) & ~15u)
struct BgzfSmem {
    uint8_t in[(GLX_BGZF_IN_PAD + 15u) & ~15u];  // BcfPadded layout
    uint8_t out[GLX_BGZF_OUT_PAD];
    BgzfShared sh;
    uint16_t src[GLX_BCF_G_MAX + 8];
};
__global__ void __launch_bounds__(GLX_BGZF_THREADS) glx_bgzf_span_kernel(BcfEnc E, uint8_t* __restrict__ slots, uint32_t* __restrict__ sizes) {
    extern __shared__ __align__(128) uint8_t s_raw[];
    BgzfSmem& M = *reinterpret_cast<BgzfSmem*>(s_raw);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const unsigned long long total = E.rec_off[E.n_sites];
    const unsigned long long lo = (unsigned long long)blockIdx.x * GLX_BCF_SPAN;
    if (lo >= total) {
        if (tid == 0) sizes[blockIdx.x] = 0;
        return;
    }
    const unsigned long long hi = min(total, lo + GLX_BCF_SPAN);
    const uint32_t n = (uint32_t)(hi - lo);
    uint32_t* out_words = reinterpret_cast<uint32_t*>(M.out + GLX_BGZF_LEAD + 18);
    for (uint32_t i = tid; i < GLX_BGZF_OUT_PAD / 4; i += nt) reinterpret_cast<uint32_t*>(M.out)[i] = 0;
    bcf_fill_span<BcfPadded>(E, lo, hi, M.in, M.src, tid, nt);
    if (tid == 0) {
        M.sh.n_in = n;
        bgzf_segments(M.sh, E, lo, hi);
    }
    bgzf_crc_setup(M.sh, n, tid, nt);
    bgzf_hist_init(M.sh, tid, nt);
    __syncthreads();
    bgzf_crc_parts(M.sh, M.in, n, tid, nt);
    bgzf_tokenize(M.sh, M.in, n, tid, nt);
    __syncthreads();
    for (uint32_t l = 0; l < 10; l++) {
        bgzf_crc_level(M.sh, l, tid, nt);
        __syncthreads();
    }
    if (tid == 0) bgzf_crc_finish(M.sh);
    __syncthreads();
    bgzf_build_codes(M.sh, out_words, tid, nt);
    bgzf_count_bits(M.sh, M.in, n, tid, nt);
    __syncthreads();
    {  // exclusive scan of the 1024 bit counts (+ header bits)
        const uint32_t mine = M.sh.wbits[tid];
        for (uint32_t d = 1; d < GLX_BGZF_WORKERS; d <<= 1) {
            const uint32_t a = tid >= d ? M.sh.wbits[tid - d] : 0;
            __syncthreads();
            M.sh.wbits[tid] += a;
            __syncthreads();
        }
        const uint32_t incl = M.sh.wbits[tid];
        __syncthreads();
        M.sh.wbits[tid] = incl - mine + M.sh.hdr_bits;
        if (tid == GLX_BGZF_WORKERS - 1) M.sh.data_bits = incl + M.sh.hdr_bits;
    }
    __syncthreads();
    bgzf_emit(M.sh, M.in, n, out_words, tid, nt);
    __syncthreads();
    if (tid == 0) bgzf_finish_deflate(M.sh, out_words);
    __syncthreads();
    if (M.sh.stored) {
        bgzf_store_raw(M.sh, M.in, M.out, tid, nt);
        __syncthreads();
    }
    if (tid == 0) bgzf_frame(M.sh, M.out);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (GLX_BGZF_LEAD + M.sh.block_len + 15u) & ~15u;
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(M.out);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(slots + (size_t)blockIdx.x * GLX_BGZF_SLOT), "r"(src), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        sizes[blockIdx.x] = M.sh.block_len;
    }
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

static int glx_bgzf_launch(glx_ctx* ctx, const BcfEnc& E, uint8_t* slots, uint8_t* packed, unsigned long long* totals, uint32_t* sizes,
                           unsigned long long* offsets, uint32_t n_spans, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        CU(cudaFuncSetAttribute(glx_bgzf_span_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BgzfSmem)));
        attr_set = true;
    }
    glx_bgzf_span_kernel<<<n_spans, GLX_BGZF_THREADS, sizeof(BgzfSmem), st>>>(E, slots, sizes);
    glx_bgzf_offsets_kernel<<<1, 1024, 0, st>>>(sizes, n_spans, offsets, totals);
    glx_bgzf_pack_kernel<<<n_spans, 256, 0, st>>>(slots, sizes, offsets, packed);
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
        cudaError_t e = ctx->take((void**)&B.meta_base, mb);
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
                      pad256(n_spans_b * 4) + pad256(n_spans_b * 8);
    if (!B.work_base || B.work_bytes < wb) {
        if (B.work_base) ctx->give(B.work_base, B.work_bytes);
        B.work_base = nullptr;
        B.work_bytes = wb;
        cudaError_t e = ctx->take((void**)&B.work_base, wb);
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
    const uint32_t n_spans = (uint32_t)((B.stream_bound + GLX_BCF_SPAN - 1) / GLX_BCF_SPAN);
    const size_t need = mode == 0 ? B.stream_bound + 256 : (size_t)n_spans * GLX_BGZF_SLOT + 256;
    if (!B.stream || B.stream_bytes < need) {
        if (B.stream) ctx->give(B.stream, B.stream_bytes);
        B.stream = nullptr;
        B.stream_bytes = need;
        cudaError_t e = ctx->take((void**)&B.stream, need);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(bcf stream)");
    }
    if (mode == 1) {
        const size_t pneed = (size_t)n_spans * GLX_BGZF_SLOT + 256;
        if (!B.packed || B.packed_bytes < pneed) {
            if (B.packed) ctx->give(B.packed, B.packed_bytes);
            B.packed = nullptr;
            B.packed_bytes = pneed;
            cudaError_t e = ctx->take((void**)&B.packed, pneed);
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
        CU(cudaGetLastError());
        if (mode == 0) {
            const size_t smem = GLX_BCF_SPAN + (GLX_BCF_G_MAX + 8) * 2;
            static bool attr_set = false;
            if (!attr_set) {
                CU(cudaFuncSetAttribute(glx_bcf_span_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_set = true;
            }
            glx_bcf_span_raw_kernel<<<n_spans, GLX_BCF_SPAN_THREADS, smem, st>>>(E, B.stream);
        } else {
            int rc = glx_bgzf_launch(ctx, E, B.stream, B.packed, B.totals, B.blk_sizes, B.blk_offsets, n_spans, st);
            if (rc != GLX_OK) return rc;
        }
        CU(cudaGetLastError());
    }
    // lengths -> the batch's pinned words {., raw stream bytes, BGZF bytes, BGZF blocks}
    CU(cudaMemcpyAsync(d->host_err + 1, B.totals, 40, cudaMemcpyDeviceToHost, st));
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

uint64_t glx_bcf_bound(const glx_dev_batch* d) { return d && d->bcf.attached ? d->bcf.stream_bound + ((d->bcf.stream_bound / GLX_BCF_SPAN) + 2) * 64 : 0; }
}
