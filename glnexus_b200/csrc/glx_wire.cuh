// glx_wire.cuh — device side of the compact wire form (glx_wire, include/glx.h): expansion kernels that rebuild the
// kernel-facing record store (DRecs) from the narrow columns. Offsets are prefix sums: one three-value scan over the records
// (non-reference ordinal, allele descriptor offset, value-pool offset) and one over the allele lengths.
#pragma once
#include "glx_dev_types.cuh"

namespace glxd {

struct DWire {
    uint64_t n_records, n_variant, n_al, n_al_pool, n_pool, n_exc;
    uint32_t n_cols, flags;
    uint8_t col_w[GLX_MAX_COLS];
    const uint8_t* base;  // device copy of the whole wire buffer
    uint64_t off[GLX_WS_COUNT];
    uint64_t col_val_off[GLX_MAX_COLS];  // byte offset of column c's scalar values
    // GLX_WF_SHAPES: the four per-record sections below are not on the wire; glx_wire_shapes_kernel writes them here first
    const uint8_t *x_flags, *x_nal, *x_gt8, *x_collen8;
    uint32_t n_shapes;
    uint64_t n_shape_exc;
    __host__ __device__ const uint8_t* sec_flags() const { return x_flags ? x_flags : base + off[GLX_WS_FLAGS]; }
    __host__ __device__ const uint8_t* sec_nal() const { return x_nal ? x_nal : base + off[GLX_WS_NALLELE]; }
    __host__ __device__ const uint8_t* sec_gt() const { return x_gt8 ? x_gt8 : base + off[GLX_WS_GT]; }
    __host__ __device__ const uint8_t* sec_collen() const { return x_collen8 ? x_collen8 : base + off[GLX_WS_COL_LEN]; }
};

#define GLX_WIRE_BLOCK 256
#define GLX_WIRE_ITEMS 4  // records per thread: 1,024 per block

struct uint3x {
    uint32_t a, b, c;
};
__device__ __forceinline__ uint3x add3(uint3x x, uint3x y) { return uint3x{x.a + y.a, x.b + y.b, x.c + y.c}; }

__device__ __forceinline__ uint32_t wire_al_len(const DWire& W, uint64_t a) {
    const uint8_t* p = W.base + W.off[GLX_WS_AL_LEN];
    if (W.flags & GLX_WF_ALLEN8) return p[a];
    return (W.flags & GLX_WF_ALLEN16) ? reinterpret_cast<const uint16_t*>(p)[a] : reinterpret_cast<const uint32_t*>(p)[a];
}
__device__ __forceinline__ uint32_t wire_filt(const DWire& W, uint64_t i) {
    const uint8_t* p = W.base + W.off[GLX_WS_FILT];
    return (W.flags & GLX_WF_FILT8) ? (uint32_t)p[i] : reinterpret_cast<const uint32_t*>(p)[i];
}
__device__ __forceinline__ uint32_t wire_col_len(const DWire& W, uint64_t c, uint64_t r) {
    if (W.flags & GLX_WF_COLLEN8) {
        const uint8_t v = W.sec_collen()[c * W.n_records + r];
        return v >= 0xFD ? 0xFF00u | v : v;
    }
    return reinterpret_cast<const uint16_t*>(W.sec_collen())[c * W.n_records + r];
}
// what record r adds to the three running sums
__device__ __forceinline__ uint3x wire_record_units(const DWire& W, uint64_t r) {
    const uint8_t fl = W.sec_flags()[r];
    const bool var = !(fl & GLX_RF_IS_REF);
    uint32_t pool = 0;
    for (uint32_t c = 0; c < W.n_cols; c++) {
        const uint32_t len = wire_col_len(W, c, r);
        if (len < 0xFFF0 && len != 1) pool += len;
    }
    return uint3x{var ? 1u : 0u, var ? (uint32_t)W.sec_nal()[r] : 0u, pool};
}

// GLX_WF_SHAPES, pass 0: a record's shape code -> its (flags, n_allele, GT pair, column lengths) in the layout the sections have
// without shapes (uint8 columns), so that everything after reads them as if they had travelled. The table (<= 255 rows of
// 4 + n_cols bytes) sits in shared memory; a record with code 0xFF finds its own row by its index in the sorted exception list.
__global__ void __launch_bounds__(256) glx_wire_shapes_kernel(const DWire W, uint8_t* __restrict__ flags, uint8_t* __restrict__ nal, uint8_t* __restrict__ gt8,
                                                              uint8_t* __restrict__ collen8) {
    __shared__ uint8_t s_table[255 * (4 + GLX_MAX_COLS)];
    const uint32_t row = 4 + W.n_cols;
    const uint8_t* table = W.base + W.off[GLX_WS_SHAPE_TABLE];
    for (uint32_t i = threadIdx.x; i < W.n_shapes * row; i += blockDim.x) s_table[i] = table[i];
    __syncthreads();
    const uint8_t* code = W.base + W.off[GLX_WS_SHAPE];
    const uint32_t* exc_idx = reinterpret_cast<const uint32_t*>(W.base + W.off[GLX_WS_SHAPE_EXC_IDX]);
    const uint8_t* exc_rows = W.base + W.off[GLX_WS_SHAPE_EXC];
    const uint64_t R = W.n_records;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = code[r];
        const uint8_t* p;
        if (k != 0xFF) p = s_table + k * row;
        else {
            uint64_t lo = 0, hi = W.n_shape_exc;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (exc_idx[mid] < (uint32_t)r) lo = mid + 1;
                else hi = mid;
            }
            p = exc_rows + lo * row;
        }
        flags[r] = p[0];
        nal[r] = p[1];
        gt8[2 * r] = p[2];
        gt8[2 * r + 1] = p[3];
        for (uint32_t c = 0; c < W.n_cols; c++) collen8[(size_t)c * R + r] = p[4 + c];
    }
}

// block-wide exclusive scan of one uint3x per thread (256 threads); returns the thread's exclusive prefix, total in *tot
__device__ __forceinline__ uint3x block_scan3(uint3x v, uint3x* tot) {
    __shared__ uint3x s_w[GLX_WIRE_BLOCK / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint3x inc = v;
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, inc.a, d), b = __shfl_up_sync(0xffffffffu, inc.b, d), c = __shfl_up_sync(0xffffffffu, inc.c, d);
        if (lane >= d) inc = add3(inc, uint3x{a, b, c});
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    uint3x base{0, 0, 0}, all{0, 0, 0};
    for (uint32_t k = 0; k < GLX_WIRE_BLOCK / 32; k++) {
        if (k < wid) base = add3(base, s_w[k]);
        all = add3(all, s_w[k]);
    }
    __syncthreads();
    *tot = all;
    return uint3x{base.a + inc.a - v.a, base.b + inc.b - v.b, base.c + inc.c - v.c};
}

// pass 1: per block of 1,024 records, the three sums
__global__ void __launch_bounds__(GLX_WIRE_BLOCK) glx_wire_sums_kernel(const DWire W, uint3x* __restrict__ block_sums) {
    const uint64_t r0 = ((uint64_t)blockIdx.x * GLX_WIRE_BLOCK + threadIdx.x) * GLX_WIRE_ITEMS;
    uint3x t{0, 0, 0};
    for (uint32_t i = 0; i < GLX_WIRE_ITEMS; i++)
        if (r0 + i < W.n_records) t = add3(t, wire_record_units(W, r0 + i));
    uint3x tot;
    block_scan3(t, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
// pass 2 (one block): exclusive scan of the block sums in place
__global__ void __launch_bounds__(GLX_WIRE_BLOCK) glx_wire_scan_sums_kernel(uint3x* __restrict__ block_sums, uint32_t n) {
    __shared__ uint3x s_carry;
    if (threadIdx.x == 0) s_carry = uint3x{0, 0, 0};
    __syncthreads();
    for (uint32_t base = 0; base < n; base += GLX_WIRE_BLOCK) {
        const uint32_t i = base + threadIdx.x;
        const uint3x v = i < n ? block_sums[i] : uint3x{0, 0, 0};
        uint3x tot;
        const uint3x ex = block_scan3(v, &tot);
        const uint3x carry = s_carry;
        if (i < n) block_sums[i] = add3(ex, carry);
        __syncthreads();
        if (threadIdx.x == 0) s_carry = add3(carry, tot);
        __syncthreads();
    }
}
// the same pair for the allele string offsets (one value)
__global__ void __launch_bounds__(GLX_WIRE_BLOCK) glx_wire_al_sums_kernel(const DWire W, uint3x* __restrict__ block_sums) {
    const uint64_t a0 = ((uint64_t)blockIdx.x * GLX_WIRE_BLOCK + threadIdx.x) * GLX_WIRE_ITEMS;
    uint3x t{0, 0, 0};
    for (uint32_t i = 0; i < GLX_WIRE_ITEMS; i++)
        if (a0 + i < W.n_al)
            t.a += wire_al_len(W, a0 + i);
    uint3x tot;
    block_scan3(t, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(GLX_WIRE_BLOCK) glx_wire_alleles_kernel(const DWire W, const uint3x* __restrict__ block_sums, uint32_t* __restrict__ al_len,
                                                                          uint32_t* __restrict__ al_str, uint64_t* __restrict__ al_prefix) {
    const uint64_t a0 = ((uint64_t)blockIdx.x * GLX_WIRE_BLOCK + threadIdx.x) * GLX_WIRE_ITEMS;
    uint32_t len[GLX_WIRE_ITEMS];
    uint3x t{0, 0, 0};
    for (uint32_t i = 0; i < GLX_WIRE_ITEMS; i++) {
        len[i] = 0;
        if (a0 + i < W.n_al)
            len[i] = wire_al_len(W, a0 + i);
        t.a += len[i];
    }
    uint3x tot;
    uint32_t off = block_scan3(t, &tot).a + block_sums[blockIdx.x].a;
    const char* pool = reinterpret_cast<const char*>(W.base + W.off[GLX_WS_AL_POOL]);
    for (uint32_t i = 0; i < GLX_WIRE_ITEMS; i++) {
        if (a0 + i >= W.n_al) break;
        al_len[a0 + i] = len[i];
        al_str[a0 + i] = off;
        uint64_t p = 0;
        for (uint32_t k = 0; k < len[i] && k < 8; k++) p |= (uint64_t)(uint8_t)pool[off + k] << (8 * k);
        al_prefix[a0 + i] = p;
        off += len[i];
    }
}

// pass 3: thread per record — every per-record column of DRecs, and the flags of the record's alleles
struct WireOut {
    int32_t *end, *maxend, *col_val, *pool;
    float* qual;
    uint32_t *filt, *gt, *allele_off;
    uint16_t* col_len;
    uint8_t* al_flags;
    const uint32_t *al_len, *al_str;
};
__device__ __forceinline__ int32_t wire_widen16(int16_t v) { return v == (int16_t)0x8000 ? GLX_INT_MISSING : (v == (int16_t)0x8001 ? GLX_INT_VECTOR_END : (int32_t)v); }
__global__ void __launch_bounds__(GLX_WIRE_BLOCK) glx_wire_records_kernel(const DWire W, const uint3x* __restrict__ block_sums, const WireOut O) {
    const uint64_t r0 = ((uint64_t)blockIdx.x * GLX_WIRE_BLOCK + threadIdx.x) * GLX_WIRE_ITEMS;
    uint3x u[GLX_WIRE_ITEMS], t{0, 0, 0};
    for (uint32_t i = 0; i < GLX_WIRE_ITEMS; i++) {
        u[i] = r0 + i < W.n_records ? wire_record_units(W, r0 + i) : uint3x{0, 0, 0};
        t = add3(t, u[i]);
    }
    uint3x tot;
    uint3x run = add3(block_scan3(t, &tot), block_sums[blockIdx.x]);
    const int32_t* beg = reinterpret_cast<const int32_t*>(W.base + W.off[GLX_WS_BEG]);
    const uint8_t* flags = W.sec_flags();
    const uint8_t* nal = W.sec_nal();
    const uint64_t R = W.n_records;
    for (uint32_t i = 0; i < GLX_WIRE_ITEMS; i++, run = add3(run, u[i - 1])) {
        const uint64_t r = r0 + i;
        if (r >= R) break;
        const uint32_t len = (W.flags & GLX_WF_LEN16) ? reinterpret_cast<const uint16_t*>(W.base + W.off[GLX_WS_LEN])[r]
                                                     : reinterpret_cast<const uint32_t*>(W.base + W.off[GLX_WS_LEN])[r];
        const int32_t e = beg[r] + (int32_t)len;
        O.end[r] = e;
        O.maxend[r] = e;  // the listed exceptions are patched afterwards
        if (W.flags & GLX_WF_GT8) {
            const uint8_t* g = W.sec_gt() + 2 * r;
            const uint32_t a = g[0] == 0xFF ? (uint32_t)(uint16_t)GLX_GT16_VECTOR_END : (uint32_t)g[0] << 1;
            const uint32_t b = g[1] == 0xFF ? (uint32_t)(uint16_t)GLX_GT16_VECTOR_END : (uint32_t)g[1] << 1;
            O.gt[r] = a | b << 16;
        } else
            O.gt[r] = reinterpret_cast<const uint32_t*>(W.sec_gt())[r];
        const bool var = !(flags[r] & GLX_RF_IS_REF);
        O.qual[r] = var ? reinterpret_cast<const float*>(W.base + W.off[GLX_WS_QUAL])[run.a] : 0.0f;
        O.filt[r] = (W.flags & GLX_WF_FILT_VARIANT) ? (var ? wire_filt(W, run.a) : 0u) : wire_filt(W, r);
        O.allele_off[r] = run.b;
        uint32_t po = run.c;
        for (uint32_t c = 0; c < W.n_cols; c++) {
            const uint32_t cl = wire_col_len(W, c, r);
            O.col_len[(size_t)c * R + r] = (uint16_t)cl;
            int32_t v = 0;
            if (cl == 1) {
                v = W.col_w[c] == 2 ? wire_widen16(reinterpret_cast<const int16_t*>(W.base + W.col_val_off[c])[r])
                                    : reinterpret_cast<const int32_t*>(W.base + W.col_val_off[c])[r];
            } else if (cl < 0xFFF0) {
                v = (int32_t)po;
                po += cl;
            }
            O.col_val[(size_t)c * R + r] = v;
        }
        if (var) {  // allele flags: is_dna / symbolic / deletion (pack_records; src/genotyper.cc:20-32,59-77, src/types.cc:26-43)
            const char* pool = reinterpret_cast<const char*>(W.base + W.off[GLX_WS_AL_POOL]);
            const uint32_t a0 = run.b, na = nal[r];
            const uint32_t ref_len = O.al_len[a0], ref_str = O.al_str[a0];
            for (uint32_t a = 0; a < na; a++) {
                const uint32_t al = O.al_len[a0 + a], as = O.al_str[a0 + a];
                bool dna = al > 0;
                for (uint32_t k = 0; k < al && dna; k++) {
                    const char ch = pool[as + k];
                    dna = ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T';
                }
                const bool sym = al > 1 && pool[as] == '<' && pool[as + al - 1] == '>';
                uint8_t f = 0;
                if (dna) f |= GLX_AF_IS_DNA;
                if (sym) f |= GLX_AF_SYMBOLIC;
                if (a > 0 && al < len && len == ref_len && al < ref_len) {
                    bool pre = true;
                    for (uint32_t k = 0; k < al && pre; k++) pre = pool[as + k] == pool[ref_str + k];
                    if (pre) f |= GLX_AF_DELETION;
                }
                O.al_flags[a0 + a] = f;
            }
        }
    }
}
__global__ void __launch_bounds__(256) glx_wire_pool_kernel(const DWire W, int32_t* __restrict__ pool) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (W.flags & GLX_WF_POOL16) {
        const int16_t* s = reinterpret_cast<const int16_t*>(W.base + W.off[GLX_WS_POOL]);
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < W.n_pool; i += stride) pool[i] = wire_widen16(s[i]);
    } else {
        const int32_t* s = reinterpret_cast<const int32_t*>(W.base + W.off[GLX_WS_POOL]);
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < W.n_pool; i += stride) pool[i] = s[i];
    }
}
__global__ void __launch_bounds__(256) glx_wire_maxend_kernel(const DWire W, int32_t* __restrict__ maxend) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.n_exc) return;
    const uint32_t* e = reinterpret_cast<const uint32_t*>(W.base + W.off[GLX_WS_MAXEND_EXC]);
    maxend[e[2 * i]] = (int32_t)e[2 * i + 1];
}

}  // namespace glxd
