// glx_dev_state.cuh — the device context and per-batch device state behind the opaque handles of include/glx.h
// (shared by glx_device.cu and glx_bcf.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/glx.h"
#include "glx_bgzf.cuh"
#include "glx_cell_fast.cuh"
#include "glx_dev_types.cuh"

using namespace glxd;

#define GLX_ARENA_CACHE 64  // a batch owns up to ten arenas and a range-batch driver keeps two or three batches in flight: no cudaFree in steady state
struct glx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;                      // worklist kernels of chunk j run here while the tiled kernel of chunk j+1 runs
    cudaEvent_t chunk_ev[GLX_MAX_CHUNKS + 2] = {};   // tiled(chunk j) done / aux drained
    uint32_t* host_ovf_pool = nullptr;               // pinned overflow words of the narrow download, same scheme as below
    unsigned long long* host_err_pool = nullptr;     // 64 pinned error words, one per live batch (allocating pinned memory
                                                     // per batch would synchronise the device and break pipelining)
    std::string err;
    int sm_count = 0;
    uint32_t* crc_pw = nullptr;  // device: x^(8 * 64 * k) mod P for the BGZF stage's CRC-32
    uint32_t func_attrs = 0;     // which large-shared-memory kernels have had their attribute set on this device
    uint8_t width_hint[GLX_MAX_FIELDS];  // narrow download: 1 when every value of the field in the previous batch fitted int8, else 2
    // kernel-path switches of the parity tests and ablations (glx_set_option); all 0 in production
    struct Options {
        int force_general = 0, force_dynamic_tiled = 0, no_resolve_sig = 0, no_fuse = 0, chunks = 1, profile_upload = 0;
    } opt;
    unsigned long long err_slots_used = 0;  // bit i: slot i of host_err_pool / host_ovf_pool belongs to a live batch
    // freed batch arenas are kept for the next batch of the same range stream (cudaMalloc/cudaFree of GB-sized
    // buffers costs milliseconds and would dominate a per-batch call)
    struct Cached {
        void* p = nullptr;
        size_t bytes = 0;
    } cache[GLX_ARENA_CACHE];
    // *bytes in: wanted; out: what the block really holds (hand that to give())
    cudaError_t take(void** p, size_t* bytes_io) {
        const size_t bytes = *bytes_io;
        int best = -1;
        for (int i = 0; i < GLX_ARENA_CACHE; i++)
            if (cache[i].p && cache[i].bytes >= bytes && (best < 0 || cache[i].bytes < cache[best].bytes)) best = i;
        if (best >= 0) {  // best fit, however loose: range batches of one stream differ in size (a shard's first and last
                          // batch are slices) and a cudaMalloc / cudaFree of a GB-sized arena stalls the whole device
            *p = cache[best].p;
            *bytes_io = cache[best].bytes;
            cache[best] = Cached();
            return cudaSuccess;
        }
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) {  // release the cache and retry once
            cudaGetLastError();
            for (int i = 0; i < GLX_ARENA_CACHE; i++)
                if (cache[i].p) {
                    cudaFree(cache[i].p);
                    cache[i] = Cached();
                }
            e = cudaMalloc(p, bytes);
        }
        return e;
    }
    void give(void* p, size_t bytes) {
        if (!p) return;
        int slot = -1;
        for (int i = 0; i < GLX_ARENA_CACHE; i++)
            if (!cache[i].p) slot = i;
        if (slot < 0) {  // evict the smallest
            slot = 0;
            for (int i = 1; i < GLX_ARENA_CACHE; i++)
                if (cache[i].bytes < cache[slot].bytes) slot = i;
            if (cache[slot].bytes >= bytes) {
                cudaFree(p);
                return;
            }
            cudaFree(cache[slot].p);
        }
        cache[slot].p = p;
        cache[slot].bytes = bytes;
    }
};

namespace glx_state {

struct Arena {
    char* base = nullptr;
    size_t size = 0, used = 0;
    template <class T>
    T* take(size_t n) {
        used = (used + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + used);
        used += n * sizeof(T);
        return p;
    }
};

inline int cuda_fail(glx_ctx* ctx, cudaError_t e, const char* what) {
    ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
    return GLX_E_CUDA;
}
#define CU(call)                                                  \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

}  // namespace glx_state
using namespace glx_state;

struct glx_dev_batch;
int glx_take_err_slot(glx_ctx* ctx, glx_dev_batch* d);

// N1: device state of the BCF2 record encoder of one batch (glx_bcf.cu)
struct GlxBcfState {
    bool attached = false;
    char* meta_base = nullptr;   // per-site strings / header ids (glx_bcf_meta) + allele_af
    size_t meta_bytes = 0;
    char* work_base = nullptr;   // min/max table, site plans, offsets, header pool
    size_t work_bytes = 0;
    uint8_t* stream = nullptr;   // raw record stream, or the BGZF slots
    size_t stream_bytes = 0;
    uint8_t* packed = nullptr;   // BGZF mode: the compressed blocks, contiguous
    size_t packed_bytes = 0;
    unsigned long long* totals = nullptr;  // device: {stream bytes, header-pool bytes, BGZF bytes, BGZF blocks}
    uint32_t* blk_sizes = nullptr;           // BGZF: size of each span's block, then
    unsigned long long* blk_offsets = nullptr;  // its offset in the contiguous output
    uint32_t* span_site = nullptr;           // first site of every span
    uint32_t* hist = nullptr;                // BGZF: symbol counts of the sampled spans
    struct glxd::BgzfTable* table = nullptr; // BGZF: the batch's Huffman code + block header
    uint64_t stream_bound = 0, hdr_bound = 0;
    BcfEnc E;
    cudaEvent_t len_ev = nullptr;  // recorded when the lengths have landed in the batch's pinned words
    cudaEvent_t meta_ev = nullptr; // recorded behind glx_bcf_attach's copies (which may be on a stream of their own)
    int mode = -1;                 // last glx_encode_bcf_on: 0 raw records, 1 BGZF blocks
};

struct glx_dev_batch {
    GlxBcfState bcf;
    void* wire_sums = nullptr;     // wire upload: block sums of the record scan / of the allele-length scan (inside the input arena)
    void* wire_al_sums = nullptr;
    void* wire_shape_scratch = nullptr;  // GLX_WF_SHAPES: expanded GT pairs [2R] + column lengths [C][R] (inside the input arena)
    DCfg cfg;
    DRecs R;
    DSites S;
    DOut O;
    char* in_base = nullptr;
    char* out_base = nullptr;
    size_t in_bytes = 0, out_bytes = 0, in_alloc = 0, out_alloc = 0;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // around resolve / tiled / single / general
    bool profile = false;
    uint32_t n_sites = 0, n_samples = 0, n_fields = 0;
    uint64_t n_cells = 0;
    uint64_t fld_total[GLX_MAX_FIELDS] = {};
    uint32_t max_slots = 0;
    uint8_t* cnt_scratch = nullptr;
    size_t cnt_scratch_bytes = 0;
    int general_blocks = 0;
    int launches_per_call = 0;
    bool force_general = false;
    uint32_t n_chunks = 1;
    uint32_t max_site_alleles = 0;
    uint32_t n_cols = 0;  // source columns of the record store
    uint32_t tile_sites = GLX_TILE_SITES;  // sites per tile of the tiled kernel in use
    bool fused_variant = false;  // the tiled kernel finishes its own lone-variant cells
    unsigned long long* host_err = nullptr;  // pinned copy of the device error word (asynchronous download)
    int err_slot = -1;                       // slot of the context's pinned pools, or -1: host_err / host_overflow are the batch's own allocation
    cudaStream_t last_stream = nullptr;      // stream of the last glx_launch (glx_check waits on it)
    cudaStream_t dl_stream = nullptr;        // stream of the last download
    cudaEvent_t up_ev[4] = {nullptr, nullptr, nullptr, nullptr};  // around the two upload-side kernels (option profile_upload)
    int16_t* narrow = nullptr;               // device int16 images of the integer fields + overflow words
    size_t narrow_bytes = 0;
    uint8_t narrow_width[GLX_MAX_FIELDS] = {};  // width each field was sent in by the last narrow download
    uint32_t* narrow_overflow = nullptr;     // [GLX_MAX_FIELDS] device
    uint32_t* host_overflow = nullptr;       // pinned
    FastPlan plan;
    uint2* worklist = nullptr;     // [n_cells] (cell id, record cursor) for the general kernel
    uint32_t* n_work = nullptr;
    uint2* decl_list = nullptr;  // declined list of the fused tiled kernel (FastPlan::decl_list), inside the worklist arena
    uint32_t* decl_n = nullptr;  // [GLX_MAX_CHUNKS]
    uint32_t decl_cap = 0;
    size_t worklist_bytes = 0;
    uint32_t n_sample_blocks = 0, n_site_tiles = 0;
    uint32_t samples_per_cta = GLX_TILE_THREADS;  // one sample per lane
    typedef void (*resolve_fn)(const DCfg, const FastPlan, const DRecs, uint32_t, int32_t*, const int32_t*, uint32_t, uint32_t*, const uint2*);
    typedef void (*tiled_fn)(const DCfg, const FastPlan, const DRecs, const DSites, const DOut, const int32_t*, const uint32_t*, uint2*, uint32_t*, uint32_t,
                             uint32_t);
    resolve_fn resolve_sig = nullptr;  // signature-specialised resolve kernel (plans with record views), else the generic one
    tiled_fn tiled_kernel = nullptr;  // signature-specialised tiled kernel, or the dynamic one
    size_t tiled_smem = 0;
    int32_t* blobs = nullptr;  // [n_records][blob_words] resolved records, then the tile cursor table and tile_q
    size_t blob_bytes = 0;
    uint32_t* tile_cursors = nullptr;  // [n_site_tiles][n_samples]
    int32_t* tile_q = nullptr;         // [n_site_tiles] suffix-min of each tile's first qbeg
    uint2* block_hints = nullptr;      // [resolve blocks + 1] (sample, t_hi) of each block's first record (glx_block_hints_kernel)
    uint32_t n_datasets = 0;
    // multi-sample (or reordered) datasets: every cell goes through glx_cells_multi_kernel
    bool multi = false;
    const int32_t* gt_pool = nullptr;   // device copy of glx_records.gt_pool
    const uint2* out_map = nullptr;     // device [n_samples]: x = dataset (0xFFFFFFFF: the sample has none), y = sample within it | samples of it << 16
    std::vector<uint2> out_map_host;    // source of that copy (lives as long as the batch)
};

