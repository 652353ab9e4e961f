// glx_bgzf.cuh — BGZF blocks on the device: DEFLATE (RFC 1951, dynamic Huffman + LZ77 restricted to two structural
// distances) and CRC-32 of one <= 65,280-byte span of the BCF record stream, framed as a BGZF block (SAM spec §4.1).
//
// The reference writes its BCF through htslib's bgzf (single writer thread, zlib level -1; src/service.cc:254-275). A
// BGZF reader only requires each block to be a valid gzip member that inflates to the same bytes, so this encoder does
// not imitate zlib's bit stream; it is built for the data it sees. A BCF record is field-major: per FORMAT field one
// typed vector for all samples. Two redundancies dominate: (1) runs of one byte (GT of hom-ref cells, RNC "..",
// zero high bytes), caught by distance-1 matches; (2) a sample's value at this site equals its value at the previous
// site whenever the same reference band covers both, caught by matches at the distance D between the same FORMAT block of
// consecutive records (known from the layout, no hashing). Everything else is literals under a Huffman code.
//
// The Huffman code is built ONCE PER BATCH (blocks of one batch hold the same mix of fields): a sampling kernel tokenizes a
// few hundred spans into a histogram, one small block turns it into code lengths, canonical codes and the finished
// dynamic-block header bits, and every BGZF block then only tokenizes, counts, scans and emits. Symbols the sample never
// saw still get a code (every frequency is smoothed by one).
//
// Every phase is a block-level function over GLX_BGZF_WORKERS virtual workers (64-byte chunks of the span), written so
// that the host build (tests/emu) runs it with one thread and checks the result with zlib.
#pragma once
#include <cstdint>
#include <cstring>

#include "glx_bcf.cuh"

namespace glxd {

#define GLX_BGZF_WORKERS 1024u
#define GLX_BGZF_CHUNK 64u    // bytes per worker: 1024 x 64 = 65,536 >= GLX_BCF_SPAN
#define GLX_BGZF_SLOT 65536u  // bytes reserved per block before compaction
#define GLX_BGZF_LEAD 14u     // the block starts at byte 14 of its staging buffer, so that the deflate data is word-aligned (14 + 18)
#define GLX_BGZF_OUT (GLX_BGZF_LEAD + 18u + 5u + GLX_BCF_SPAN + 8u + 16u)
#define GLX_BGZF_IN_PAD (GLX_BCF_SPAN + (GLX_BCF_SPAN / 64u + 2u) * 4u)  // the span with 4 bytes of padding after every 64 (BcfPadded)
#define GLX_BGZF_LITS 286u
#define GLX_BGZF_DISTS 30u
#define GLX_BGZF_HIST (288u + 32u)
#define GLX_BGZF_HDR_WORDS 96u
#define GLX_BGZF_SAMPLES 384u  // spans tokenized for the batch's histogram

// the batch's Huffman code and the finished header of every dynamic block
struct BgzfTable {
    uint8_t lit_len[288], dist_len[32];
    uint16_t lit_code[288], dist_code[32];  // bit-reversed canonical codes
    uint32_t lit_cl[288], dist_cl[32];      // code | length << 16: one load per symbol when emitting
    uint32_t hdr_bits;
    uint32_t hdr_words[GLX_BGZF_HDR_WORDS];
};

struct BgzfShared {
    BgzfTable tab;
    uint32_t freq[GLX_BGZF_HIST];               // sampling: literal/length counts [0, 288), distance counts [288, 320)
    uint32_t wbits[GLX_BGZF_WORKERS];           // per-worker bit counts -> bit offsets
    unsigned long long tok[GLX_BGZF_WORKERS];   // bit i: a token starts at byte i of the worker's chunk (literal = 1 byte, match = to the next start)
    unsigned long long useD[GLX_BGZF_WORKERS];  // bit i: the match starting there copies from the structural distance D, not from distance 1
    uint32_t n_in, data_bits, crc, crc_acc, crc_tail, stored, block_len;
    uint32_t crc_table[4][256];
};

__host__ __device__ __forceinline__ int bgzf_log2(uint32_t x) {  // floor(log2 x), x > 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
__host__ __device__ __forceinline__ uint32_t bgzf_or(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicOr(p, v);
#else
    const uint32_t o = *p;
    *p |= v;
    return o;
#endif
}
__host__ __device__ __forceinline__ void bgzf_add(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}

// ---- CRC-32 (IEEE, reflected) ----
#define GLX_CRC_POLY 0xedb88320u
__host__ __device__ inline uint32_t crc_multmodp(uint32_t a, uint32_t b) {  // a(x) * b(x) mod P in the reflected representation
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = b & 1 ? (b >> 1) ^ GLX_CRC_POLY : b >> 1;
    }
    return p;
}
__host__ __device__ inline uint32_t crc_x_pow_8n(uint32_t n) {  // x^(8n) mod P
    uint32_t p = 1u << 31;   // x^0
    uint32_t sq = 1u << 23;  // x^8
    while (n) {
        if (n & 1) p = crc_multmodp(sq, p);
        sq = crc_multmodp(sq, sq);
        n >>= 1;
    }
    return p;
}
// CRC of the span from the workers' chunk CRCs without a combine tree: with the data right-aligned in the 1024 x 64 byte
// frame (leading zeros leave a raw, init-0 register untouched), crc_raw(span) = XOR_w crc_raw(chunk w) * x^(8 * 64 * (1023 - w)).
// The 1024 powers are computed once on the host (bgzf_crc_powers) and live in global memory, followed by, per power, the
// byte-sliced table of "multiply by this power" (the product is GF(2)-linear in the other factor): four loads and three XORs
// per chunk instead of a 32-step shift-and-add (6 % of the BGZF kernel's instructions). 4 MB, L2-resident.
#define GLX_BGZF_PW_WORDS 1028u  // the powers, padded
#define GLX_BGZF_CRC_WORDS (GLX_BGZF_PW_WORDS + GLX_BGZF_WORKERS * 1024u)
inline void bgzf_crc_powers(uint32_t* pw /* [GLX_BGZF_CRC_WORDS]; pw[GLX_BGZF_WORKERS] = x^(8 * GLX_BCF_SPAN) */) {
    uint32_t p = 1u << 31;
    const uint32_t step = crc_x_pow_8n(GLX_BGZF_CHUNK);
    for (uint32_t k = 0; k < GLX_BGZF_WORKERS; k++) {
        pw[k] = p;
        p = crc_multmodp(step, p);
    }
    pw[GLX_BGZF_WORKERS] = crc_x_pow_8n(GLX_BCF_SPAN);
    for (uint32_t k = GLX_BGZF_WORKERS + 1; k < GLX_BGZF_PW_WORDS; k++) pw[k] = 0;
    uint32_t* T = pw + GLX_BGZF_PW_WORDS;
    for (uint32_t k = 0; k < GLX_BGZF_WORKERS; k++) {
        uint32_t bit[32];
        for (uint32_t i = 0; i < 32; i++) bit[i] = crc_multmodp(pw[k], 1u << i);
        for (uint32_t j = 0; j < 4; j++)
            for (uint32_t v = 0; v < 256; v++) {
                uint32_t x = 0;
                for (uint32_t b = 0; b < 8; b++)
                    if (v >> b & 1u) x ^= bit[8 * j + b];
                T[(k * 4 + j) * 256 + v] = x;
            }
    }
}
// c(x) * x^(8 * 64 * k) mod P through the tables of bgzf_crc_powers
__host__ __device__ __forceinline__ uint32_t bgzf_crc_shift(const uint32_t* __restrict__ crc_pw, uint32_t k, uint32_t c) {
    const uint32_t* M = crc_pw + GLX_BGZF_PW_WORDS + (size_t)k * 1024u;
#if defined(__CUDA_ARCH__)
    return __ldg(M + (c & 0xffu)) ^ __ldg(M + 256u + (c >> 8 & 0xffu)) ^ __ldg(M + 512u + (c >> 16 & 0xffu)) ^ __ldg(M + 768u + (c >> 24));
#else
    return M[c & 0xffu] ^ M[256u + (c >> 8 & 0xffu)] ^ M[512u + (c >> 16 & 0xffu)] ^ M[768u + (c >> 24)];
#endif
}
__host__ __device__ __forceinline__ void bgzf_xor(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicXor(p, v);
#else
    *p ^= v;
#endif
}
__host__ __device__ inline void bgzf_crc_setup(BgzfShared& sh, uint32_t tid, uint32_t nt) {
    for (uint32_t i = tid; i < 256; i += nt) {  // slice-by-4 tables: crc_table[k][i] = CRC of byte i followed by k zero bytes
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = c & 1 ? (c >> 1) ^ GLX_CRC_POLY : c >> 1;
        sh.crc_table[0][i] = c;
        for (int t = 1; t < 4; t++) {
            for (int k = 0; k < 8; k++) c = c & 1 ? (c >> 1) ^ GLX_CRC_POLY : c >> 1;
            sh.crc_table[t][i] = c;
        }
    }
    if (tid == 0) sh.crc_acc = sh.crc_tail = 0;
}
// The span's CRC from the workers' chunk CRCs (computed inside the tokenizer's word loop): raw (init 0, no final xor) CRCs
// combine as crc(A || B) = crc(A) * x^(8|B|) + crc(B); with 64-byte chunks, chunk w of n_full full chunks is followed by
// 64 * (n_full - 1 - w) + (n mod 64) bytes, so crc_raw(span) = (XOR_w crc_w * pw[n_full - 1 - w]) * x^(8 (n mod 64)) + crc(tail).
__host__ __device__ inline void bgzf_crc_finish(BgzfShared& sh, const uint32_t* __restrict__ pw) {  // one thread
    const uint32_t n = sh.n_in, r = n % GLX_BGZF_CHUNK;
    const uint32_t raw = (r ? crc_multmodp(crc_x_pow_8n(r), sh.crc_acc) : sh.crc_acc) ^ sh.crc_tail;
    const uint32_t xn = n == GLX_BCF_SPAN ? pw[GLX_BGZF_WORKERS] : crc_x_pow_8n(n);
    sh.crc = ~(raw ^ crc_multmodp(xn, 0xffffffffu));
}

// ---- DEFLATE symbols ----
__host__ __device__ __forceinline__ void bgzf_len_sym(uint32_t len, uint32_t& sym, uint32_t& ebits, uint32_t& eval) {  // 3 <= len <= 258
    if (len == 258) {
        sym = 285;
        ebits = eval = 0;
        return;
    }
    const uint32_t l = len - 3;
    if (l < 8) {
        sym = 257 + l;
        ebits = eval = 0;
        return;
    }
    const uint32_t e = (uint32_t)bgzf_log2(l) - 2;
    sym = 257 + 4 * e + (l >> e);
    ebits = e;
    eval = l & ((1u << e) - 1);
}
__host__ __device__ __forceinline__ void bgzf_dist_sym(uint32_t dist, uint32_t& sym, uint32_t& ebits, uint32_t& eval) {  // 1 <= dist <= 32768
    const uint32_t x = dist - 1;
    if (x < 4) {
        sym = x;
        ebits = eval = 0;
        return;
    }
    const uint32_t e = (uint32_t)bgzf_log2(x) - 1;
    sym = 2 * e + 2 + ((x >> e) & 1u);
    ebits = e;
    eval = x & ((1u << e) - 1);
}

// ---- tokenizer ----
__host__ __device__ __forceinline__ uint32_t bgzf_eq_nibble(uint32_t x, uint32_t y) {  // bit b: byte b of x == byte b of y
    const uint32_t d = x ^ y;
    const uint32_t z = ~(((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d) & 0x80808080u;  // 0x80 in every byte of d that is zero
    return (z * 0x00204081u) >> 28;
}
__host__ __device__ __forceinline__ uint32_t bgzf_run(unsigned long long m, uint32_t i) {  // ones of m from bit i up
    const unsigned long long z = ~(m >> i);
    if (z == 0) return 64 - i;
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffsll((long long)z) - 1;
#else
    return (uint32_t)__builtin_ffsll((long long)z) - 1;
#endif
}
__host__ __device__ __forceinline__ uint32_t bgzf_lowbit(unsigned long long m) {  // index of the lowest set bit, m != 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffsll((long long)m) - 1;
#else
    return (uint32_t)__builtin_ffsll((long long)m) - 1;
#endif
}
__host__ __device__ __forceinline__ uint32_t bgzf_popcll(unsigned long long m) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popcll(m);
#else
    return (uint32_t)__builtin_popcountll(m);
#endif
}
// What a worker needs to know about the pieces under its 64-byte chunk: where another piece starts inside it (a D-match
// must not run across: the distance changes) and the structural distance of the first four pieces (more: no D-matches there).
struct BgzfChunkCtx {
    unsigned long long brk;  // bit i (i >= 1): a piece starts at chunk byte i
    uint32_t D[4];
    uint32_t a[4], e[4];  // span-relative byte range of those pieces inside the chunk
    uint32_t np;
};
__host__ __device__ inline void bgzf_chunk_ctx(const BcfSpanTab& T, uint32_t c0, uint32_t L, BgzfChunkCtx& C) {
    C.brk = 0;
    C.np = 0;
    for (int i = 0; i < 4; i++) C.D[i] = C.a[i] = C.e[i] = 0;
    if (!T.complete || !T.n) return;
    for (uint32_t pc = bcf_tab_find(T, (long long)c0); pc < T.n && T.start[pc] < (long long)(c0 + L); pc++) {
        const long long ps = T.start[pc], pe = T.start[pc + 1];
        if (pe <= ps) continue;
        if (ps > (long long)c0) C.brk |= 1ull << (ps - c0);
        if (C.np < 4) {
            const uint32_t D = T.dist[pc];
            C.D[C.np] = (D > 1 && D <= 32768) ? D : 0u;
            C.a[C.np] = ps > (long long)c0 ? (uint32_t)ps : c0;
            C.e[C.np] = pe < (long long)(c0 + L) ? (uint32_t)pe : c0 + L;
        }
        C.np++;
    }
}
__host__ __device__ __forceinline__ uint32_t bgzf_D_at(const BgzfChunkCtx& C, uint32_t p) {  // distance of the piece holding chunk byte p
    const uint32_t ord = bgzf_popcll(C.brk & ((2ull << p) - 1ull));
    return ord < 4 ? C.D[ord] : 0u;
}

__host__ __device__ __forceinline__ uint32_t bgzf_low32(uint32_t m) {  // index of the lowest set bit, m != 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)m) - 1;
#else
    return (uint32_t)__builtin_ffs((int)m) - 1;
#endif
}
// Worker w owns bytes [64w, 64w+64) of the span (in = the BcfPadded staging buffer). It finds, per byte, whether the byte
// repeats the one before it (distance 1) and whether it repeats the byte at the structural distance D of its piece, as two
// 64-bit masks (word-wise compares), then walks them greedily — a D-match of >= 4 bytes that is longer than the distance-1
// run, else a distance-1 run of >= 3, else a literal — recording the token starts. Matches end at the chunk's end and at
// piece boundaries. HIST: count the symbols into sh.freq (sampling). COUNT: the worker's bit count under sh.tab into sh.wbits.
template <bool HIST, bool COUNT>
__host__ __device__ inline void bgzf_tokenize(BgzfShared& sh, const BcfSpanTab& T, const uint8_t* in, uint32_t n, const uint32_t* __restrict__ crc_pw, uint32_t tid,
                                              uint32_t nt) {
    for (uint32_t w = tid; w < GLX_BGZF_WORKERS; w += nt) {
        const uint32_t c0 = w * GLX_BGZF_CHUNK;
        if (c0 >= n) {
            sh.tok[w] = sh.useD[w] = 0;
            if (COUNT) sh.wbits[w] = 0;
            continue;
        }
        const uint32_t L = n - c0 < GLX_BGZF_CHUNK ? n - c0 : GLX_BGZF_CHUNK;
        const unsigned long long live = L == 64 ? ~0ull : ((1ull << L) - 1);
        const uint32_t* W = reinterpret_cast<const uint32_t*>(in + BcfPadded::phys(c0));  // the chunk is contiguous and word-aligned
        unsigned long long eq1 = 0, zero = 0;
        uint32_t carry = c0 ? in[BcfPadded::phys(c0 - 1)] : 0u, crc = 0;
        for (uint32_t k = 0; k < 16; k++) {
            const uint32_t x = W[k];
            eq1 |= (unsigned long long)bgzf_eq_nibble(x, (x << 8) | carry) << (4 * k);
            if (HIST) zero |= (unsigned long long)bgzf_eq_nibble(x, 0u) << (4 * k);
            carry = x >> 24;
            if (COUNT && L == 64) {  // CRC-32 of the chunk, four bytes per step (the compressing pass only)
                const uint32_t y = crc ^ x;
                crc = sh.crc_table[3][y & 0xff] ^ sh.crc_table[2][y >> 8 & 0xff] ^ sh.crc_table[1][y >> 16 & 0xff] ^ sh.crc_table[0][y >> 24];
            }
        }
        if (COUNT) {
            if (L == 64) {
                if (crc) bgzf_xor(&sh.crc_acc, bgzf_crc_shift(crc_pw, n / GLX_BGZF_CHUNK - 1 - w, crc));
            } else {  // the span's short last chunk
                for (uint32_t i = 0; i < L; i++) crc = sh.crc_table[0][(crc ^ in[BcfPadded::phys(c0 + i)]) & 0xff] ^ (crc >> 8);
                sh.crc_tail = crc;
            }
        }
        if (!c0) eq1 &= ~1ull;  // the span's first byte has no predecessor
        eq1 &= live;
        BgzfChunkCtx C;
        bgzf_chunk_ctx(T, c0, L, C);
        unsigned long long eqD = 0;
        for (uint32_t pi = 0; pi < (C.np < 4 ? C.np : 4u); pi++) {
            const uint32_t D = C.D[pi];
            if (!D) continue;
            const uint32_t a = C.a[pi] < D ? D : C.a[pi], e = C.e[pi];  // positions below D have no source inside the span
            if (a >= e) continue;
            unsigned long long m = 0;
            for (uint32_t k = (a - c0) >> 2; k <= (e - 1 - c0) >> 2; k++) {
                const uint32_t i0 = c0 + 4 * k;
                uint32_t ref;
                if (i0 >= D) {
                    const uint32_t r0 = i0 - D, al = r0 & ~3u, shf = (r0 & 3u) * 8;
                    const uint32_t lo = *reinterpret_cast<const uint32_t*>(in + BcfPadded::phys(al));
                    const uint32_t hi = *reinterpret_cast<const uint32_t*>(in + BcfPadded::phys(al + 4));
                    ref = shf ? (lo >> shf) | (hi << (32 - shf)) : lo;
                } else {  // the word straddles position D: its first bytes have no source; they are masked below
                    ref = 0;
                    for (uint32_t bb = 0; bb < 4; bb++)
                        if (i0 + bb >= D) ref |= (uint32_t)in[BcfPadded::phys(i0 + bb - D)] << (8 * bb);
                }
                m |= (unsigned long long)bgzf_eq_nibble(W[k], ref) << (4 * k);
            }
            const uint32_t ra = a - c0, re = e - c0;  // keep positions [ra, re)
            const unsigned long long keep = (re >= 64 ? ~0ull : ((1ull << re) - 1)) & ~((1ull << ra) - 1);
            eqD |= m & keep;
        }
        // greedy walk, jumping from candidate to candidate: a distance-1 run of >= 3 or a D-run of >= 4 can only start where
        // the next 3 (4) mask bits are set; everything in between is literals
        const unsigned long long contD = eqD & ~C.brk;  // a D-run continues into byte j iff eqD[j] and no piece starts there
        const unsigned long long cand1 = eq1 & (eq1 >> 1) & (eq1 >> 2);
        const unsigned long long candD = eqD & (contD >> 1) & (contD >> 2) & (contD >> 3);
        const unsigned long long cand = cand1 | candD;
        unsigned long long tok = 0, useD = 0, lits = 0;
        uint32_t i = 0;
        while (i < L) {
            const unsigned long long c = cand >> i;
            if (!c) {  // literals to the end
                lits |= live & ~((1ull << i) - 1);
                break;
            }
            const uint32_t j = i + bgzf_lowbit(c);
            if (j > i) lits |= ((1ull << j) - 1) & ~((1ull << i) - 1);
            const uint32_t r1 = bgzf_run(eq1, j);
            const uint32_t rD = (eqD >> j & 1ull) ? 1 + (j < 63 ? bgzf_run(contD, j + 1) : 0u) : 0u;
            tok |= 1ull << j;
            if (rD >= 4 && rD > r1) {
                useD |= 1ull << j;
                i = j + rD;
            } else {
                i = j + r1;  // a candidate that is not a D-match is a distance-1 run of >= 3
            }
        }
        tok |= lits;
        sh.tok[w] = tok;
        sh.useD[w] = useD;
        if (HIST) {
            // zero is by far the most common literal of typed integer vectors: one add for all of the worker's
            const uint32_t nz = bgzf_popcll(lits & zero);
            if (nz) bgzf_add(&sh.freq[0], nz);
            for (unsigned long long m = lits & ~zero; m; m &= m - 1) {
                const uint32_t p = bgzf_lowbit(m);
                bgzf_add(&sh.freq[(W[p >> 2] >> (8 * (p & 3))) & 0xffu], 1);
            }
        }
        if (HIST || COUNT) {
            uint32_t bits = 0;
            if (COUNT) {
                const uint32_t lo = (uint32_t)lits, hi = (uint32_t)(lits >> 32);
                for (uint32_t k = 0; k < 16; k++) {
                    const uint32_t nib = ((k < 8 ? lo : hi) >> (4 * (k & 7))) & 15u;
                    if (!nib) continue;
                    const uint32_t x = W[k];
                    if (nib & 1u) bits += sh.tab.lit_len[x & 0xffu];
                    if (nib & 2u) bits += sh.tab.lit_len[x >> 8 & 0xffu];
                    if (nib & 4u) bits += sh.tab.lit_len[x >> 16 & 0xffu];
                    if (nib & 8u) bits += sh.tab.lit_len[x >> 24];
                }
            }
            for (unsigned long long t = tok & ~lits; t; t &= t - 1) {  // matches
                const uint32_t p = bgzf_lowbit(t);
                const unsigned long long rest = tok >> p >> 1;
                const uint32_t q = rest ? p + 1 + bgzf_lowbit(rest) : L;
                uint32_t sym, eb, ev, dsym, deb;
                bgzf_len_sym(q - p, sym, eb, ev);
                bgzf_dist_sym((useD >> p & 1ull) ? bgzf_D_at(C, p) : 1u, dsym, deb, ev);
                if (HIST) {
                    bgzf_add(&sh.freq[sym], 1);
                    bgzf_add(&sh.freq[288 + dsym], 1);
                }
                if (COUNT) bits += sh.tab.lit_len[sym] + eb + sh.tab.dist_len[dsym] + deb;
            }
            if (COUNT) sh.wbits[w] = bits;
        }
    }
}

// ---- Huffman code of the batch (one thread; runs once per batch). Code lengths <= maxbits forming a COMPLETE prefix code
// (zlib's inflate rejects incomplete literal/length sets): start from ceil(log2(total / f)) per symbol, repair the Kraft sum
// if the clamp at maxbits oversubscribed it, then hand the slack out by shortening symbols until the sum is exactly one. ----
__host__ __device__ inline void bgzf_code_lengths(const uint32_t* freq, uint32_t n, uint32_t maxbits, uint8_t* len) {
    uint32_t used = 0, last = 0;
    unsigned long long total = 0;
    for (uint32_t i = 0; i < n; i++) {
        len[i] = 0;
        if (freq[i]) {
            used++;
            last = i;
            total += freq[i];
        }
    }
    if (used == 0) return;
    if (used == 1) {
        len[last] = 1;
        return;
    }
    const uint32_t full = 1u << maxbits;
    unsigned long long K = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (!freq[i]) continue;
        uint32_t l = 1;
        while (l < maxbits && ((unsigned long long)freq[i] << l) < total) l++;
        len[i] = (uint8_t)l;
        K += full >> l;
    }
    for (uint32_t L = maxbits - 1; K > full && L >= 1; L--)  // oversubscribed: lengthen the longest codes first
        for (uint32_t i = 0; i < n && K > full; i++)
            if (len[i] == L) {
                len[i]++;
                K -= full >> (L + 1);
            }
    unsigned long long slack = full - K;
    while (slack) {  // each pass shortens at least one code of the current maximum length
        for (uint32_t i = 0; i < n && slack; i++) {
            const uint32_t l = len[i];
            if (l > 1 && (full >> l) <= slack) {
                len[i] = (uint8_t)(l - 1);
                slack -= full >> l;
            }
        }
    }
}
// canonical codes, bit-reversed for LSB-first packing
__host__ __device__ inline void bgzf_assign_codes(const uint8_t* len, uint32_t n, uint16_t* code) {
    uint32_t bl_count[16] = {0}, next[16];
    for (uint32_t i = 0; i < n; i++) bl_count[len[i]]++;
    bl_count[0] = 0;
    uint32_t c = 0;
    for (uint32_t b = 1; b < 16; b++) {
        c = (c + bl_count[b - 1]) << 1;
        next[b] = c;
    }
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t l = len[i];
        if (!l) {
            code[i] = 0;
            continue;
        }
        uint32_t v = next[l]++, r = 0;
        for (uint32_t b = 0; b < l; b++) {
            r = r << 1 | (v & 1);
            v >>= 1;
        }
        code[i] = (uint16_t)r;
    }
}

// serial bit writer over a zeroed, word-aligned buffer
struct BgzfBits {
    uint32_t* words;
    uint32_t pos;  // in bits
    __host__ __device__ void put(uint32_t v, uint32_t nbits) {
        if (!nbits) return;
        const uint32_t w = pos >> 5, o = pos & 31;
        words[w] |= v << o;
        if (o + nbits > 32) words[w + 1] |= v >> (32 - o);
        pos += nbits;
    }
};

// freq[320] (literal/length, then distance counts of the sampled spans) -> the batch's table: smoothed frequencies, both
// codes, and the dynamic-block header (BFINAL = 1, BTYPE = 10, the run-length coded code lengths). A block-level function
// (every thread calls it): per-symbol work runs in parallel, the two Kraft settles and the header walk on single threads.
// scratch: 64 words of shared memory.
__host__ __device__ inline void bgzf_table_build(const uint32_t* freq, BgzfTable& tab, uint32_t* scratch, uint32_t tid, uint32_t nt) {
    uint32_t* tot = scratch;       // [0] literal/length total, [1] distance total, [2] / [3] their Kraft sums (units of 2^-15)
    uint32_t* cnt = scratch + 8;   // [0..16) literal/length code-length counts -> first codes, [16..32) distance
    for (uint32_t i = tid; i < 48; i += nt) scratch[i] = 0;
    GLX_BLOCK_SYNC();
    for (uint32_t i = tid; i < GLX_BGZF_HIST; i += nt) {  // + 1: every symbol can be coded
        const bool lit = i < 288;
        const uint32_t j = lit ? i : i - 288;
        if (lit ? j < GLX_BGZF_LITS : j < GLX_BGZF_DISTS) bgzf_add(&tot[lit ? 0 : 1], freq[i] + 1);
    }
    GLX_BLOCK_SYNC();
    for (uint32_t i = tid; i < GLX_BGZF_HIST; i += nt) {
        const bool lit = i < 288;
        const uint32_t j = lit ? i : i - 288;
        uint32_t l = 0;
        if (lit ? j < GLX_BGZF_LITS : j < GLX_BGZF_DISTS) {
            const uint32_t f = freq[i] + 1, total = tot[lit ? 0 : 1];
            l = 1;
            while (l < 15 && ((unsigned long long)f << l) < total) l++;
            bgzf_add(&tot[lit ? 2 : 3], 32768u >> l);
        }
        (lit ? tab.lit_len : tab.dist_len)[j] = (uint8_t)l;
    }
    GLX_BLOCK_SYNC();
    // settle each code to a Kraft sum of exactly one: lengthen the longest codes while oversubscribed, then shorten codes in
    // index order while slack remains (each pass shortens at least one code of the current maximum length)
    for (uint32_t a = 0; a < 2; a++) {
        if (tid != (nt > 32 ? a * 32 : 0u)) continue;
        uint8_t* len = a ? tab.dist_len : tab.lit_len;
        const uint32_t n = a ? GLX_BGZF_DISTS : GLX_BGZF_LITS, full = 32768u;
        uint32_t K = tot[2 + a];
        for (uint32_t L = 14; K > full && L >= 1; L--)
            for (uint32_t i = 0; i < n && K > full; i++)
                if (len[i] == L) {
                    len[i]++;
                    K -= full >> (L + 1);
                }
        uint32_t slack = full - K;
        while (slack)
            for (uint32_t i = 0; i < n && slack; i++) {
                const uint32_t l = len[i];
                if (l > 1 && (full >> l) <= slack) {
                    len[i] = (uint8_t)(l - 1);
                    slack -= full >> l;
                }
            }
    }
    GLX_BLOCK_SYNC();
    for (uint32_t i = tid; i < GLX_BGZF_HIST; i += nt) {  // canonical codes: count per length ...
        const bool lit = i < 288;
        const uint32_t l = lit ? tab.lit_len[i] : tab.dist_len[i - 288];
        if (l) bgzf_add(&cnt[(lit ? 0 : 16) + l], 1);
    }
    GLX_BLOCK_SYNC();
    if (tid == 0)  // ... first code per length ...
        for (uint32_t a = 0; a < 2; a++) {
            uint32_t* c = cnt + 16 * a;
            uint32_t code = 0, prev = 0;
            for (uint32_t b = 1; b < 16; b++) {
                code = (code + prev) << 1;
                prev = c[b];
                c[b] = code;
            }
        }
    GLX_BLOCK_SYNC();
    for (uint32_t i = tid; i < GLX_BGZF_HIST; i += nt) {  // ... code = first + rank among earlier symbols of that length, bit-reversed
        const bool lit = i < 288;
        const uint32_t j = lit ? i : i - 288;
        const uint8_t* len = lit ? tab.lit_len : tab.dist_len;
        const uint32_t l = len[j];
        uint32_t rank = 0;
        for (uint32_t e = 0; e < j; e++) rank += len[e] == l;
        uint32_t v = l ? cnt[(lit ? 0 : 16) + l] + rank : 0u, r = 0;
        for (uint32_t b = 0; b < l; b++) {
            r = r << 1 | (v & 1);
            v >>= 1;
        }
        (lit ? tab.lit_code : tab.dist_code)[j] = (uint16_t)r;
        (lit ? tab.lit_cl : tab.dist_cl)[j] = r | l << 16;
    }
    for (uint32_t i = tid; i < GLX_BGZF_HDR_WORDS; i += nt) tab.hdr_words[i] = 0;
    GLX_BLOCK_SYNC();
    if (tid != 0) return;
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const uint32_t hlit = 286, hdist = 30;  // every symbol has a code
    uint32_t cl_freq[19];
    uint8_t cl_len[19];
    uint16_t cl_code[19];
    for (int i = 0; i < 19; i++) cl_freq[i] = 0;
    const uint32_t n_all = hlit + hdist;
    auto seq = [&](uint32_t i) -> uint32_t { return i < hlit ? tab.lit_len[i] : tab.dist_len[i - hlit]; };
    auto walk = [&](auto&& sink) {  // all lengths are >= 1 here: a length, then repeats of it in groups of 3..6
        uint32_t i = 0;
        while (i < n_all) {
            const uint32_t v = seq(i);
            uint32_t run = 1;
            while (i + run < n_all && seq(i + run) == v) run++;
            i += run;
            sink(v, 0, 0);
            run--;
            while (run >= 3) {
                const uint32_t t = run > 6 ? 6 : run;
                sink(16, t - 3, 2);
                run -= t;
            }
            while (run--) sink(v, 0, 0);
        }
    };
    walk([&](uint32_t sym, uint32_t, uint32_t) { cl_freq[sym]++; });
    bgzf_code_lengths(cl_freq, 19, 7, cl_len);
    {
        uint32_t used = 0;
        for (int k = 0; k < 19; k++) used += cl_len[k] != 0;
        if (used == 1) {  // a lone code-length symbol needs a partner for the code to be complete
            for (int k = 0; k < 19; k++)
                if (!cl_len[k]) {
                    cl_len[k] = 1;
                    break;
                }
        }
    }
    bgzf_assign_codes(cl_len, 19, cl_code);
    uint32_t hclen = 19;
    while (hclen > 4 && cl_len[order[hclen - 1]] == 0) hclen--;
    BgzfBits bw{tab.hdr_words, 0};
    bw.put(1, 1);  // BFINAL
    bw.put(2, 2);  // BTYPE = dynamic Huffman
    bw.put(hlit - 257, 5);
    bw.put(hdist - 1, 5);
    bw.put(hclen - 4, 4);
    for (uint32_t k = 0; k < hclen; k++) bw.put(cl_len[order[k]], 3);
    walk([&](uint32_t sym, uint32_t ev, uint32_t eb) {
        bw.put(cl_code[sym], cl_len[sym]);
        bw.put(ev, eb);
    });
    tab.hdr_bits = bw.pos;
}

// exclusive scan of wbits (+ header bits); one thread here, a block scan in the kernel
__host__ __device__ inline void bgzf_scan_bits_serial(BgzfShared& sh) {
    uint32_t acc = sh.tab.hdr_bits;
    for (uint32_t w = 0; w < GLX_BGZF_WORKERS; w++) {
        const uint32_t b = sh.wbits[w];
        sh.wbits[w] = acc;
        acc += b;
    }
    sh.data_bits = acc;  // position of the end-of-block code
}
// the header bits (whole words; the last, partial one is OR-ed because worker 0 shares it)
__host__ __device__ inline void bgzf_put_header(const BgzfShared& sh, uint32_t* out_words, uint32_t tid, uint32_t nt) {
    const uint32_t nw = (sh.tab.hdr_bits + 31) >> 5;
    for (uint32_t i = tid; i < nw; i += nt) bgzf_or(&out_words[i], sh.tab.hdr_words[i]);
}
// Each worker writes the codes of its recorded tokens at its bit offset: interior words are its own (plain stores), the two
// boundary words are shared (ORs). A token starts at every set bit of the worker's start mask and ends where the next one
// starts; a token of one byte is a literal. Literal-heavy chunks set the pace of their warp, so four literals out of one
// aligned word go out as two double codes (<= 30 bits each).
__host__ __device__ inline void bgzf_emit(BgzfShared& sh, const BcfSpanTab& T, const uint8_t* in, uint32_t n, uint32_t* out_words, uint32_t tid, uint32_t nt) {
    for (uint32_t w = tid; w < GLX_BGZF_WORKERS; w += nt) {
        const uint32_t c0 = w * GLX_BGZF_CHUNK;
        if (c0 >= n) continue;
        const uint32_t L = n - c0 < GLX_BGZF_CHUNK ? n - c0 : GLX_BGZF_CHUNK;
        const uint32_t pos = sh.wbits[w];
        uint32_t word = pos >> 5;
        unsigned long long acc = 0;
        uint32_t fill = pos & 31;  // the low `fill` bits of the first word belong to the previous worker
        bool first = true;
        auto put = [&](uint32_t v, uint32_t nb) {  // nb <= 32
            acc |= (unsigned long long)v << fill;
            fill += nb;
            if (fill >= 32) {
                if (first) bgzf_or(&out_words[word], (uint32_t)acc);
                else out_words[word] = (uint32_t)acc;
                first = false;
                word++;
                acc >>= 32;
                fill -= 32;
            }
        };
        const unsigned long long useD = sh.useD[w];
        BgzfChunkCtx C;
        if (useD) bgzf_chunk_ctx(T, c0, L, C);
        const uint32_t* W = reinterpret_cast<const uint32_t*>(in + BcfPadded::phys(c0));
        unsigned long long rest = sh.tok[w];  // the start mask shifted down to the current token: bit 0 is set
        uint32_t p = 0;
        while (p < L) {
            if (!(p & 3u) && p + 4 <= L && ((uint32_t)rest & 0xFu) == 0xFu && (p + 4 == L || ((uint32_t)rest >> 4 & 1u))) {
                const uint32_t x = W[p >> 2];
                const uint32_t a = sh.tab.lit_cl[x & 0xffu], b = sh.tab.lit_cl[x >> 8 & 0xffu], c = sh.tab.lit_cl[x >> 16 & 0xffu], d = sh.tab.lit_cl[x >> 24];
                put((a & 0xffffu) | (b & 0xffffu) << (a >> 16), (a >> 16) + (b >> 16));
                put((c & 0xffffu) | (d & 0xffffu) << (c >> 16), (c >> 16) + (d >> 16));
                rest >>= 4;
                p += 4;
                continue;
            }
            const unsigned long long nx = rest >> 1;
            const uint32_t len = nx ? 1u + bgzf_lowbit(nx) : L - p;  // start bits lie below L
            if (len < 3) {
                const uint32_t cl = sh.tab.lit_cl[(W[p >> 2] >> (8 * (p & 3))) & 0xffu];
                put(cl & 0xffffu, cl >> 16);
            } else {
                uint32_t s, eb, ev;
                bgzf_len_sym(len, s, eb, ev);
                uint32_t cl = sh.tab.lit_cl[s];
                put((cl & 0xffffu) | ev << (cl >> 16), (cl >> 16) + eb);  // code (<= 15 bits) and its extra bits (<= 5) in one go
                bgzf_dist_sym((useD >> p & 1ull) ? bgzf_D_at(C, p) : 1u, s, eb, ev);
                cl = sh.tab.dist_cl[s];
                put((cl & 0xffffu) | ev << (cl >> 16), (cl >> 16) + eb);  // <= 15 + 13 bits
            }
            rest = len < 64 ? rest >> len : 0ull;
            p += len;
        }
        if (fill) bgzf_or(&out_words[word], (uint32_t)acc);
    }
}

// one thread: end-of-block, stored fallback decision, total block length
__host__ __device__ inline void bgzf_finish_deflate(BgzfShared& sh, uint32_t* out_words) {
    BgzfBits bw{out_words, sh.data_bits};
    bw.put(sh.tab.lit_code[256], sh.tab.lit_len[256]);
    const uint32_t dbytes = (bw.pos + 7) >> 3;
    sh.stored = dbytes >= sh.n_in + 5 ? 1u : 0u;
    sh.block_len = 18 + (sh.stored ? sh.n_in + 5 : dbytes) + 8;
}
__host__ __device__ inline void bgzf_frame(BgzfShared& sh, uint8_t* stage) {  // one thread, after a stored block's bytes are in place
    uint8_t* b = stage + GLX_BGZF_LEAD;
    const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    for (int i = 0; i < 16; i++) b[i] = hdr[i];
    const uint32_t bsize = sh.block_len - 1;
    b[16] = (uint8_t)(bsize & 0xff);
    b[17] = (uint8_t)(bsize >> 8);
    uint8_t* t = b + sh.block_len - 8;
    for (int i = 0; i < 4; i++) t[i] = (uint8_t)(sh.crc >> (8 * i));
    for (int i = 0; i < 4; i++) t[4 + i] = (uint8_t)(sh.n_in >> (8 * i));
}
// stored fallback: deflate data = 01 LEN NLEN raw bytes
__host__ __device__ inline void bgzf_store_raw(const BgzfShared& sh, const uint8_t* in, uint8_t* stage, uint32_t tid, uint32_t nt) {
    uint8_t* d = stage + GLX_BGZF_LEAD + 18;
    const uint32_t n = sh.n_in;
    for (uint32_t i = tid; i < 5; i += nt) d[i] = i == 0 ? 1 : (i == 1 ? (uint8_t)(n & 0xff) : (i == 2 ? (uint8_t)(n >> 8) : (i == 3 ? (uint8_t)(~n & 0xff) : (uint8_t)((~n >> 8) & 0xff))));
    for (uint32_t i = tid; i < n; i += nt) d[5 + i] = in[BcfPadded::phys(i)];
}

}  // namespace glxd
