// glx_bgzf.cuh — BGZF blocks on the device: DEFLATE (RFC 1951, dynamic Huffman + LZ77 restricted to two structural
// distances) and CRC-32 of one <= 65,280-byte span of the BCF record stream, framed as a BGZF block (SAM spec §4.1).
//
// The reference writes its BCF through htslib's bgzf (single writer thread, zlib level -1; src/service.cc:254-275). A
// BGZF reader only requires each block to be a valid gzip member that inflates to the same bytes, so this encoder does
// not imitate zlib's bit stream; it is built for the data it sees. A BCF record is field-major: per FORMAT field one
// typed vector for all samples. Two redundancies dominate: (1) runs of one byte (GT of hom-ref cells, RNC "..",
// zero high bytes), caught by distance-1 matches; (2) a sample's value at this site equals its value at the previous
// site whenever the same reference band covers both, caught by matches at the distance D between the same FORMAT block of
// consecutive records (known from the layout, no hashing). Everything else is literals under a per-block Huffman code.
//
// Every phase is a block-level function over GLX_BGZF_WORKERS virtual workers (64-byte chunks of the span), written so
// that the host build (tests/emu) runs it with one thread and checks the result with zlib.
#pragma once
#include <cstdint>
#include <cstring>

#include "glx_bcf.cuh"

namespace glxd {

#define GLX_BGZF_WORKERS 1024u
#define GLX_BGZF_CHUNK 64u                      // bytes per worker: 1024 x 64 = 65,536 >= GLX_BCF_SPAN
#define GLX_BGZF_SLOT 65536u                    // bytes reserved per block before compaction
#define GLX_BGZF_LEAD 14u                       // the block starts at byte 14 of its staging buffer, so that the deflate data is word-aligned (14 + 18)
#define GLX_BGZF_OUT (GLX_BGZF_LEAD + 18u + 5u + GLX_BCF_SPAN + 8u + 16u)
#define GLX_BGZF_IN_PAD (GLX_BCF_SPAN + (GLX_BCF_SPAN / 64u + 2u) * 4u)  // the span with 4 bytes of padding after every 64 (BcfPadded)
#define GLX_BGZF_SEG_MAX 512u
#define GLX_BGZF_LITS 286u
#define GLX_BGZF_DISTS 30u

struct BgzfShared {
    uint32_t lit_freq[288], dist_freq[32];
    uint8_t lit_len[288], dist_len[32];
    uint16_t lit_code[288], dist_code[32];  // bit-reversed canonical codes
    uint32_t wbits[GLX_BGZF_WORKERS];       // CRC parts, then per-worker bit counts -> bit offsets
    unsigned long long tok[GLX_BGZF_WORKERS];   // bit i: a token starts at byte i of the worker's chunk (literal = 1 byte, match = to the next start)
    unsigned long long useD[GLX_BGZF_WORKERS];  // bit i: the match starting there copies from the structural distance D, not from distance 1
    uint16_t hdr_tok[320 + 8];                  // run-length coded code lengths of the block header: symbol | extra value << 5
    uint32_t seg_start[GLX_BGZF_SEG_MAX], seg_dist[GLX_BGZF_SEG_MAX];  // data segments of the span and their "same value, previous record" distance
    uint32_t n_seg, n_in, hdr_bits, data_bits, crc, stored, block_len, n_hdr_tok, lit_total, dist_total, lit_kraft, dist_kraft;
    uint32_t crc_table[256];
    uint32_t crc_shift[12];  // x^(8 * 64 * 2^l) mod P, l = 0..9; [10] = x^(8 n) mod P for this span
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
    uint32_t p = 1u << 31;  // x^0
    uint32_t sq = 1u << 23;  // x^8
    while (n) {
        if (n & 1) p = crc_multmodp(sq, p);
        sq = crc_multmodp(sq, sq);
        n >>= 1;
    }
    return p;
}
__host__ __device__ inline void bgzf_crc_setup(BgzfShared& sh, uint32_t n, uint32_t tid, uint32_t nt) {
    for (uint32_t i = tid; i < 256; i += nt) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = c & 1 ? (c >> 1) ^ GLX_CRC_POLY : c >> 1;
        sh.crc_table[i] = c;
    }
    for (uint32_t l = tid; l < 11; l += nt) sh.crc_shift[l] = l < 10 ? crc_x_pow_8n(GLX_BGZF_CHUNK << l) : crc_x_pow_8n(n);
}
// raw (init 0, no final xor) CRC of each worker's chunk; the data is right-aligned in the 1024 x 64 byte frame, so that
// leading zero padding leaves the raw register untouched and every combine step shifts by a fixed power
__host__ __device__ inline void bgzf_crc_parts(BgzfShared& sh, const uint8_t* in, uint32_t n, uint32_t tid, uint32_t nt) {
    const uint32_t pad = GLX_BGZF_WORKERS * GLX_BGZF_CHUNK - n;
    for (uint32_t w = tid; w < GLX_BGZF_WORKERS; w += nt) {
        const uint32_t v0 = w * GLX_BGZF_CHUNK, v1 = v0 + GLX_BGZF_CHUNK;
        uint32_t c = 0;
        for (uint32_t v = v0 < pad ? pad : v0; v < v1; v++) c = sh.crc_table[(c ^ in[BcfPadded::phys(v - pad)]) & 0xff] ^ (c >> 8);
        sh.wbits[w] = c;
    }
}
// one level of the combine tree: pairs (2i, 2i+1) at stride 2^l -> crc(A||B) = crc(A) * x^(8|B|) + crc(B)
__host__ __device__ inline void bgzf_crc_level(BgzfShared& sh, uint32_t l, uint32_t tid, uint32_t nt) {
    const uint32_t stride = 1u << l, pairs = GLX_BGZF_WORKERS >> (l + 1);
    for (uint32_t i = tid; i < pairs; i += nt) {
        const uint32_t a = i * 2 * stride, b = a + stride;
        sh.wbits[a] = crc_multmodp(sh.crc_shift[l], sh.wbits[a]) ^ sh.wbits[b];
    }
}
__host__ __device__ inline void bgzf_crc_finish(BgzfShared& sh) {  // one thread
    sh.crc = ~(sh.wbits[0] ^ crc_multmodp(sh.crc_shift[10], 0xffffffffu));
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

// Pass 1. Worker w owns bytes [64w, 64w+64) of the span (in = the BcfPadded staging buffer). It finds, per byte, whether the
// byte repeats the one before it (distance 1) and whether it repeats the byte at the structural distance D of its segment,
// as two 64-bit masks, then walks them greedily — a D-match of >= 4 bytes that is longer than the distance-1 run, else a
// distance-1 run of >= 3, else a literal — recording the token starts and counting symbols. Matches end at the chunk's end.
__host__ __device__ inline void bgzf_tokenize(BgzfShared& sh, const uint8_t* in, uint32_t n, uint32_t tid, uint32_t nt) {
    for (uint32_t w = tid; w < GLX_BGZF_WORKERS; w += nt) {
        const uint32_t c0 = w * GLX_BGZF_CHUNK;
        if (c0 >= n) {
            sh.tok[w] = sh.useD[w] = 0;
            continue;
        }
        const uint32_t L = n - c0 < GLX_BGZF_CHUNK ? n - c0 : GLX_BGZF_CHUNK;
        const unsigned long long live = L == 64 ? ~0ull : ((1ull << L) - 1);
        const uint32_t* W = reinterpret_cast<const uint32_t*>(in + BcfPadded::phys(c0));  // the chunk is contiguous and word-aligned
        unsigned long long eq1 = 0;
        uint32_t carry = c0 ? in[BcfPadded::phys(c0 - 1)] : 0u;
        for (uint32_t k = 0; k < 16; k++) {
            const uint32_t x = W[k];
            eq1 |= (unsigned long long)bgzf_eq_nibble(x, (x << 8) | carry) << (4 * k);
            carry = x >> 24;
        }
        if (!c0) eq1 &= ~1ull;  // the span's first byte has no predecessor
        eq1 &= live;
        unsigned long long eqD = 0, brk = 0;  // brk: a segment (another D) starts here: a D-match must not run across
        {  // segments overlapping the chunk: [a, b) with distance D
            uint32_t lo = 0, hi = sh.n_seg;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (sh.seg_start[mid] <= c0) lo = mid + 1;
                else hi = mid;
            }
            uint32_t sg = lo ? lo - 1 : 0;  // segment holding c0 (or the first one)
            for (; sg < sh.n_seg && sh.seg_start[sg] < c0 + L; sg++) {
                const uint32_t D = sh.seg_dist[sg];
                if (sh.seg_start[sg] > c0) brk |= 1ull << (sh.seg_start[sg] - c0);
                if (D <= 1 || D > 32768) continue;
                const uint32_t a = sh.seg_start[sg] > c0 ? sh.seg_start[sg] : c0;
                const uint32_t e = sg + 1 < sh.n_seg && sh.seg_start[sg + 1] < c0 + L ? sh.seg_start[sg + 1] : c0 + L;
                for (uint32_t i = a < D ? D : a; i < e; i++)
                    if (in[BcfPadded::phys(i)] == in[BcfPadded::phys(i - D)]) eqD |= 1ull << (i - c0);
            }
        }
        unsigned long long tok = 0, useD = 0;
        uint32_t i = 0;
        while (i < L) {
            const uint32_t r1 = bgzf_run(eq1, i);
            uint32_t rD = bgzf_run(eqD, i);
            if (i < 63) {
                const unsigned long long nb = brk >> (i + 1);
                if (nb) {
                    const uint32_t lim = bgzf_lowbit(nb) + 1;
                    rD = rD < lim ? rD : lim;
                }
            }
            tok |= 1ull << i;
            if (rD >= 4 && rD > r1) {
                useD |= 1ull << i;
                i += rD;
            } else if (r1 >= 3) {
                i += r1;
            } else {
                bgzf_add(&sh.lit_freq[in[BcfPadded::phys(c0 + i)]], 1);
                i++;
            }
        }
        sh.tok[w] = tok;
        sh.useD[w] = useD;
        // match symbols: second walk over the recorded starts (keeps the first loop short)
        unsigned long long t = tok;
        uint32_t sg = 0xffffffffu;
        while (t) {
            const uint32_t p = bgzf_lowbit(t);
            t &= t - 1;
            const uint32_t q = t ? bgzf_lowbit(t) : L;
            if (q - p < 3) continue;
            uint32_t sym, eb, ev;
            bgzf_len_sym(q - p, sym, eb, ev);
            bgzf_add(&sh.lit_freq[sym], 1);
            uint32_t D = 1;
            if (useD >> p & 1ull) {
                if (sg == 0xffffffffu) {
                    uint32_t lo = 0, hi = sh.n_seg;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (sh.seg_start[mid] <= c0 + p) lo = mid + 1;
                        else hi = mid;
                    }
                    sg = lo;
                } else
                    while (sg < sh.n_seg && sh.seg_start[sg] <= c0 + p) sg++;
                D = sh.seg_dist[sg - 1];
            }
            bgzf_dist_sym(D, sym, eb, ev);
            bgzf_add(&sh.dist_freq[sym], 1);
        }
    }
}

// the recorded tokens of worker w, in order: emit(is_match, literal byte | match length, distance)
template <class F>
__host__ __device__ inline void bgzf_replay(const BgzfShared& sh, const uint8_t* in, uint32_t n, uint32_t w, F&& emit) {
    const uint32_t c0 = w * GLX_BGZF_CHUNK;
    if (c0 >= n) return;
    const uint32_t L = n - c0 < GLX_BGZF_CHUNK ? n - c0 : GLX_BGZF_CHUNK;
    unsigned long long t = sh.tok[w];
    const unsigned long long useD = sh.useD[w];
    uint32_t sg = 0xffffffffu;
    while (t) {
        const uint32_t p = bgzf_lowbit(t);
        t &= t - 1;
        const uint32_t q = t ? bgzf_lowbit(t) : L;
        if (q - p < 3) {
            emit(false, (uint32_t)in[BcfPadded::phys(c0 + p)], 0u);
            continue;
        }
        uint32_t D = 1;
        if (useD >> p & 1ull) {
            if (sg == 0xffffffffu) {
                uint32_t lo = 0, hi = sh.n_seg;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (sh.seg_start[mid] <= c0 + p) lo = mid + 1;
                    else hi = mid;
                }
                sg = lo;
            } else
                while (sg < sh.n_seg && sh.seg_start[sg] <= c0 + p) sg++;
            D = sh.seg_dist[sg - 1];
        }
        emit(true, q - p, D);
    }
}

__host__ __device__ inline void bgzf_hist_init(BgzfShared& sh, uint32_t tid, uint32_t nt) {
    for (uint32_t i = tid; i < 288; i += nt) sh.lit_freq[i] = i == 256 ? 1u : 0u;  // one end-of-block
    for (uint32_t i = tid; i < 32; i += nt) sh.dist_freq[i] = 0;
}

// ---- Huffman codes. Code lengths <= maxbits forming a COMPLETE prefix code (zlib's inflate rejects incomplete literal/length
// sets): start from ceil(log2(total / f)) per symbol (parallel), repair the Kraft sum if the clamp at maxbits oversubscribed
// it, then hand the slack out by shortening symbols until the sum is exactly one (one thread per alphabet). ----
__host__ __device__ __forceinline__ uint32_t bgzf_len0(uint32_t f, uint32_t total, uint32_t maxbits) {
    uint32_t l = 1;
    while (l < maxbits && ((unsigned long long)f << l) < total) l++;
    return l;
}
// serial part for one alphabet: lengths initialised, K = their Kraft sum in units of 2^-maxbits, used = symbols with f > 0
__host__ __device__ inline void bgzf_len_settle(const uint32_t* freq, uint32_t n, uint32_t maxbits, uint8_t* len, uint32_t K, uint32_t used) {
    if (used == 0) return;
    if (used == 1) {
        for (uint32_t i = 0; i < n; i++)
            if (freq[i]) len[i] = 1;
        return;
    }
    const uint32_t full = 1u << maxbits;
    for (uint32_t L = maxbits - 1; K > full && L >= 1; L--)  // oversubscribed: lengthen the longest codes first
        for (uint32_t i = 0; i < n && K > full; i++)
            if (len[i] == L) {
                len[i]++;
                K -= full >> (L + 1);
            }
    uint32_t slack = full - K;
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
// stand-alone form (the 19-symbol code-length alphabet of the header). One thread.
__host__ __device__ inline void bgzf_code_lengths(const uint32_t* freq, uint32_t n, uint32_t maxbits, uint8_t* len) {
    uint32_t used = 0, total = 0, K = 0;
    for (uint32_t i = 0; i < n; i++) {
        len[i] = 0;
        used += freq[i] != 0;
        total += freq[i];
    }
    for (uint32_t i = 0; i < n; i++)
        if (freq[i]) {
            len[i] = (uint8_t)bgzf_len0(freq[i], total, maxbits);
            K += (1u << maxbits) >> len[i];
        }
    bgzf_len_settle(freq, n, maxbits, len, K, used);
}
__host__ __device__ __forceinline__ uint32_t bgzf_rev(uint32_t v, uint32_t l) {  // low l bits of v, reversed
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - l);
#else
    uint32_t r = 0;
    for (uint32_t b = 0; b < l; b++) {
        r = r << 1 | (v & 1);
        v >>= 1;
    }
    return r;
#endif
}
// canonical codes, bit-reversed for LSB-first packing. One thread (small alphabets).
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
        code[i] = l ? (uint16_t)bgzf_rev(next[l]++, l) : (uint16_t)0;
    }
}

// serial bit writer over a zeroed, word-aligned buffer (the block header, end-of-block)
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

// Dynamic-block header (BFINAL = 1, BTYPE = 10) from the two length tables. One thread. Sets sh.hdr_bits.
__host__ __device__ inline void bgzf_write_header(BgzfShared& sh, uint32_t* out_words) {
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint32_t hlit = 286, hdist = 30;
    while (hlit > 257 && sh.lit_len[hlit - 1] == 0) hlit--;
    while (hdist > 1 && sh.dist_len[hdist - 1] == 0) hdist--;
    // run-length code of the concatenated length sequence, kept as tokens (symbol | extra value << 5)
    uint32_t cl_freq[19];
    uint8_t cl_len[19];
    uint16_t cl_code[19];
    for (int i = 0; i < 19; i++) cl_freq[i] = 0;
    const uint32_t n_all = hlit + hdist;
    uint32_t nt = 0;
    auto sink = [&](uint32_t sym, uint32_t ev) {
        sh.hdr_tok[nt++] = (uint16_t)(sym | ev << 5);
        cl_freq[sym]++;
    };
    uint32_t i = 0;
    while (i < n_all) {
        const uint32_t v = i < hlit ? sh.lit_len[i] : sh.dist_len[i - hlit];
        uint32_t run = 1;
        while (i + run < n_all && (i + run < hlit ? sh.lit_len[i + run] : sh.dist_len[i + run - hlit]) == v) run++;
        i += run;
        if (v == 0) {
            while (run >= 11) {
                const uint32_t t = run > 138 ? 138 : run;
                sink(18, t - 11);
                run -= t;
            }
            if (run >= 3) {
                sink(17, run - 3);
                run = 0;
            }
            while (run--) sink(0, 0);
        } else {
            sink(v, 0);
            run--;
            while (run >= 3) {
                const uint32_t t = run > 6 ? 6 : run;
                sink(16, t - 3);
                run -= t;
            }
            while (run--) sink(v, 0);
        }
    }
    sh.n_hdr_tok = nt;
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
    BgzfBits bw{out_words, 0};
    bw.put(1, 1);  // BFINAL
    bw.put(2, 2);  // BTYPE = dynamic Huffman
    bw.put(hlit - 257, 5);
    bw.put(hdist - 1, 5);
    bw.put(hclen - 4, 4);
    for (uint32_t k = 0; k < hclen; k++) bw.put(cl_len[order[k]], 3);
    for (uint32_t k = 0; k < nt; k++) {
        const uint32_t sym = sh.hdr_tok[k] & 31u, ev = sh.hdr_tok[k] >> 5;
        bw.put(cl_code[sym], cl_len[sym]);
        bw.put(ev, sym == 16 ? 2u : (sym == 17 ? 3u : (sym == 18 ? 7u : 0u)));
    }
    sh.hdr_bits = bw.pos;
}

// Both Huffman codes and the block header, by the whole block (every thread calls it). out_words = word-aligned, zeroed
// start of the deflate data.
__host__ __device__ inline void bgzf_build_codes(BgzfShared& sh, uint32_t* out_words, uint32_t tid, uint32_t nt) {
    uint32_t* cnt = sh.wbits;  // scratch: [0..16) literal/length code-length counts, [16..32) distance, [32..34) used symbols
    for (uint32_t i = tid; i < 40; i += nt) cnt[i] = 0;
    if (tid == 0) sh.lit_total = sh.dist_total = sh.lit_kraft = sh.dist_kraft = 0;
    GLX_BLOCK_SYNC();
    for (uint32_t i = tid; i < 288 + 32; i += nt) {
        const bool lit = i < 288;
        const uint32_t f = lit ? (i < GLX_BGZF_LITS ? sh.lit_freq[i] : 0u) : (i - 288 < GLX_BGZF_DISTS ? sh.dist_freq[i - 288] : 0u);
        if (f) {
            bgzf_add(lit ? &sh.lit_total : &sh.dist_total, f);
            bgzf_add(&cnt[lit ? 32 : 33], 1);
        }
    }
    GLX_BLOCK_SYNC();
    for (uint32_t i = tid; i < 288 + 32; i += nt) {
        const bool lit = i < 288;
        const uint32_t j = lit ? i : i - 288;
        const uint32_t f = lit ? (j < GLX_BGZF_LITS ? sh.lit_freq[j] : 0u) : (j < GLX_BGZF_DISTS ? sh.dist_freq[j] : 0u);
        uint32_t l = 0;
        if (f) {
            l = bgzf_len0(f, lit ? sh.lit_total : sh.dist_total, 15);
            bgzf_add(lit ? &sh.lit_kraft : &sh.dist_kraft, 32768u >> l);
        }
        (lit ? sh.lit_len : sh.dist_len)[j] = (uint8_t)l;
    }
    GLX_BLOCK_SYNC();
    // the two serial settles run side by side on two warps
    if (tid == 0) bgzf_len_settle(sh.lit_freq, GLX_BGZF_LITS, 15, sh.lit_len, sh.lit_kraft, cnt[32]);
    if (tid == (nt > 32 ? 32u : 0u)) {
        bgzf_len_settle(sh.dist_freq, GLX_BGZF_DISTS, 15, sh.dist_len, sh.dist_kraft, cnt[33]);
        if (cnt[33] == 0) sh.dist_len[0] = 1;  // "one distance code of one bit" keeps every inflater happy
    }
    GLX_BLOCK_SYNC();
    // canonical codes: count per length, first code per length, then code = first + rank among earlier symbols of that length
    for (uint32_t i = tid; i < 288 + 32; i += nt) {
        const bool lit = i < 288;
        const uint32_t l = lit ? sh.lit_len[i] : sh.dist_len[i - 288];
        if (l) bgzf_add(&cnt[(lit ? 0 : 16) + l], 1);
    }
    GLX_BLOCK_SYNC();
    if (tid == 0 || tid == (nt > 32 ? 32u : 0u)) {
        for (uint32_t a = 0; a < 2; a++) {
            if (nt > 32 && a != (tid ? 1u : 0u)) continue;
            uint32_t* c = cnt + 16 * a;
            uint32_t code = 0, prev = 0;
            for (uint32_t b = 1; b < 16; b++) {  // c[b]: count -> first code of length b
                code = (code + prev) << 1;
                prev = c[b];
                c[b] = code;
            }
        }
    }
    GLX_BLOCK_SYNC();
    for (uint32_t i = tid; i < 288 + 32; i += nt) {
        const bool lit = i < 288;
        const uint32_t j = lit ? i : i - 288;
        const uint8_t* len = lit ? sh.lit_len : sh.dist_len;
        const uint32_t l = len[j];
        uint32_t rank = 0;
        for (uint32_t e = 0; e < j; e++) rank += len[e] == l;
        (lit ? sh.lit_code : sh.dist_code)[j] = l ? (uint16_t)bgzf_rev(cnt[(lit ? 0 : 16) + l] + rank, l) : (uint16_t)0;
    }
    GLX_BLOCK_SYNC();
    if (tid == 0) bgzf_write_header(sh, out_words);
    GLX_BLOCK_SYNC();
}

// per-worker bit counts
__host__ __device__ inline void bgzf_count_bits(BgzfShared& sh, const uint8_t* in, uint32_t n, uint32_t tid, uint32_t nt) {
    for (uint32_t w = tid; w < GLX_BGZF_WORKERS; w += nt) {
        uint32_t bits = 0;
        bgzf_replay(sh, in, n, w, [&](bool match, uint32_t v, uint32_t d) {
            if (!match) bits += sh.lit_len[v];
            else {
                uint32_t s, eb, ev;
                bgzf_len_sym(v, s, eb, ev);
                bits += sh.lit_len[s] + eb;
                bgzf_dist_sym(d, s, eb, ev);
                bits += sh.dist_len[s] + eb;
            }
        });
        sh.wbits[w] = bits;
    }
}
// exclusive scan of wbits (+ header bits); one thread here, a block scan in the kernel
__host__ __device__ inline void bgzf_scan_bits_serial(BgzfShared& sh) {
    uint32_t acc = sh.hdr_bits;
    for (uint32_t w = 0; w < GLX_BGZF_WORKERS; w++) {
        const uint32_t b = sh.wbits[w];
        sh.wbits[w] = acc;
        acc += b;
    }
    sh.data_bits = acc;  // position of the end-of-block code
}
// each worker writes its codes at its bit offset: interior words are its own (plain stores), the two boundary words are shared
__host__ __device__ inline void bgzf_emit(BgzfShared& sh, const uint8_t* in, uint32_t n, uint32_t* out_words, uint32_t tid, uint32_t nt) {
    for (uint32_t w = tid; w < GLX_BGZF_WORKERS; w += nt) {
        if (w * GLX_BGZF_CHUNK >= n) continue;
        const uint32_t pos = sh.wbits[w];
        uint32_t word = pos >> 5;
        unsigned long long acc = 0;
        uint32_t fill = pos & 31;  // the low `fill` bits of the first word belong to the previous worker
        bool first = true;
        auto put = [&](uint32_t v, uint32_t nb) {
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
        bgzf_replay(sh, in, n, w, [&](bool match, uint32_t v, uint32_t d) {
            if (!match) put(sh.lit_code[v], sh.lit_len[v]);
            else {
                uint32_t s, eb, ev;
                bgzf_len_sym(v, s, eb, ev);
                put(sh.lit_code[s], sh.lit_len[s]);
                put(ev, eb);
                bgzf_dist_sym(d, s, eb, ev);
                put(sh.dist_code[s], sh.dist_len[s]);
                put(ev, eb);
            }
        });
        if (fill) bgzf_or(&out_words[word], (uint32_t)acc);
    }
}

// one thread: end-of-block, stored fallback decision, BGZF header and trailer. `stage` = staging buffer (block at stage + GLX_BGZF_LEAD).
// Returns via sh.block_len the total BGZF block size; sh.stored says whether the caller must copy the raw bytes (stored block).
__host__ __device__ inline void bgzf_finish_deflate(BgzfShared& sh, uint32_t* out_words) {
    BgzfBits bw{out_words, sh.data_bits};
    bw.put(sh.lit_code[256], sh.lit_len[256]);
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

// "same value at the previous record" distances of the span's FORMAT data segments (one thread). Segment starts are
// span-relative; a segment that begins before the span starts at 0.
__host__ __device__ inline void bgzf_segments(BgzfShared& sh, const BcfEnc& E, unsigned long long lo, unsigned long long hi) {
    uint32_t ns = 0;
    uint32_t prev = 0xffffffffu;  // previous site that has a record
    const uint32_t s0 = bcf_first_site(E, lo);
    for (uint32_t p = s0; p-- > 0;)
        if (E.rec_off[p + 1] > E.rec_off[p]) {
            prev = p;
            break;
        }
    for (uint32_t s = s0; s < E.n_sites && E.rec_off[s] < hi && ns + 2 < GLX_BGZF_SEG_MAX; s++) {
        const unsigned long long r0 = E.rec_off[s];
        if (E.rec_off[s + 1] == r0) continue;
        const BcfSitePlan& P = E.plan[s];
        const unsigned long long ibase = r0 + 8 + P.l_shared;
        // header + shared bytes: no structural distance
        if (r0 >= lo || s == s0) {
            sh.seg_start[ns] = r0 > lo ? (uint32_t)(r0 - lo) : 0u;
            sh.seg_dist[ns] = 0;
            ns++;
        }
        for (uint32_t b = 0; b < P.n_blk && ns + 2 < GLX_BGZF_SEG_MAX; b++) {
            const unsigned long long dstart = ibase + P.blk_off[b] + P.blk_hdr_len[b];
            const unsigned long long dend = dstart + (unsigned long long)E.n_samples * P.blk_cnt[b] * P.blk_width[b];
            if (dend <= lo) continue;
            if (dstart >= hi) break;
            uint32_t dist = 0;
            if (prev != 0xffffffffu) {
                const BcfSitePlan& Q = E.plan[prev];
                if (Q.n_blk == P.n_blk && Q.blk_cnt[b] == P.blk_cnt[b] && Q.blk_width[b] == P.blk_width[b]) {
                    const unsigned long long pstart = E.rec_off[prev] + 8 + Q.l_shared + Q.blk_off[b] + Q.blk_hdr_len[b];
                    const unsigned long long dd = dstart - pstart;
                    if (dd <= 32768) dist = (uint32_t)dd;
                }
            }
            // the block's own key/descriptor bytes belong to the previous segment; data starts here
            sh.seg_start[ns] = dstart > lo ? (uint32_t)(dstart - lo) : 0u;
            sh.seg_dist[ns] = dist;
            ns++;
        }
        prev = s;
    }
    sh.n_seg = ns;
}

}  // namespace glxd
