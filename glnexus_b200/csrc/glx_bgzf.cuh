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
#define GLX_BGZF_SEG_MAX 512u
#define GLX_BGZF_LITS 286u
#define GLX_BGZF_DISTS 30u

struct BgzfShared {
    uint32_t lit_freq[288], dist_freq[32];
    uint8_t lit_len[288], dist_len[32];
    uint16_t lit_code[288], dist_code[32];  // bit-reversed canonical codes
    uint32_t wbits[GLX_BGZF_WORKERS];       // CRC parts, then per-worker bit counts -> bit offsets
    uint32_t seg_start[GLX_BGZF_SEG_MAX], seg_dist[GLX_BGZF_SEG_MAX];  // data segments of the span and their "same value, previous record" distance
    uint32_t n_seg, n_in, hdr_bits, data_bits, crc, stored, block_len;
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
        for (uint32_t v = v0 < pad ? pad : v0; v < v1; v++) c = sh.crc_table[(c ^ in[v - pad]) & 0xff] ^ (c >> 8);
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

// ---- tokenizer: worker chunk [c0, c1) of in[0, n) -> literals and matches at distance 1 or D(position) ----
template <class F>
__host__ __device__ inline void bgzf_tokenize(const BgzfShared& sh, const uint8_t* in, uint32_t c0, uint32_t c1, F&& emit) {
    // segment of c0: last segment with start <= c0
    uint32_t sg = 0;
    {
        uint32_t a = 0, b = sh.n_seg;
        while (a < b) {
            const uint32_t mid = (a + b) >> 1;
            if (sh.seg_start[mid] <= c0) a = mid + 1;
            else b = mid;
        }
        sg = a;  // number of segments starting at or before c0
    }
    uint32_t i = c0;
    while (i < c1) {
        while (sg < sh.n_seg && sh.seg_start[sg] <= i) sg++;
        const uint32_t D = sg ? sh.seg_dist[sg - 1] : 0;
        const uint32_t maxlen = c1 - i < 258 ? c1 - i : 258;
        uint32_t l1 = 0, lD = 0;
        if (i >= 1) {
            const uint8_t b = in[i - 1];
            while (l1 < maxlen && in[i + l1] == b) l1++;
        }
        if (D > 1 && D <= i && D <= 32768) {
            while (lD < maxlen && in[i + lD] == in[i + lD - D]) lD++;
        }
        if (lD >= 4 && lD > l1) {
            emit(true, lD, D);
            i += lD;
        } else if (l1 >= 3) {
            emit(true, l1, 1u);
            i += l1;
        } else {
            emit(false, (uint32_t)in[i], 0u);
            i++;
        }
    }
}

__host__ __device__ inline void bgzf_hist_init(BgzfShared& sh, uint32_t tid, uint32_t nt) {
    for (uint32_t i = tid; i < 288; i += nt) sh.lit_freq[i] = i == 256 ? 1u : 0u;  // one end-of-block
    for (uint32_t i = tid; i < 32; i += nt) sh.dist_freq[i] = 0;
}
__host__ __device__ inline void bgzf_hist(BgzfShared& sh, const uint8_t* in, uint32_t n, uint32_t tid, uint32_t nt) {
    for (uint32_t w = tid; w < GLX_BGZF_WORKERS; w += nt) {
        const uint32_t c0 = w * GLX_BGZF_CHUNK, c1 = c0 + GLX_BGZF_CHUNK < n ? c0 + GLX_BGZF_CHUNK : n;
        if (c0 >= n) continue;
        bgzf_tokenize(sh, in, c0, c1, [&](bool match, uint32_t v, uint32_t d) {
            if (!match) bgzf_add(&sh.lit_freq[v], 1);
            else {
                uint32_t s, eb, ev;
                bgzf_len_sym(v, s, eb, ev);
                bgzf_add(&sh.lit_freq[s], 1);
                bgzf_dist_sym(d, s, eb, ev);
                bgzf_add(&sh.dist_freq[s], 1);
            }
        });
    }
}

// Code lengths <= maxbits forming a COMPLETE prefix code (zlib's inflate rejects incomplete literal/length sets): start from
// ceil(log2(total / f)), repair the Kraft sum if the clamp at maxbits oversubscribed it, then hand the slack out by
// shortening symbols until the sum is exactly one. One thread; n <= 288.
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
// canonical codes, bit-reversed for LSB-first packing. One thread.
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

// Dynamic-block header (BFINAL = 1, BTYPE = 10) from the two length tables. One thread. Returns the header's bit count.
__host__ __device__ inline uint32_t bgzf_write_header(BgzfShared& sh, uint32_t* out_words) {
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint32_t hlit = 286, hdist = 30;
    while (hlit > 257 && sh.lit_len[hlit - 1] == 0) hlit--;
    while (hdist > 1 && sh.dist_len[hdist - 1] == 0) hdist--;
    // run-length code of the concatenated length sequence: two passes (count, then write)
    uint32_t cl_freq[19];
    uint8_t cl_len[19];
    uint16_t cl_code[19];
    for (int i = 0; i < 19; i++) cl_freq[i] = 0;
    const uint32_t n_all = hlit + hdist;
    auto seq = [&](uint32_t i) -> uint32_t { return i < hlit ? sh.lit_len[i] : sh.dist_len[i - hlit]; };
    auto walk = [&](auto&& sink) {
        uint32_t i = 0;
        while (i < n_all) {
            const uint32_t v = seq(i);
            uint32_t run = 1;
            while (i + run < n_all && seq(i + run) == v) run++;
            if (v == 0) {
                uint32_t r = run;
                while (r >= 11) {
                    const uint32_t t = r > 138 ? 138 : r;
                    sink(18, t - 11, 7);
                    r -= t;
                }
                if (r >= 3) {
                    sink(17, r - 3, 3);
                    r = 0;
                }
                while (r--) sink(0, 0, 0);
            } else {
                sink(v, 0, 0);
                uint32_t r = run - 1;
                while (r >= 3) {
                    const uint32_t t = r > 6 ? 6 : r;
                    sink(16, t - 3, 2);
                    r -= t;
                }
                while (r--) sink(v, 0, 0);
            }
            i += run;
        }
    };
    walk([&](uint32_t sym, uint32_t, uint32_t) { cl_freq[sym]++; });
    bgzf_code_lengths(cl_freq, 19, 7, cl_len);
    {
        uint32_t used = 0;
        for (int i = 0; i < 19; i++) used += cl_len[i] != 0;
        if (used == 1) {  // a lone code-length symbol needs a partner for the code to be complete
            for (int i = 0; i < 19; i++)
                if (!cl_len[i]) {
                    cl_len[i] = 1;
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
    for (uint32_t i = 0; i < hclen; i++) bw.put(cl_len[order[i]], 3);
    walk([&](uint32_t sym, uint32_t ev, uint32_t eb) {
        bw.put(cl_code[sym], cl_len[sym]);
        bw.put(ev, eb);
    });
    return bw.pos;
}

// one thread: both Huffman codes + the header. out_words = word-aligned start of the deflate data (zeroed)
__host__ __device__ inline void bgzf_build_codes(BgzfShared& sh, uint32_t* out_words) {
    bgzf_code_lengths(sh.lit_freq, GLX_BGZF_LITS, 15, sh.lit_len);
    bgzf_code_lengths(sh.dist_freq, GLX_BGZF_DISTS, 15, sh.dist_len);
    {
        uint32_t used = 0;
        for (uint32_t i = 0; i < GLX_BGZF_DISTS; i++) used += sh.dist_len[i] != 0;
        if (used == 0) sh.dist_len[0] = 1;  // "one distance code of one bit" keeps every inflater happy
    }
    bgzf_assign_codes(sh.lit_len, GLX_BGZF_LITS, sh.lit_code);
    bgzf_assign_codes(sh.dist_len, GLX_BGZF_DISTS, sh.dist_code);
    sh.hdr_bits = bgzf_write_header(sh, out_words);
}

// per-worker bit counts
__host__ __device__ inline void bgzf_count_bits(BgzfShared& sh, const uint8_t* in, uint32_t n, uint32_t tid, uint32_t nt) {
    for (uint32_t w = tid; w < GLX_BGZF_WORKERS; w += nt) {
        const uint32_t c0 = w * GLX_BGZF_CHUNK, c1 = c0 + GLX_BGZF_CHUNK < n ? c0 + GLX_BGZF_CHUNK : n;
        uint32_t bits = 0;
        if (c0 < n)
            bgzf_tokenize(sh, in, c0, c1, [&](bool match, uint32_t v, uint32_t d) {
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
        const uint32_t c0 = w * GLX_BGZF_CHUNK, c1 = c0 + GLX_BGZF_CHUNK < n ? c0 + GLX_BGZF_CHUNK : n;
        if (c0 >= n) continue;
        uint32_t pos = sh.wbits[w];
        uint32_t word = pos >> 5;
        unsigned long long acc = 0;
        uint32_t fill = pos & 31;  // bits of `acc` already spoken for (the low `fill` bits belong to the previous worker)
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
        bgzf_tokenize(sh, in, c0, c1, [&](bool match, uint32_t v, uint32_t d) {
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
    for (uint32_t i = tid; i < n; i += nt) d[5 + i] = in[i];
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
