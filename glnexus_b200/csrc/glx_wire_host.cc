// glx_wire_host.cc — builds the compact wire form of a packed batch's record store (glx_wire, include/glx.h): what crosses
// PCIe instead of the kernel-facing columns. SURVEY.md §8f N2 (start): the packer's output at the width the data needs.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>

#include "glx_host.hpp"

namespace glx {

namespace {
// fn(t, lo, hi) over [0, n) cut into T contiguous chunks, one thread each
template <class F>
void parallel_chunks(uint64_t n, unsigned T, F&& fn) {
    if (T <= 1 || n < 65536) {
        fn(0u, (uint64_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; t++) pool.emplace_back([&, t]() { fn(t, n * t / T, n * (t + 1) / T); });
    for (auto& th : pool) th.join();
}
size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
bool fits16(int32_t v) { return v == GLX_INT_MISSING || v == GLX_INT_VECTOR_END || (v >= -32760 && v <= 32767); }
int16_t to16(int32_t v) { return v == GLX_INT_MISSING ? (int16_t)0x8000 : (v == GLX_INT_VECTOR_END ? (int16_t)0x8001 : (int16_t)v); }
}  // namespace

// rows of the shape table (GLX_WF_SHAPES): 255 = as many as a code byte can name; tests lower it so that records of rarer shapes
// take the exception path on small inputs
static std::atomic<uint32_t> g_shape_cap{255};
void set_wire_shape_cap(uint32_t cap) { g_shape_cap = cap > 255 ? 255 : cap; }

Status build_wire(const PackedBatch& b, WireBytes& w) {
    const glx_records& r = b.recs;
    const uint64_t R = r.n_records, D = r.n_datasets, C = r.n_cols;
    if (r.n_datasets != r.n_samples_out) return Status::NotImplemented("build_wire: multi-sample datasets have no wire form");
    for (uint32_t d = 0; d < D; d++)
        if (r.ds_n_sample[d] != 1 || r.sample_out[r.ds_sample_off[d]] != (int32_t)d) return Status::NotImplemented("build_wire: multi-sample datasets have no wire form");
    if (R >= (1ull << 32)) return Status::NotImplemented("build_wire: more than 2^32 records");
    glx_wire h;
    memset(&h, 0, sizeof h);
    h.magic = GLX_WIRE_MAGIC;
    h.n_records = R;
    h.n_datasets = (uint32_t)D;
    h.n_cols = (uint32_t)C;
    h.n_al = r.n_al;
    h.n_al_pool = r.n_al_pool;
    h.n_pool = r.n_pool;
    // ---- what fits (one pass over the records on T threads; per-chunk results merged below) ----
    const unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    struct Part {
        bool len16 = true, gt8 = true, collen8 = true, pool16 = true, pool_order = true, filt8 = true;
        uint64_t n_variant = 0, pool_units = 0, al_units = 0;
        bool al_order = true;
        std::vector<uint32_t> exc;
        bool any[GLX_MAX_COLS] = {}, f16[GLX_MAX_COLS];
    };
    std::vector<Part> parts(T);
    std::vector<uint64_t> chunk_lo(T + 1, R);
    parallel_chunks(R, T, [&](unsigned t, uint64_t lo, uint64_t hi) {
        Part& P = parts[t];
        chunk_lo[t] = lo;
        for (uint64_t c = 0; c < C; c++) P.f16[c] = true;
        for (uint64_t i = lo; i < hi; i++) {
            const int64_t len = (int64_t)r.end[i] - r.beg[i];
            if (len < 0 || len > 65535) P.len16 = false;
            if (r.filt[i] > 255u) P.filt8 = false;
            if (!(r.flags[i] & GLX_RF_IS_REF)) {
                P.n_variant++;
                P.al_units += r.n_allele[i];
            }
            for (int hlf = 0; hlf < 2; hlf++) {
                const int16_t g = (int16_t)(r.gt[i] >> (16 * hlf));
                if (!(g == GLX_GT16_VECTOR_END || (g >= 0 && g <= 508 && !(g & 1)))) P.gt8 = false;
            }
            if (r.maxend[i] != r.end[i]) {
                P.exc.push_back((uint32_t)i);
                P.exc.push_back((uint32_t)r.maxend[i]);
            }
            for (uint64_t c = 0; c < C; c++) {
                const uint16_t cl = r.col_len[c * R + i];
                if (cl >= 0xFD && cl < 0xFFF0) P.collen8 = false;
                if (cl == 1) {
                    P.any[c] = true;
                    if (!fits16(r.col_val[c * R + i])) P.f16[c] = false;
                } else if (cl < 0xFFF0)
                    P.pool_units += cl;
            }
        }
    });
    bool len16 = true, gt8 = true, collen8 = true, pool16 = true, allen16 = true, allen8 = true, filt8 = true;
    uint64_t n_variant = 0;
    std::vector<uint32_t> exc;  // (record, maxend) pairs where maxend != end
    std::vector<uint64_t> pool_base(T + 1, 0), var_base(T + 1, 0), al_base(T + 1, 0);
    for (unsigned t = 0; t < T; t++) {
        const Part& P = parts[t];
        len16 = len16 && P.len16;
        gt8 = gt8 && P.gt8;
        collen8 = collen8 && P.collen8;
        filt8 = filt8 && P.filt8;
        exc.insert(exc.end(), P.exc.begin(), P.exc.end());
        pool_base[t + 1] = pool_base[t] + P.pool_units;
        var_base[t + 1] = var_base[t] + P.n_variant;
        al_base[t + 1] = al_base[t] + P.al_units;
        for (uint64_t c = 0; c < C; c++) h.col_w[c] = std::max<uint8_t>(h.col_w[c], !P.any[c] ? 0 : (P.f16[c] ? 2 : 4));
    }
    n_variant = var_base[T];
    if (pool_base[T] != r.n_pool) return Status::NotImplemented("build_wire: value pool has entries no record refers to");
    // the pool must be laid out record-major, column-minor (what pack_records / slice_batch / the generator write): offsets are then prefix sums
    // ... and the allele descriptors must follow the non-reference records in order (their offsets are prefix sums too)
    parallel_chunks(R, T, [&](unsigned t, uint64_t lo, uint64_t hi) {
        uint64_t run = pool_base[t], ao = al_base[t];
        for (uint64_t i = lo; i < hi; i++) {
            for (uint64_t c = 0; c < C; c++) {
                const uint16_t cl = r.col_len[c * R + i];
                if (cl >= 0xFFF0 || cl == 1) continue;
                if ((uint64_t)(uint32_t)r.col_val[c * R + i] != run) parts[t].pool_order = false;
                run += cl;
            }
            if (!(r.flags[i] & GLX_RF_IS_REF)) {
                if (r.allele_off[i] != ao) parts[t].al_order = false;
                ao += r.n_allele[i];
            }
        }
    });
    for (unsigned t = 0; t < T; t++) {
        if (!parts[t].pool_order) return Status::NotImplemented("build_wire: value pool is not in record order");
        if (!parts[t].al_order) return Status::NotImplemented("build_wire: allele descriptors are not in record order");
    }
    if (al_base[T] != r.n_al) return Status::NotImplemented("build_wire: allele descriptors no record refers to");
    {
        std::vector<char> ok(T, 1);
        parallel_chunks(r.n_pool, T, [&](unsigned t, uint64_t lo, uint64_t hi) {
            for (uint64_t i = lo; i < hi; i++)
                if (!fits16(r.pool[i])) {
                    ok[t] = 0;
                    break;
                }
        });
        for (char o : ok) pool16 = pool16 && o;
    }
    // allele strings must be laid out in allele order for their offsets to be prefix sums (checked chunk by chunk: the lengths' sums first)
    {
        std::vector<uint64_t> len_base(T + 1, 0);
        std::vector<char> ok16(T, 1), ok8(T, 1), okstr(T, 1);
        parallel_chunks(r.n_al, T, [&](unsigned t, uint64_t lo, uint64_t hi) {
            uint64_t sum = 0;
            for (uint64_t a = lo; a < hi; a++) {
                sum += r.al_len[a];
                if (r.al_len[a] > 65535) ok16[t] = 0;
                if (r.al_len[a] > 255) ok8[t] = 0;
            }
            len_base[t + 1] = sum;
        });
        for (unsigned t = 0; t < T; t++) len_base[t + 1] += len_base[t];
        parallel_chunks(r.n_al, T, [&](unsigned t, uint64_t lo, uint64_t hi) {
            uint64_t run = len_base[t];  // (one chunk when the work is small: parallel_chunks then passes t = 0, lo = 0)
            for (uint64_t a = lo; a < hi; a++) {
                if (r.al_str[a] != run) okstr[t] = 0;
                run += r.al_len[a];
            }
        });
        for (unsigned t = 0; t < T; t++) {
            allen16 = allen16 && ok16[t];
            allen8 = allen8 && ok8[t];
            if (!okstr[t]) return Status::NotImplemented("build_wire: allele strings are not in allele order");
        }
        if (len_base[T] != r.n_al_pool) return Status::NotImplemented("build_wire: allele string pool has bytes no allele refers to");
    }
    bool filt_variant = true;
    for (uint32_t k = 0; k < b.cfg.n_fields; k++)
        if (b.cfg.fields[k].kind == GLX_FK_FT || b.cfg.fields[k].kind == GLX_FK_STRING) filt_variant = false;
    h.n_variant = n_variant;
    h.n_maxend_exc = exc.size() / 2;
    h.flags = (len16 ? GLX_WF_LEN16 : 0) | (gt8 ? GLX_WF_GT8 : 0) | (collen8 ? GLX_WF_COLLEN8 : 0) | (pool16 ? GLX_WF_POOL16 : 0) |
              (filt_variant ? GLX_WF_FILT_VARIANT : 0) | (allen16 ? GLX_WF_ALLEN16 : 0) | (allen8 ? GLX_WF_ALLEN8 : 0) | (filt8 ? GLX_WF_FILT8 : 0);
    // ---- record shapes: (flags, n_allele, GT pair, column lengths) as one byte per record (GLX_WF_SHAPES) ----
    // A row is 4 + n_cols <= 16 bytes (more columns: no shapes), handled as two 64-bit words. Runs of equal shapes (consecutive
    // reference bands) cost one comparison; the rest goes through a small open-addressing table per thread.
    const size_t row_bytes = 4 + C;
    std::vector<uint8_t> shape_code, shape_table, shape_exc;
    std::vector<uint32_t> shape_exc_idx;
    const bool shapes = gt8 && collen8 && R > 0 && row_bytes <= 16;
    struct Key {
        uint64_t a, b;
        bool operator==(const Key& o) const { return a == o.a && b == o.b; }
    };
    auto shape_key = [&](uint64_t i) -> Key {
        uint8_t row[16] = {0};
        row[0] = r.flags[i];
        row[1] = r.n_allele[i];
        for (int hlf = 0; hlf < 2; hlf++) {
            const int16_t g = (int16_t)(r.gt[i] >> (16 * hlf));
            row[2 + hlf] = g == GLX_GT16_VECTOR_END ? (uint8_t)0xFF : (uint8_t)(g >> 1);
        }
        for (uint64_t c = 0; c < C; c++) {
            const uint16_t cl = r.col_len[c * R + i];
            row[4 + c] = cl >= 0xFFF0 ? (uint8_t)(cl & 0xFF) : (uint8_t)cl;
        }
        Key k;
        memcpy(&k.a, row, 8);
        memcpy(&k.b, row + 8, 8);
        return k;
    };
    struct Table {  // key -> value (a count, or a code); linear probing, grows at half full
        std::vector<Key> keys;
        std::vector<uint64_t> vals;
        std::vector<char> used;
        size_t n = 0;
        Table() : keys(256), vals(256, 0), used(256, 0) {}
        static size_t hash(const Key& k) {
            uint64_t h = (k.a * 0x9e3779b97f4a7c15ull) ^ (k.b * 0xc2b2ae3d27d4eb4full);
            return (size_t)(h ^ (h >> 29));
        }
        uint64_t* find(const Key& k, bool insert) {
            size_t m = keys.size() - 1, i = hash(k) & m;
            while (used[i]) {
                if (keys[i] == k) return &vals[i];
                i = (i + 1) & m;
            }
            if (!insert) return nullptr;
            if (2 * (n + 1) > keys.size()) {
                Table bigger;
                bigger.keys.assign(keys.size() * 2, Key{0, 0});
                bigger.vals.assign(keys.size() * 2, 0);
                bigger.used.assign(keys.size() * 2, 0);
                for (size_t j = 0; j < keys.size(); j++)
                    if (used[j]) *bigger.find(keys[j], true) = vals[j];
                *this = std::move(bigger);
                return find(k, true);
            }
            used[i] = 1;
            keys[i] = k;
            vals[i] = 0;
            n++;
            return &vals[i];
        }
    };
    if (shapes) {
        // pass 1: how often each shape occurs
        std::vector<Table> counts(T);
        parallel_chunks(R, T, [&](unsigned t, uint64_t lo, uint64_t hi) {
            Key prev{0, 0};
            uint64_t* slot = nullptr;
            uint64_t run = 0;
            for (uint64_t i = lo; i < hi; i++) {
                const Key cur = shape_key(i);
                if (!slot || !(cur == prev)) {
                    if (slot) *slot += run;
                    run = 0;
                    slot = counts[t].find(cur, true);  // (stays valid: the table only grows inside find, and slot is re-fetched here)
                    prev = cur;
                }
                run++;
            }
            if (slot) *slot += run;
        });
        Table all;
        for (auto& m : counts)
            for (size_t j = 0; j < m.keys.size(); j++)
                if (m.used[j]) *all.find(m.keys[j], true) += m.vals[j];
        std::vector<std::pair<uint64_t, Key>> order;
        order.reserve(all.n);
        for (size_t j = 0; j < all.keys.size(); j++)
            if (all.used[j]) order.emplace_back(all.vals[j], all.keys[j]);
        std::sort(order.begin(), order.end(), [](const auto& x, const auto& y) {
            if (x.first != y.first) return x.first > y.first;
            return x.second.a != y.second.a ? x.second.a < y.second.a : x.second.b < y.second.b;
        });
        const size_t n_shapes = std::min<size_t>(order.size(), g_shape_cap.load());
        Table code_of;
        shape_table.resize(n_shapes * row_bytes);
        for (size_t k = 0; k < order.size(); k++) {
            *code_of.find(order[k].second, true) = k < n_shapes ? k : 0xFF;
            if (k < n_shapes) {
                uint8_t row[16];
                memcpy(row, &order[k].second.a, 8);
                memcpy(row + 8, &order[k].second.b, 8);
                memcpy(&shape_table[k * row_bytes], row, row_bytes);
            }
        }
        // pass 2: the codes; records of rarer shapes keep their row (in record order)
        shape_code.resize(R);
        std::vector<std::vector<uint32_t>> exc_idx(T);
        std::vector<std::vector<uint8_t>> exc_rows(T);
        parallel_chunks(R, T, [&](unsigned t, uint64_t lo, uint64_t hi) {
            Key prev{0, 0};
            int code = -2;
            for (uint64_t i = lo; i < hi; i++) {
                const Key cur = shape_key(i);
                if (code == -2 || !(cur == prev)) {
                    code = (int)*code_of.find(cur, false);
                    prev = cur;
                }
                shape_code[i] = (uint8_t)code;
                if (code == 0xFF) {
                    uint8_t row[16];
                    memcpy(row, &cur.a, 8);
                    memcpy(row + 8, &cur.b, 8);
                    exc_idx[t].push_back((uint32_t)i);
                    exc_rows[t].insert(exc_rows[t].end(), row, row + row_bytes);
                }
            }
        });
        for (unsigned t = 0; t < T; t++) {  // (chunks are ascending record ranges: the concatenation stays sorted)
            shape_exc_idx.insert(shape_exc_idx.end(), exc_idx[t].begin(), exc_idx[t].end());
            shape_exc.insert(shape_exc.end(), exc_rows[t].begin(), exc_rows[t].end());
        }
        h.flags |= GLX_WF_SHAPES;
        h.n_shapes = (uint32_t)n_shapes;
        h.n_shape_exc = shape_exc_idx.size();
    }
    // ---- layout ----
    size_t off = align16(sizeof(glx_wire));
    std::vector<std::pair<size_t, size_t>> pads;  // [end of a section's bytes, start of the next): zeroed by hand, the buffer comes uninitialised
    pads.emplace_back(sizeof(glx_wire), off);
    auto place = [&](int sec, size_t bytes) {
        h.off[sec] = off;
        pads.emplace_back(off + bytes, align16(off + bytes));
        off = align16(off + bytes);
    };
    size_t colval_bytes = 0;
    for (uint64_t c = 0; c < C; c++) colval_bytes += align16(R * h.col_w[c]);
    place(GLX_WS_DS_REC_OFF, (D + 1) * 8);
    place(GLX_WS_BEG, R * 4);
    place(GLX_WS_LEN, R * (len16 ? 2 : 4));
    place(GLX_WS_FLAGS, shapes ? 0 : R);
    place(GLX_WS_NALLELE, shapes ? 0 : R);
    place(GLX_WS_GT, shapes ? 0 : R * (gt8 ? 2 : 4));
    place(GLX_WS_QUAL, n_variant * 4);
    place(GLX_WS_FILT, (filt_variant ? n_variant : R) * (filt8 ? 1 : 4));
    place(GLX_WS_COL_LEN, shapes ? 0 : C * R * (collen8 ? 1 : 2));
    place(GLX_WS_COL_VAL, colval_bytes);
    place(GLX_WS_POOL, r.n_pool * (pool16 ? 2 : 4));
    place(GLX_WS_AL_LEN, r.n_al * (allen8 ? 1 : (allen16 ? 2 : 4)));
    place(GLX_WS_AL_POOL, r.n_al_pool);
    place(GLX_WS_MAXEND_EXC, exc.size() * 4);
    place(GLX_WS_SHAPE, shape_code.size());
    place(GLX_WS_SHAPE_TABLE, shape_table.size());
    place(GLX_WS_SHAPE_EXC_IDX, shape_exc_idx.size() * 4);
    place(GLX_WS_SHAPE_EXC, shape_exc.size());
    h.n_bytes = off;
    w.clear();
    w.resize(off);  // uninitialised (WireBytes): every page is first touched by the thread that fills it
    memcpy(w.data(), &h, sizeof h);
    uint8_t* p = w.data();
    for (const auto& g : pads) memset(p + g.first, 0, g.second - g.first);
    {  // the scalar columns sit back to back, each padded to 16 bytes
        size_t co = h.off[GLX_WS_COL_VAL];
        for (uint64_t c = 0; c < C; c++) {
            memset(p + co + R * h.col_w[c], 0, align16(R * h.col_w[c]) - R * h.col_w[c]);
            co += align16(R * h.col_w[c]);
        }
    }
    memcpy(p + h.off[GLX_WS_DS_REC_OFF], r.ds_rec_off, (D + 1) * 8);
    if (shapes) {
        memcpy(p + h.off[GLX_WS_SHAPE_TABLE], shape_table.data(), shape_table.size());
        if (!shape_exc_idx.empty()) {
            memcpy(p + h.off[GLX_WS_SHAPE_EXC_IDX], shape_exc_idx.data(), shape_exc_idx.size() * 4);
            memcpy(p + h.off[GLX_WS_SHAPE_EXC], shape_exc.data(), shape_exc.size());
        }
    }
    {
        uint16_t* l16 = reinterpret_cast<uint16_t*>(p + h.off[GLX_WS_LEN]);
        uint32_t* l32 = reinterpret_cast<uint32_t*>(p + h.off[GLX_WS_LEN]);
        uint8_t* g8 = p + h.off[GLX_WS_GT];
        uint32_t* g32 = reinterpret_cast<uint32_t*>(p + h.off[GLX_WS_GT]);
        float* q = reinterpret_cast<float*>(p + h.off[GLX_WS_QUAL]);
        uint32_t* f = reinterpret_cast<uint32_t*>(p + h.off[GLX_WS_FILT]);
        uint8_t* f8 = p + h.off[GLX_WS_FILT];
        uint8_t* c8 = p + h.off[GLX_WS_COL_LEN];
        uint16_t* c16 = reinterpret_cast<uint16_t*>(p + h.off[GLX_WS_COL_LEN]);
        std::vector<size_t> col_off(C + 1, h.off[GLX_WS_COL_VAL]);
        for (uint64_t c = 0; c < C; c++) col_off[c + 1] = col_off[c] + align16(R * h.col_w[c]);
        parallel_chunks(R, T, [&](unsigned t, uint64_t lo, uint64_t hi) {
            uint64_t v = var_base[t];
            memcpy(p + h.off[GLX_WS_BEG] + lo * 4, r.beg + lo, (hi - lo) * 4);
            if (shapes) memcpy(p + h.off[GLX_WS_SHAPE] + lo, shape_code.data() + lo, hi - lo);
            else {
                memcpy(p + h.off[GLX_WS_FLAGS] + lo, r.flags + lo, hi - lo);
                memcpy(p + h.off[GLX_WS_NALLELE] + lo, r.n_allele + lo, hi - lo);
            }
            for (uint64_t i = lo; i < hi; i++) {
                const uint32_t len = (uint32_t)(r.end[i] - r.beg[i]);
                if (len16) l16[i] = (uint16_t)len;
                else l32[i] = len;
                if (shapes) {
                } else if (gt8) {
                    for (int hlf = 0; hlf < 2; hlf++) {
                        const int16_t g = (int16_t)(r.gt[i] >> (16 * hlf));
                        g8[2 * i + hlf] = g == GLX_GT16_VECTOR_END ? (uint8_t)0xFF : (uint8_t)(g >> 1);
                    }
                } else
                    g32[i] = r.gt[i];
                if (!(r.flags[i] & GLX_RF_IS_REF)) {
                    q[v] = r.qual[i];
                    if (filt_variant) {
                        if (filt8) f8[v] = (uint8_t)r.filt[i];
                        else f[v] = r.filt[i];
                    }
                    v++;
                }
                if (!filt_variant) {
                    if (filt8) f8[i] = (uint8_t)r.filt[i];
                    else f[i] = r.filt[i];
                }
            }
            for (uint64_t c = 0; c < C; c++) {
                int16_t* d16 = reinterpret_cast<int16_t*>(p + col_off[c]);
                int32_t* d32 = reinterpret_cast<int32_t*>(p + col_off[c]);
                for (uint64_t i = lo; i < hi; i++) {
                    const uint16_t cl = r.col_len[c * R + i];
                    if (shapes) {
                    } else if (collen8) c8[c * R + i] = cl >= 0xFFF0 ? (uint8_t)(cl & 0xFF) : (uint8_t)cl;  // 0xFFFF -> 0xFF, 0xFFFE -> 0xFE, 0xFFFD -> 0xFD
                    else c16[c * R + i] = cl;
                    if (h.col_w[c] == 2) d16[i] = cl == 1 ? to16(r.col_val[c * R + i]) : (int16_t)0;
                    else if (h.col_w[c] == 4) d32[i] = cl == 1 ? r.col_val[c * R + i] : 0;
                }
            }
        });
    }
    if (pool16) {
        int16_t* d = reinterpret_cast<int16_t*>(p + h.off[GLX_WS_POOL]);
        parallel_chunks(r.n_pool, T, [&](unsigned, uint64_t lo, uint64_t hi) {
            for (uint64_t i = lo; i < hi; i++) d[i] = to16(r.pool[i]);
        });
    } else
        memcpy(p + h.off[GLX_WS_POOL], r.pool, r.n_pool * 4);
    if (allen8) {
        uint8_t* d = p + h.off[GLX_WS_AL_LEN];
        parallel_chunks(r.n_al, T, [&](unsigned, uint64_t lo, uint64_t hi) {
            for (uint64_t a = lo; a < hi; a++) d[a] = (uint8_t)r.al_len[a];
        });
    } else if (allen16) {
        uint16_t* d = reinterpret_cast<uint16_t*>(p + h.off[GLX_WS_AL_LEN]);
        parallel_chunks(r.n_al, T, [&](unsigned, uint64_t lo, uint64_t hi) {
            for (uint64_t a = lo; a < hi; a++) d[a] = (uint16_t)r.al_len[a];
        });
    } else
        memcpy(p + h.off[GLX_WS_AL_LEN], r.al_len, r.n_al * 4);
    memcpy(p + h.off[GLX_WS_AL_POOL], r.al_pool, r.n_al_pool);
    if (!exc.empty()) memcpy(p + h.off[GLX_WS_MAXEND_EXC], exc.data(), exc.size() * 4);
    return Status::OK();
}

}  // namespace glx
