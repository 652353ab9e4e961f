// glx_emu.cc — TEST INFRASTRUCTURE: the device cell logic compiled for the host (see cuda_host_shim.h) and driven
// cell by cell over host arrays. Used by tests/test_emu_parity.py to diff the kernels' control flow against the
// oracle without a GPU. Not part of libglx.so.
#include "cuda_host_shim.h"

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../../glnexus_b200/csrc/glx_cell_fast.cuh"
#include "../../glnexus_b200/csrc/glx_cell_general.cuh"
#include "../../glnexus_b200/csrc/glx_cell_variant.cuh"
#include "../../glnexus_b200/csrc/glx_tiled_sig.cuh"
#include "../../glnexus_b200/csrc/glx_bcf.cuh"

using namespace glxd;

static void bind(const glx_config* cfg, const glx_records* r, const glx_sites* s, glx_out* o, unsigned long long* err, DCfg& C, DRecs& R, DSites& S, DOut& O) {
    C = make_dcfg(*cfg);
    memset(&R, 0, sizeof R);
    R.n_records = r->n_records;
    R.ds_rec_off = r->ds_rec_off; R.beg = r->beg; R.end = r->end; R.maxend = r->maxend; R.qual = r->qual;
    R.n_allele = r->n_allele; R.flags = r->flags; R.filt = r->filt; R.gt = r->gt; R.allele_off = r->allele_off;
    R.col_val = r->col_val; R.col_len = r->col_len; R.pool = r->pool;
    R.al_len = r->al_len; R.al_flags = r->al_flags; R.al_prefix = r->al_prefix; R.al_str = r->al_str; R.al_pool = r->al_pool;
    memset(&S, 0, sizeof S);
    S.n_sites = s->n_sites; S.n_samples = s->n_samples;
    S.beg = s->beg; S.end = s->end; S.qbeg = s->qbeg; S.qend = s->qend; S.n_allele = s->n_allele; S.monoallelic = s->monoallelic;
    S.allele_off = s->allele_off; S.allele_is_indel = s->allele_is_indel; S.hom_log_prior = s->hom_log_prior; S.lost_log_prior = s->lost_log_prior;
    S.unif_off = s->unif_off; S.unif_beg = s->unif_beg; S.unif_end = s->unif_end; S.unif_to = s->unif_to; S.unif_len = s->unif_len;
    S.unif_prefix = s->unif_prefix; S.unif_str = s->unif_str; S.unif_pool = s->unif_pool;
    for (int k = 0; k < GLX_MAX_FIELDS; k++) { S.fld_cnt[k] = s->fld_cnt[k]; S.fld_off[k] = s->fld_off[k]; }
    static std::vector<int4> site_rows, unif_rows;  // as glx_upload packs them
    build_site_rows(*s, site_rows, unif_rows);
    S.site_rows = site_rows.data();
    S.unif_rows = unif_rows.data();
    memset(&O, 0, sizeof O);
    O.gt = o->gt; O.rnc = o->rnc; O.allele_used = o->allele_used; O.site_flags = o->site_flags; O.error = err;
    for (int k = 0; k < GLX_MAX_FIELDS; k++) O.fld[k] = o->fld[k];
}

// multi-sample (or reordered) datasets: the loop of glx_cells_multi_kernel, the validation of glx_upload_on
static int emu_multi_batch(const glx_config* cfg, const glx_records* recs, const glx_sites* sites, glx_out* out) {
    const uint32_t Sn = sites->n_sites, N = sites->n_samples;
    if (recs->n_samples_out != N) return GLX_E_UNSUPPORTED;
    std::vector<uint2> out_map(N, uint2{0xFFFFFFFFu, 0u});
    for (uint32_t d = 0; d < recs->n_datasets; d++) {
        const uint32_t ns = recs->ds_n_sample[d];
        if (ns == 0 || ns > 0xFFFFu) return GLX_E_INVALID;
        for (uint32_t k = 0; k < ns; k++) {
            const int32_t o = recs->sample_out[recs->ds_sample_off[d] + k];
            if (o < 0 || (uint32_t)o >= N || out_map[o].x != 0xFFFFFFFFu) return GLX_E_UNSUPPORTED;
            out_map[o] = uint2{d, k | ns << 16};
        }
    }
    DCfg C; DRecs R0; DSites S; DOut O;
    unsigned long long err = ~0ull;
    bind(cfg, recs, sites, out, &err, C, R0, S, O);
    std::vector<uint8_t> flags(Sn + 8, 0);
    O.site_flags = flags.data();
    uint32_t max_slots = 1;
    for (uint32_t s = 0; s < Sn; s++) {
        uint32_t t = 0;
        for (uint32_t k = 0; k < cfg->n_fields; k++) t += sites->fld_cnt[k][s];
        max_slots = std::max(max_slots, t);
        out->allele_used[s] = 0;
    }
    std::vector<uint8_t> cnt(max_slots);
    MRecs R;
    static_cast<DRecs&>(R) = R0;
    R.gt_pool = recs->gt_pool;
    for (uint32_t s = 0; s < Sn; s++)
        for (uint32_t n = 0; n < N; n++) {
            const uint2 m = out_map[n];
            uint64_t lo = 0, hi = 0;
            R.ns = 1;
            R.si = 0;
            uint32_t ds = n;
            if (m.x != 0xFFFFFFFFu) {
                ds = m.x;
                lo = R.ds_rec_off[m.x];
                hi = R.ds_rec_off[m.x + 1];
                R.si = m.y & 0xFFFFu;
                R.ns = m.y >> 16;
            }
            const unsigned long long err_cell = (unsigned long long)s * N + (ds < N ? ds : N - 1);
            genotype_cell_general(C, R, S, O, s, n, lo < hi ? first_candidate(R.maxend, lo, hi, S.qbeg[s]) : lo, hi, cnt.data(), 1, err_cell);
        }
    memcpy(out->site_flags, flags.data(), Sn);
    if (err == ~0ull) return GLX_OK;
    const int code = (int)(err & 0xff);
    return code == 101 ? GLX_E_UNSUPPORTED : -code;
}

extern "C" int glx_emu_general_batch(const glx_config* cfg, const glx_records* recs, const glx_sites* sites, glx_out* out) {
    bool multi = recs->n_datasets != sites->n_samples;
    for (uint32_t d = 0; !multi && d < recs->n_datasets; d++)
        if (recs->ds_n_sample[d] != 1 || recs->sample_out[recs->ds_sample_off[d]] != (int32_t)d) multi = true;
    if (multi) return emu_multi_batch(cfg, recs, sites, out);
    DCfg C; DRecs R; DSites S; DOut O;
    unsigned long long err = ~0ull;
    bind(cfg, recs, sites, out, &err, C, R, S, O);
    const uint32_t Sn = sites->n_sites, N = sites->n_samples;
    // site_flags is OR-ed through an aligned 32-bit word: give it padded storage
    std::vector<uint8_t> flags(Sn + 8, 0);
    O.site_flags = flags.data();
    uint32_t max_slots = 1;
    for (uint32_t s = 0; s < Sn; s++) {
        uint32_t t = 0;
        for (uint32_t k = 0; k < cfg->n_fields; k++) t += sites->fld_cnt[k][s];
        max_slots = std::max(max_slots, t);
        out->allele_used[s] = 0;
    }
    std::vector<uint8_t> cnt(max_slots);
    for (uint32_t s = 0; s < Sn; s++)
        for (uint32_t n = 0; n < N; n++) {
            const uint64_t lo = R.ds_rec_off[n], hi = R.ds_rec_off[n + 1];
            genotype_cell_general(C, R, S, O, s, n, first_candidate(R.maxend, lo, hi, S.qbeg[s]), hi, cnt.data(), 1);
        }
    memcpy(out->site_flags, flags.data(), Sn);
    if (err == ~0ull) return GLX_OK;
    const int code = (int)(err & 0xff);
    return code == 101 ? GLX_E_UNSUPPORTED : -code;
}

// composition of the cells the fast path declines: [0] finished by the single-record closed form, [1] general path,
// [2] finished by the all-bands closed form, [3] all-band cells that nevertheless went to the general path
static int32_t vsm[64];  // the shared-memory column of genotype_cell_variant, stride 1
static uint64_t g_emu_stats[9];  // [8] of [4]: finished by the biallelic form; [0] single-record closed form [1] general [2] all-band closed form [3] bands declined [4] of [0]: through the record view; [5..7] see note_general
extern "C" void glx_emu_stats(uint64_t out[8]) { memcpy(out, g_emu_stats, 8 * sizeof(uint64_t)); }
extern "C" uint64_t glx_emu_biallelic_cells() { return g_emu_stats[8]; }
static void note_general(const DRecs& R, const DSites& S, uint32_t s, uint64_t first, uint64_t hi) {
    g_emu_stats[1]++;
    int n = 0;
    bool all_ref = true;
    for (uint64_t r = first; r < hi && R.beg[r] < S.qend[s]; r++) {
        if (!(R.end[r] > S.qbeg[s])) continue;
        n++;
        if (!(R.flags[r] & GLX_RF_IS_REF)) all_ref = false;
    }
    if (all_ref && n >= 2) g_emu_stats[3]++;  // all-band cells the bands form declined
    if (S.monoallelic[s]) g_emu_stats[5]++;   // [5] cells of monoallelic sites
    else if (!all_ref) {
        int nv = 0;
        for (uint64_t r = first; r < hi && R.beg[r] < S.qend[s]; r++)
            if (R.end[r] > S.qbeg[s] && !(R.flags[r] & GLX_RF_IS_REF)) nv++;
        if (nv == 1) g_emu_stats[6]++;        // [6] bands + exactly one variant record
        else g_emu_stats[7]++;                // [7] several variant records
    }
}

// signature-specialised lanes (glx_tiled_sig.cuh), lane by lane
template <class Sig>
static void run_tiled_sig(const DCfg& C, const FastPlan& plan, const DRecs& R, const DSites& S, const DOut& O, const int32_t* blobs, const uint32_t* cursors,
                          std::vector<uint8_t>& flags, std::vector<uint8_t>& cnt, uint64_t& n_slow) {
    const uint32_t Sn = S.n_sites, N = S.n_samples;
    for (uint32_t n = 0; n < N; n++) {
        for (uint32_t site0 = 0; site0 < Sn; site0 += GLX_TILE_SITES) {
            const uint32_t nsite = std::min<uint32_t>(GLX_TILE_SITES, Sn - site0);
            TiledLane<Sig> L;
            L.init(R, blobs, n, cursors[(size_t)(site0 / GLX_TILE_SITES) * N + n]);
            for (uint32_t i = 0; i < nsite; i++) {
                const uint32_t s = site0 + i;
                int rnc = RNC_M;
                bool called = false;
                const int type = L.classify(C, blobs, S.beg[s], S.end[s], S.qbeg[s], S.qend[s], S.monoallelic[s] != 0, i > 0 && S.qbeg[s] < S.qbeg[s - 1], called, rnc);
                if (type == CT_SLOW) {
                    n_slow++;
                    const bool single = !S.monoallelic[s] && L.c < L.n && L.beg_c < S.qend[s] && L.beg_n >= S.qend[s];
                    if (single && genotype_cell_biallelic<Sig>(C, plan, R, S, O, blobs, s, n, R.ds_rec_off[n] + L.c)) {  // the fused kernel's form
                        g_emu_stats[0]++;
                        g_emu_stats[4]++;
                        g_emu_stats[8]++;
                        continue;
                    }
                    if (single && genotype_cell_variant<8, Sig>(C, plan, R, S, O, blobs, s, n, R.ds_rec_off[n] + L.c, vsm, 1)) {
                        g_emu_stats[0]++;
                        g_emu_stats[4]++;
                        continue;
                    }
                    if (single && genotype_cell_single(C, R, S, O, s, n, R.ds_rec_off[n] + L.c)) {
                        g_emu_stats[0]++;
                        continue;
                    }
                    if (!single && genotype_cell_bands(C, plan, R, S, O, blobs, s, n, R.ds_rec_off[n] + L.c, R.ds_rec_off[n + 1])) {
                        g_emu_stats[2]++;
                        continue;
                    }
                    note_general(R, S, s, R.ds_rec_off[n] + L.c, R.ds_rec_off[n + 1]);
                    genotype_cell_general(C, R, S, O, s, n, R.ds_rec_off[n] + L.c, R.ds_rec_off[n + 1], cnt.data(), 1);
                    continue;
                }
                const size_t c2 = ((size_t)s * N + n) * 2;
                O.gt[c2] = O.gt[c2 + 1] = called ? 0 : -1;
                O.rnc[c2] = O.rnc[c2 + 1] = (uint8_t)kRncChar[rnc];
                if (called) O.allele_used[s] |= 1u;
                if (rnc == RNC_I) flags[s] |= 1;
                L.emit(plan, O, true, type == CT_REF, s, N, n, S.n_allele[s], [&](uint32_t k) { return (unsigned long long)S.fld_off[k][s]; }, nullptr, 0, 1);
            }
        }
    }
}

// The tiled kernel's per-lane state machine run lane by lane: same tiles, same cursor walk, same classify /
// resolve / value functions; cells it declines go through the general cell. (The warp-transposed stores and
// the worklist hand-off are CUDA-only and are covered by the GPU tests.)
extern "C" void glx_emu_stats_reset() { memset(g_emu_stats, 0, sizeof g_emu_stats); }
extern "C" int glx_emu_tiled_batch(const glx_config* cfg, const glx_records* recs, const glx_sites* sites, glx_out* out, uint64_t* n_slow_out, int dynamic_only) {
    {  // a batch with multi-sample datasets has no tiled path: the device runs glx_cells_multi_kernel whatever the options
        bool multi = recs->n_datasets != sites->n_samples;
        for (uint32_t d = 0; !multi && d < recs->n_datasets; d++)
            if (recs->ds_n_sample[d] != 1 || recs->sample_out[recs->ds_sample_off[d]] != (int32_t)d) multi = true;
        if (multi) {
            if (n_slow_out) *n_slow_out = 0;
            return emu_multi_batch(cfg, recs, sites, out);
        }
    }
    DCfg C; DRecs R; DSites S; DOut O;
    unsigned long long err = ~0ull;
    bind(cfg, recs, sites, out, &err, C, R, S, O);
    const FastPlan plan = make_fast_plan(C);
    const uint32_t Sn = sites->n_sites, N = sites->n_samples;
    std::vector<uint8_t> flags(Sn + 8, 0);
    O.site_flags = flags.data();
    uint32_t max_slots = 1;
    for (uint32_t s = 0; s < Sn; s++) {
        uint32_t t = 0;
        for (uint32_t k = 0; k < cfg->n_fields; k++) t += sites->fld_cnt[k][s];
        max_slots = std::max(max_slots, t);
        out->allele_used[s] = 0;
    }
    std::vector<uint8_t> cnt(max_slots);
    uint64_t n_slow = 0;
    // glx_resolve_records_kernel: one blob per record
    const uint32_t n_rv = plan.n_rv, W = blob_words(n_rv);
    std::vector<int32_t> blobs((size_t)(R.n_records + 1) * W);
    if (plan.enabled)
        for (uint32_t d = 0; d < N; d++)
            for (uint64_t r = R.ds_rec_off[d]; r < R.ds_rec_off[d + 1]; r++) build_record_blob(C, plan, R, r, r + 1 == R.ds_rec_off[d + 1], blobs.data() + r * W);
    // tile cursor table, as glx_upload + glx_resolve_records_kernel build it
    const uint32_t n_tiles = (Sn + GLX_TILE_SITES - 1) / GLX_TILE_SITES;
    std::vector<int32_t> tq(n_tiles + 1);
    {
        int32_t run = INT32_MAX;
        for (uint32_t tile = n_tiles; tile-- > 0;) {
            run = std::min(run, S.qbeg[(size_t)tile * GLX_TILE_SITES]);
            tq[tile] = run;
        }
    }
    std::vector<uint32_t> cursors((size_t)n_tiles * N + 1, 0xffffffffu);
    for (uint32_t d = 0; d < N; d++)
        for (uint64_t r = R.ds_rec_off[d]; r < R.ds_rec_off[d + 1]; r++) fill_tile_cursors(R, r, R.ds_rec_off[d], d, N, tq.data(), n_tiles, cursors.data());
    {
        uint32_t nf;
        uint64_t modes, nums, nrvs;
        bool handled = false;
        if (!dynamic_only && sig_of_plan(C, plan, nf, modes, nums, nrvs)) {  // same dispatch as glx_upload
            auto match = [&](uint32_t NF, uint64_t M, uint64_t Nu, uint64_t Nr) { return nf == NF && modes == M && nums == Nu && nrvs == Nr; };
            // glx_resolve_sig_kernel builds its rows with the signature's unrolled builder: must be the same rows
            auto check_rows = [&](auto sig) -> bool {
                using Sig = decltype(sig);
                if (!plan.vview) return true;
                for (uint32_t d = 0; d < N; d++)
                    for (uint64_t r = R.ds_rec_off[d]; r < R.ds_rec_off[d + 1]; r++) {
                        int32_t w[Sig::W];
                        build_record_blob_sig<Sig>(C, plan, R, r, r + 1 == R.ds_rec_off[d + 1], w);
                        if (memcmp(w, blobs.data() + r * W, sizeof w) != 0) return false;
                    }
                return true;
            };
            if (match(4, 0x1010ull, 0x2030ull, 0x4141ull) ? !check_rows(SigDeepVariant()) : (match(5, 0x20010ull, 0x20030ull, 0x01441ull) ? !check_rows(SigGatk()) : false))
                return GLX_E_INVALID - 1000;  // signature builder disagrees with build_record_blob
            if (match(4, 0x1010ull, 0x2030ull, 0x4141ull)) {
                run_tiled_sig<SigDeepVariant>(C, plan, R, S, O, blobs.data(), cursors.data(), flags, cnt, n_slow);
                handled = true;
            } else if (match(5, 0x20010ull, 0x20030ull, 0x01441ull)) {
                run_tiled_sig<SigGatk>(C, plan, R, S, O, blobs.data(), cursors.data(), flags, cnt, n_slow);
                handled = true;
            }
        }
        if (handled) {
            memcpy(out->site_flags, flags.data(), Sn);
            if (n_slow_out) *n_slow_out = n_slow;
            if (err == ~0ull) return GLX_OK;
            const int code = (int)(err & 0xff);
            return code == 101 ? GLX_E_UNSUPPORTED : -code;
        }
    }
    for (uint32_t n = 0; n < N; n++) {
        for (uint32_t site0 = 0; site0 < Sn; site0 += GLX_TILE_SITES) {
            const uint32_t nsite = std::min<uint32_t>(GLX_TILE_SITES, Sn - site0);
            LaneCursor L;
            int32_t rv[GLX_FAST_RV];
            if (plan.enabled) cursor_init(R, blobs.data(), W, n_rv, L, n, cursors[(size_t)(site0 / GLX_TILE_SITES) * N + n], rv, 1);
            for (uint32_t i = 0; i < nsite; i++) {
                const uint32_t s = site0 + i;
                int type = CT_SLOW, rnc = RNC_M;
                bool called = false;
                if (plan.enabled)
                    type = fast_classify(C, blobs.data(), W, n_rv, L, S.beg[s], S.end[s], S.qbeg[s], S.qend[s], S.monoallelic[s] != 0,
                                         i > 0 && S.qbeg[s] < S.qbeg[s - 1], rv, 1, called, rnc);
                if (type == CT_SLOW) {
                    n_slow++;
                    const uint64_t lo = R.ds_rec_off[n], hi = R.ds_rec_off[n + 1];
                    // the kernel passes the lane's cursor instead of searching again
                    const uint64_t first = plan.enabled ? lo + L.c : first_candidate(R.maxend, lo, hi, S.qbeg[s]);
                    const bool single = plan.enabled && !S.monoallelic[s] && L.c < L.n && L.beg_c < S.qend[s] && L.beg_n >= S.qend[s];
                    if (single && genotype_cell_variant<4>(C, plan, R, S, O, blobs.data(), s, n, first, vsm, 1)) continue;
                    if (single && genotype_cell_single(C, R, S, O, s, n, first)) continue;
                    if (!single && plan.enabled && genotype_cell_bands(C, plan, R, S, O, blobs.data(), s, n, first, hi)) continue;
                    genotype_cell_general(C, R, S, O, s, n, first, hi, cnt.data(), 1);
                    continue;
                }
                const size_t c2 = ((size_t)s * N + n) * 2;
                O.gt[c2] = O.gt[c2 + 1] = called ? 0 : -1;
                O.rnc[c2] = O.rnc[c2 + 1] = (uint8_t)kRncChar[rnc];
                if (called) O.allele_used[s] |= 1u;
                if (rnc == RNC_I) flags[s] |= 1;
                for (uint32_t k = 0; k < C.n_fields; k++) {
                    const int count = S.fld_cnt[k][s];
                    fast_emit_field(plan, k, type == CT_REF, count, O.fld[k] + S.fld_off[k][s] + (uint64_t)n * count, rv, 1);
                }
            }
        }
    }
    memcpy(out->site_flags, flags.data(), Sn);
    if (n_slow_out) *n_slow_out = n_slow;
    if (err == ~0ull) return GLX_OK;
    const int code = (int)(err & 0xff);
    return code == 101 ? GLX_E_UNSUPPORTED : -code;
}

// The resolve kernels confine their per-record searches with per-block hints (glx_block_hints_kernel) and take t_lo from
// the neighbouring record. Replay that arithmetic for every record — block size `block`, tiles of `tile_sites` sites —
// and compare the claimed tile range with fill_tile_cursors' definition (two full searches). Returns the mismatches.
static uint32_t lower_tile(const int32_t* tq, uint32_t lo, uint32_t hi, int v) {
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tq[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
extern "C" int glx_emu_check_hints(const glx_records* recs, const glx_sites* sites, uint32_t tile_sites, uint32_t block) {
    const uint64_t Rn = recs->n_records;
    const uint32_t D = recs->n_datasets, Sn = sites->n_sites;
    if (!Rn || !Sn) return 0;
    const uint32_t n_tiles = (Sn + tile_sites - 1) / tile_sites;
    std::vector<int32_t> tq(n_tiles);
    int32_t run = INT32_MAX;
    for (uint32_t tile = n_tiles; tile-- > 0;) {
        run = std::min(run, sites->qbeg[(size_t)tile * tile_sites]);
        tq[tile] = run;
    }
    auto sample_of = [&](uint64_t r, uint32_t lo, uint32_t hi) {
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (recs->ds_rec_off[mid] <= r) lo = mid;
            else hi = mid;
        }
        return lo;
    };
    const uint64_t n_blocks = (Rn + block - 1) / block;
    std::vector<uint32_t> hs(n_blocks + 1), ht(n_blocks + 1);
    for (uint64_t b = 0; b < n_blocks; b++) {
        hs[b] = sample_of(b * block, 0, D);
        ht[b] = lower_tile(tq.data(), 0, n_tiles, recs->maxend[b * block]);
    }
    hs[n_blocks] = D - 1;
    ht[n_blocks] = n_tiles;
    int bad = 0;
    std::vector<uint32_t> thi(Rn);
    for (uint64_t r = 0; r < Rn; r++) {
        const uint64_t b = r / block;
        const uint32_t sample = sample_of(r, hs[b], hs[b + 1] + 1);
        if (!(recs->ds_rec_off[sample] <= r && r < recs->ds_rec_off[sample + 1])) bad++;
        const bool one = hs[b] == hs[b + 1];
        thi[r] = lower_tile(tq.data(), one ? ht[b] : 0u, one ? ht[b + 1] : n_tiles, recs->maxend[r]);
        if (thi[r] != lower_tile(tq.data(), 0, n_tiles, recs->maxend[r])) bad++;
        const uint64_t base = recs->ds_rec_off[sample];
        uint32_t t_lo = 0;
        if (r > base) {
            const bool first_of_warp = (r % 32) == 0, first_of_block = (r % block) == 0;
            t_lo = first_of_warp ? lower_tile(tq.data(), (one && !first_of_block) ? ht[b] : 0u, thi[r], recs->maxend[r - 1]) : thi[r - 1];
            if (t_lo != lower_tile(tq.data(), 0, thi[r], recs->maxend[r - 1])) bad++;
        }
    }
    return bad;
}


// The device BCF2 record encoder (glx_bcf.cuh) run stage by stage with one virtual thread per block.
struct EmuBcf {
    BcfEnc E;
    std::vector<int32_t> mm;
    std::vector<BcfSitePlan> plan;
    std::vector<unsigned long long> roff, hoff;
    std::vector<uint8_t> pool;
    unsigned long long total = 0;
    EmuBcf(const glx_config* cfg, const glx_sites* sites, const glx_bcf_meta* meta, const glx_out* out, uint32_t chunk) {
        const uint32_t S = sites->n_sites, N = sites->n_samples, F = cfg->n_fields;
        memset(&E, 0, sizeof E);
        bcf_enc_scalars(E, *cfg, *meta, S, N);
        E.beg = sites->beg; E.n_allele = sites->n_allele; E.monoallelic = sites->monoallelic; E.allele_off = sites->allele_off;
        for (uint32_t k = 0; k < F; k++) { E.fld_cnt[k] = sites->fld_cnt[k]; E.fld_off[k] = sites->fld_off[k]; E.fld[k] = out->fld[k]; }
        E.gt = out->gt; E.rnc = out->rnc; E.allele_used = out->allele_used;
        E.m.rid = meta->rid; E.m.qual = meta->qual; E.m.dna_off = meta->dna_off; E.m.dna_pool = meta->dna_pool;
        E.m.norm_beg = meta->norm_beg; E.m.norm_end = meta->norm_end; E.m.norm_off = meta->norm_off; E.m.norm_pool = meta->norm_pool;
        E.m.aq = meta->aq; E.m.af = sites->allele_af; E.m.contig_off = meta->contig_off; E.m.contig_pool = meta->contig_pool;
        E.m.filter_off = meta->filter_off; E.m.filter_pool = meta->filter_pool;
        mm.assign((size_t)S * F * 2 + 2, 0);
        for (size_t i = 0; i < (size_t)S * F; i++) { mm[2 * i] = INT32_MAX; mm[2 * i + 1] = INT32_MIN + 1; }
        plan.resize(S + 1);
        roff.assign(S + 1, 0);
        hoff.assign(S + 1, 0);
        E.mm = mm.data(); E.plan = plan.data(); E.rec_off = roff.data(); E.hdr_off = hoff.data();
        if (chunk == 0) chunk = 2048;
        std::vector<int32_t> s_mm(2 * GLX_MAX_FIELDS);
        for (uint32_t s = 0; s < S; s++)
            for (uint32_t n0 = 0; n0 < N; n0 += chunk) {
                bcf_scan_init(E, s_mm.data(), 0, 1);
                bcf_scan_accum(E, s, n0, std::min(N, n0 + chunk), s_mm.data(), 0, 1);
                bcf_scan_flush(E, s, s_mm.data(), 0, 1);
            }
        for (uint32_t s = 0; s < S; s++) bcf_layout_site(E, s);
        unsigned long long ra = 0, ha = 0;
        for (uint32_t s = 0; s <= S; s++) {  // exclusive scans
            const unsigned long long r = s < S ? roff[s] : 0, h = s < S ? hoff[s] : 0;
            roff[s] = ra; hoff[s] = ha;
            ra += r; ha += h;
        }
        pool.resize(ha + 16);
        E.hdr_pool = pool.data();
        for (uint32_t s = 0; s < S; s++) bcf_header_site(E, s);
        total = roff[S];
    }
};

// `span` = bytes per span (the device uses GLX_BCF_SPAN; small values exercise the clipping at span boundaries), `chunk` =
// samples per scan block. *bytes is malloc'd (free()).
#include "../../glnexus_b200/csrc/glx_bgzf.cuh"
extern "C" int glx_emu_bcf_records(const glx_config* cfg, const glx_sites* sites, const glx_bcf_meta* meta, const glx_out* out, uint32_t span, uint32_t chunk,
                                   uint8_t** bytes, uint64_t* n_bytes, uint64_t* rec_off) {
    EmuBcf X(cfg, sites, meta, out, chunk);
    uint8_t* o = (uint8_t*)malloc(X.total ? X.total + 8 : 8);
    memset(o, 0xEE, X.total);
    auto tab = std::make_unique<BcfSpanTab>();
    if (span == 0) span = GLX_BCF_SPAN;
    for (unsigned long long lo = 0; lo < X.total; lo += span)
        bcf_fill_span<BcfLinear>(X.E, lo, std::min<unsigned long long>(X.total, lo + span), bcf_first_site(X.E, lo), o + lo, *tab, 0, 1);
    for (uint32_t s = 0; s <= sites->n_sites; s++) rec_off[s] = X.roff[s];
    *bytes = o;
    *n_bytes = X.total;
    return 0;
}

// The BGZF stage (glx_bgzf.cuh) on top: sample spans -> the batch's Huffman table -> every span becomes one BGZF block; the
// blocks are concatenated. `span` <= GLX_BCF_SPAN.
extern "C" int glx_emu_bgzf(const glx_config* cfg, const glx_sites* sites, const glx_bcf_meta* meta, const glx_out* out, uint32_t span, uint8_t** bytes,
                            uint64_t* n_bytes, uint32_t* n_stored) {
    EmuBcf X(cfg, sites, meta, out, 0);
    if (span == 0 || span > GLX_BCF_SPAN) span = GLX_BCF_SPAN;
    std::vector<uint8_t> all;
    std::vector<uint8_t> in(GLX_BGZF_IN_PAD + 64);
    std::vector<uint32_t> stage_w((GLX_BGZF_OUT + 64) / 4);
    static std::vector<uint32_t> pw;  // powers + multiplication tables, once per process
    if (pw.empty()) {
        pw.resize(GLX_BGZF_CRC_WORDS);
        bgzf_crc_powers(pw.data());
    }
    auto sh = std::make_unique<BgzfShared>();
    auto tab = std::make_unique<BcfSpanTab>();
    *n_stored = 0;
    const unsigned long long n_spans = (X.total + span - 1) / span;
    // sampling pass
    std::vector<uint32_t> hist(GLX_BGZF_HIST, 0);
    const unsigned long long n_samp = std::min<unsigned long long>(n_spans, GLX_BGZF_SAMPLES);
    for (unsigned long long i = 0; i < n_samp; i++) {
        const unsigned long long b = n_spans <= GLX_BGZF_SAMPLES ? i : i * n_spans / GLX_BGZF_SAMPLES;
        const unsigned long long lo = b * span, hi = std::min<unsigned long long>(X.total, lo + span);
        for (uint32_t k = 0; k < GLX_BGZF_HIST; k++) sh->freq[k] = 0;
        bcf_fill_span<BcfPadded>(X.E, lo, hi, bcf_first_site(X.E, lo), in.data(), *tab, 0, 1);
        bgzf_tokenize<true, false>(*sh, *tab, in.data(), (uint32_t)(hi - lo), nullptr, 0, 1);
        for (uint32_t k = 0; k < GLX_BGZF_HIST; k++) hist[k] += sh->freq[k];
    }
    uint32_t scratch[64];
    bgzf_table_build(hist.data(), sh->tab, scratch, 0, 1);
    for (unsigned long long lo = 0; lo < X.total; lo += span) {
        const unsigned long long hi = std::min<unsigned long long>(X.total, lo + span);
        const uint32_t n = (uint32_t)(hi - lo);
        std::fill(stage_w.begin(), stage_w.end(), 0u);
        uint8_t* stage = reinterpret_cast<uint8_t*>(stage_w.data());
        uint32_t* out_words = reinterpret_cast<uint32_t*>(stage + GLX_BGZF_LEAD + 18);
        sh->n_in = n;
        bgzf_crc_setup(*sh, 0, 1);
        bcf_fill_span<BcfPadded>(X.E, lo, hi, bcf_first_site(X.E, lo), in.data(), *tab, 0, 1);
        bgzf_put_header(*sh, out_words, 0, 1);
        bgzf_tokenize<false, true>(*sh, *tab, in.data(), n, pw.data(), 0, 1);
        bgzf_crc_finish(*sh, pw.data());
        bgzf_scan_bits_serial(*sh);
        bgzf_emit(*sh, *tab, in.data(), n, out_words, 0, 1);
        bgzf_finish_deflate(*sh, out_words);
        if (sh->stored) {
            bgzf_store_raw(*sh, in.data(), stage, 0, 1);
            (*n_stored)++;
        }
        bgzf_frame(*sh, stage);
        all.insert(all.end(), stage + GLX_BGZF_LEAD, stage + GLX_BGZF_LEAD + sh->block_len);
    }
    uint8_t* o = (uint8_t*)malloc(all.size() ? all.size() : 1);
    memcpy(o, all.data(), all.size());
    *bytes = o;
    *n_bytes = all.size();
    return 0;
}

// GLX_WF_SHAPES (include/glx.h): decode every record's shape code the way glx_wire_shapes_kernel does and compare with the record
// store the wire form was built from. Returns the number of records that differ, or -1 when the wire form carries no shapes;
// *n_exc / *n_shapes report how the batch was coded.
extern "C" long long glx_emu_wire_shapes_check(const glx_wire* w, const glx_records* recs, uint64_t* n_shapes, uint64_t* n_exc) {
    if (!(w->flags & GLX_WF_SHAPES)) return -1;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(w);
    const uint64_t R = w->n_records, C = w->n_cols, row = 4 + C;
    if (R != recs->n_records || C != recs->n_cols) return R + 1;
    const uint8_t* code = base + w->off[GLX_WS_SHAPE];
    const uint8_t* table = base + w->off[GLX_WS_SHAPE_TABLE];
    const uint32_t* exc_idx = reinterpret_cast<const uint32_t*>(base + w->off[GLX_WS_SHAPE_EXC_IDX]);
    const uint8_t* exc_rows = base + w->off[GLX_WS_SHAPE_EXC];
    if (n_shapes) *n_shapes = w->n_shapes;
    if (n_exc) *n_exc = w->n_shape_exc;
    long long bad = 0;
    for (uint64_t r = 0; r < R; r++) {
        const uint8_t* p;
        if (code[r] != 0xFF) {
            if (code[r] >= w->n_shapes) {
                bad++;
                continue;
            }
            p = table + code[r] * row;
        } else {
            const uint32_t* it = std::lower_bound(exc_idx, exc_idx + w->n_shape_exc, (uint32_t)r);
            if (it == exc_idx + w->n_shape_exc || *it != (uint32_t)r) {
                bad++;
                continue;
            }
            p = exc_rows + (it - exc_idx) * row;
        }
        bool ok = p[0] == recs->flags[r] && p[1] == recs->n_allele[r];
        for (int h = 0; h < 2; h++) {
            const int16_t g = (int16_t)(recs->gt[r] >> (16 * h));
            ok = ok && p[2 + h] == (g == GLX_GT16_VECTOR_END ? (uint8_t)0xFF : (uint8_t)(g >> 1));
        }
        for (uint64_t c = 0; c < C; c++) {
            const uint16_t cl = recs->col_len[c * R + r];
            ok = ok && p[4 + c] == (cl >= 0xFFF0 ? (uint8_t)(cl & 0xFF) : (uint8_t)cl);
        }
        bad += !ok;
    }
    return bad;
}
