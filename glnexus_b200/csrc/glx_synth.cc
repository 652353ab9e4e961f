// glx_synth.cc — seeded synthetic gVCF cohorts in record-store form (BASELINE.json configs[1..4]).
//
// There are no datasets in this environment, so the measured workloads are generated: a DeepVariant- or
// GATK-style cohort of single-sample gVCFs over one contiguous region of chr20 — reference bands
// (`REF,<*>` with GT/GQ/MIN_DP/PL), variant records (`REF,ALT..,<*>` with GT/GQ/DP/AD/PL[/SB]), coverage gaps,
// RefCall records, overlapping variant records, lost alleles — plus the unified sites the unifier would have
// produced for them (alleles, AF priors, implicit unification). Output is a PackedBatch identical in layout to
// what pack_sites/pack_records build from decoded gVCFs, so the oracle, the kernels and the pVCF assembler all
// consume it unchanged. Deterministic in (params, seed), independent of the number of generator threads.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include "../../include/glx_host.h"
#include "glx_host.hpp"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {  // splitmix64
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return n ? (uint32_t)(next() % n) : 0; }
    int geometric(double mean) {
        if (mean <= 1.0) return 1;
        const double u = std::max(uni(), 1e-12);
        return 1 + (int)(std::log(u) / std::log(1.0 - 1.0 / mean));
    }
    int poissonish(double lambda) {  // normal approximation, clamped at 0
        const double u1 = std::max(uni(), 1e-12), u2 = uni();
        const double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
        const int v = (int)std::lround(lambda + std::sqrt(lambda) * z);
        return v < 0 ? 0 : v;
    }
};

const char BASES[4] = {'A', 'C', 'G', 'T'};

struct SynSite {
    int beg, end;
    std::vector<std::string> alleles;  // REF + ALTs
    std::vector<float> af;             // per allele (af[0] unused)
    double p_alt_total;
};

// per-thread record buffers (same shapes as PackedBatch's)
struct RecBuf {
    std::vector<uint64_t> ds_count;
    std::vector<int32_t> beg, end, maxend;
    std::vector<float> qual;
    std::vector<uint8_t> n_allele, flags;
    std::vector<uint32_t> gt, allele_off;
    std::vector<std::vector<int32_t>> cv;
    std::vector<std::vector<uint16_t>> cl;
    std::vector<int32_t> pool;
    std::vector<uint32_t> al_len, al_str;
    std::vector<uint8_t> al_flags;
    std::vector<uint64_t> al_prefix;
    std::vector<char> al_pool;
};

uint64_t prefix8(const std::string& s) {
    uint64_t p = 0;
    for (size_t i = 0; i < s.size() && i < 8; i++) p |= (uint64_t)(uint8_t)s[i] << (8 * i);
    return p;
}

bool is_deletion(const std::string& ref, const std::string& alt) {
    if (alt.size() >= ref.size()) return false;
    for (size_t j = 0; j < alt.size(); j++)
        if (alt[j] != ref[j]) return false;
    return true;
}

struct Cols {
    int ref_dp = -1, ad = -1, pl = -1, gq = -1, dp = -1, sb = -1;
};

struct Gen {
    const glxh_synth_params& p;
    const glx_config& cfg;
    const std::vector<SynSite>& sites;
    Cols c;
    uint32_t n_cols;

    void set_col(RecBuf& b, int col, const int32_t* v, int n) {
        if (col < 0) return;
        if (n == 1) {
            b.cv[col].back() = v[0];
            b.cl[col].back() = 1;
        } else {
            b.cv[col].back() = (int32_t)b.pool.size();
            b.cl[col].back() = (uint16_t)n;
            b.pool.insert(b.pool.end(), v, v + n);
        }
    }
    void begin_rec(RecBuf& b, int beg, int end, int& maxend, int na, uint8_t flags, int g0, int g1, float qual) {
        maxend = std::max(maxend, end);
        b.beg.push_back(beg);
        b.end.push_back(end);
        b.maxend.push_back(maxend);
        b.qual.push_back(qual);
        b.n_allele.push_back((uint8_t)na);
        b.flags.push_back(flags);
        auto enc = [](int a) -> uint32_t { return a < 0 ? 0u : (uint32_t)(uint16_t)(int16_t)((a + 1) << 1); };
        b.gt.push_back(enc(g0) | (enc(g1) << 16));
        b.allele_off.push_back((uint32_t)b.al_len.size());
        for (uint32_t k = 0; k < n_cols; k++) {
            b.cv[k].push_back(0);
            b.cl[k].push_back(GLX_LEN_ABSENT);
        }
    }
    void add_allele(RecBuf& b, const std::string& al, const std::string& ref, int rlen, bool first, bool symbolic) {
        uint8_t f = 0;
        if (symbolic) f |= GLX_AF_SYMBOLIC;
        else {
            f |= GLX_AF_IS_DNA;
            if (!first && al.size() < (size_t)rlen && (size_t)rlen == ref.size() && is_deletion(ref, al)) f |= GLX_AF_DELETION;
        }
        b.al_len.push_back((uint32_t)al.size());
        b.al_flags.push_back(f);
        b.al_prefix.push_back(prefix8(al));
        b.al_str.push_back((uint32_t)b.al_pool.size());
        b.al_pool.insert(b.al_pool.end(), al.begin(), al.end());
    }

    void band(RecBuf& b, Rng& rng, int beg, int end, int& maxend) {
        const bool noncalled = rng.uni() < 0.003;
        const int dp = rng.poissonish(28.0);
        const int gq = noncalled ? 0 : (int)rng.below(51);
        begin_rec(b, beg, end, maxend, 2, GLX_RF_IS_REF | GLX_RF_SYMBOLIC_LAST, noncalled ? -1 : 0, noncalled ? -1 : 0, NAN);
        b.qual.back() = 0.0f;
        int32_t v = dp;
        set_col(b, c.ref_dp, &v, 1);
        v = gq;
        set_col(b, c.gq, &v, 1);
        if (p.gatk_style) {
            v = dp + (int)rng.below(4);
            set_col(b, c.dp, &v, 1);
        }
        int32_t pl[3] = {0, std::min(990, 3 * dp), std::min(990, 30 * dp)};
        if (noncalled) {
            pl[0] = 20;
            pl[1] = 0;
        }
        set_col(b, c.pl, pl, 3);
    }

    // one variant record over site `st` with record alleles REF + alts[] + symbolic; gt indexes record alleles
    void variant(RecBuf& b, Rng& rng, const SynSite& st, const std::vector<std::string>& alts, int g0, int g1, int& maxend, bool weak_hom) {
        const std::string& ref = st.alleles[0];
        const int na = 2 + (int)alts.size();
        const int rlen = st.end - st.beg;
        const char* sym = p.gatk_style ? "<NON_REF>" : "<*>";
        begin_rec(b, st.beg, st.end, maxend, na, GLX_RF_SYMBOLIC_LAST, g0, g1, (float)(1 + rng.below(600)) / 10.0f * 5.0f);
        add_allele(b, ref, ref, rlen, true, false);
        for (const auto& a : alts) add_allele(b, a, ref, rlen, false, false);
        add_allele(b, sym, ref, rlen, false, true);
        const int dp = std::max(2, rng.poissonish(30.0));
        int32_t ad[GLX_MAX_ALLELES] = {0};
        int called[2] = {g0, g1};
        int rem = dp;
        for (int k = 0; k < 2; k++) {
            if (called[k] < 0) continue;
            const int share = (k == 0) ? rem / 2 + (int)rng.below(5) - 2 : rem;
            const int s = std::max(0, std::min(rem, share));
            ad[called[k]] += s;
            rem -= s;
        }
        if (rng.uni() < 0.02) ad[0] = 0;
        set_col(b, c.ad, ad, na);
        int32_t v = dp;
        set_col(b, c.dp, &v, 1);
        const int nG = na * (na + 1) / 2;
        int32_t pl[GLX_MAX_ALLELES * (GLX_MAX_ALLELES + 1) / 2];
        for (int k = 0; k < nG; k++) pl[k] = 20 + (int)rng.below(971);
        const int lo = std::min(g0, g1), hi = std::max(g0, g1);
        int called_idx = -1;
        if (lo >= 0) {
            called_idx = hi * (hi + 1) / 2 + lo;
            pl[called_idx] = 0;
            if (weak_hom && lo == hi && lo > 0) pl[lo * (lo + 1) / 2] = 2 + (int)rng.below(40);  // het with REF close behind
        } else {
            pl[0] = 0;
        }
        set_col(b, c.pl, pl, nG);
        int second = 990;
        for (int k = 0; k < nG; k++)
            if (k != called_idx && pl[k] < second) second = pl[k];
        v = std::min(99, second);
        set_col(b, c.gq, &v, 1);
        if (p.gatk_style && c.sb >= 0) {
            int32_t sb[4] = {(int32_t)rng.below(20), (int32_t)rng.below(20), (int32_t)rng.below(20), (int32_t)rng.below(20)};
            set_col(b, c.sb, sb, 4);
        }
    }

    void fill_bands(RecBuf& b, Rng& rng, int a, int e, int& maxend) {
        while (a < e) {
            int len = rng.geometric(p.band_mean_len);
            if (len > e - a) len = e - a;
            if (!(rng.uni() < p.frac_uncovered)) band(b, rng, a, a + len, maxend);
            a += len;
        }
    }

    int draw_allele(Rng& rng, const SynSite& st) {
        double u = rng.uni();
        if (u >= st.p_alt_total) return 0;
        for (size_t a = 1; a < st.alleles.size(); a++) {
            if (u < st.af[a]) return (int)a;
            u -= st.af[a];
        }
        return (int)st.alleles.size() - 1;
    }

    void sample(RecBuf& b, uint32_t sample_idx) {
        Rng rng(p.seed * 0x9e3779b97f4a7c15ull + 0x51ed270b + (uint64_t)sample_idx * 0xd1342543de82ef95ull);
        const size_t r0 = b.beg.size();
        int pos = p.region_beg, maxend = INT32_MIN;
        const int region_end = p.region_beg + p.region_len;
        for (const SynSite& st : sites) {
            int h0 = draw_allele(rng, st), h1 = draw_allele(rng, st);
            const bool refcall = h0 == 0 && h1 == 0 && rng.uni() < p.frac_two_record_cells * 0.5;
            if (h0 == 0 && h1 == 0 && !refcall) continue;
            if (st.beg < pos) continue;  // covered by this sample's previous variant record
            fill_bands(b, rng, pos, st.beg, maxend);
            if (h0 > h1) std::swap(h0, h1);
            if (h0 == 0 && h1 > 0 && rng.uni() < p.frac_homalt) h0 = h1;
            std::vector<std::string> alts;
            int g0, g1;
            if (refcall) {
                alts.push_back(st.alleles[1]);
                g0 = g1 = 0;
            } else if (h0 == 0) {
                alts.push_back(st.alleles[h1]);
                g0 = 0;
                g1 = 1;
            } else if (h0 == h1) {
                alts.push_back(st.alleles[h1]);
                g0 = g1 = 1;
            } else {
                alts.push_back(st.alleles[h0]);
                alts.push_back(st.alleles[h1]);
                g0 = 1;
                g1 = 2;
            }
            if (!refcall && rng.uni() < p.frac_lost_allele) alts[0] = st.alleles[0] + "TTTTTTTTTT";  // not a unified allele
            if (!refcall && rng.uni() < 0.002) g0 = -1;                                              // half-missing input GT
            variant(b, rng, st, alts, g0, g1, maxend, rng.uni() < 0.5);
            if (!refcall && rng.uni() < p.frac_two_record_cells * 0.5) {
                // a second, overlapping variant record at the same position (unphased / overlapping variants)
                std::vector<std::string> alts2{st.alleles[1 + rng.below((uint32_t)st.alleles.size() - 1)]};
                variant(b, rng, st, alts2, 0, 1, maxend, false);
            }
            pos = st.end;
        }
        fill_bands(b, rng, pos, region_end, maxend);
        b.ds_count.push_back(b.beg.size() - r0);
    }
};

}  // namespace

extern "C" int glxh_synth_batch(const char* preset, const glxh_synth_params* pp, glxh_batch** ans, char* err, size_t errlen) {
    using namespace glx;
    auto fail = [&](const Status& s) {
        if (err && errlen) snprintf(err, errlen, "%s", s.str().c_str());
        return s.c_code();
    };
    if (!pp || !ans) return GLX_E_INVALID;
    *ans = nullptr;
    glxh_synth_params p = *pp;
    if (p.n_samples == 0 || p.region_len <= 0 || p.max_alt < 1 || p.max_alt > 8) return fail(Status::Invalid("glxh_synth_batch: bad parameters"));
    if (p.n_sites > (uint32_t)p.region_len) p.n_sites = (uint32_t)p.region_len;
    genotyper_config cfg;
    Status s = load_config_preset(preset && *preset ? preset : "DeepVariant", cfg);
    if (s.bad()) return fail(s);

    // ---- unified sites ----
    Rng rng(p.seed);
    std::vector<int> positions;
    {
        uint32_t need = p.n_sites;
        for (int i = 0; i < p.region_len && need; i++) {
            const uint32_t remaining = (uint32_t)(p.region_len - i);
            if (rng.below(remaining) < need) {
                positions.push_back(p.region_beg + i);
                need--;
            }
        }
    }
    const double af_lo = 1.0 / (2.0 * p.n_samples), af_hi = std::max((double)p.af_max, af_lo * 1.0001);
    std::vector<SynSite> syn(positions.size());
    std::vector<unified_site> usites;
    usites.reserve(positions.size());
    for (size_t i = 0; i < positions.size(); i++) {
        SynSite& st = syn[i];
        const int gap = (i + 1 < positions.size() ? positions[i + 1] : p.region_beg + p.region_len) - positions[i];
        int n_alt = 1;
        if (p.max_alt > 1 && rng.uni() < p.frac_multiallelic) n_alt = 2 + (int)rng.below(p.max_alt - 1);
        // allele shapes: 's' substitution of the first base, 'd' deletion of trailing bases, 'i' insertion
        std::vector<char> shape(n_alt + 1, 's');
        int n_sub = 0, n_del = 0;
        for (int a = 1; a <= n_alt; a++) {
            bool indel = rng.uni() < p.frac_indel || n_sub >= 3;
            if (!indel) {
                n_sub++;
                continue;
            }
            if (rng.uni() < 0.5 && n_del + 2 <= gap && n_del < 4) {
                shape[a] = 'd';
                n_del++;
            } else
                shape[a] = 'i';
        }
        const int L = 1 + n_del;
        std::string ref;
        for (int k = 0; k < L; k++) ref.push_back(BASES[rng.below(4)]);
        st.beg = positions[i];
        st.end = positions[i] + L;
        st.alleles.push_back(ref);
        int sub_k = 0, del_k = 0;
        const int ref_b = (int)(std::find(BASES, BASES + 4, ref[0]) - BASES);
        for (int a = 1; a <= n_alt; a++) {
            std::string al;
            if (shape[a] == 's') {
                al = ref;
                al[0] = BASES[(ref_b + 1 + sub_k++) % 4];
            } else if (shape[a] == 'd') {
                al = ref.substr(0, L - 1 - del_k++);
            } else {
                al = ref;
                for (int k = 0; k < a; k++) al.push_back(BASES[(ref_b + a + k) % 4]);
            }
            st.alleles.push_back(al);
        }
        const double af_total = af_lo * std::pow(af_hi / af_lo, rng.uni());
        st.af.assign(n_alt + 1, 0.0f);
        double wsum = 0;
        for (int a = 1; a <= n_alt; a++) wsum += std::ldexp(1.0, -a);
        st.p_alt_total = 0;
        for (int a = 1; a <= n_alt; a++) {
            st.af[a] = (float)(af_total * std::ldexp(1.0, -a) / wsum);
            st.p_alt_total += st.af[a];
        }
        unified_site us(range(0, st.beg, st.end));
        for (int a = 0; a <= n_alt; a++) {
            unified_allele ua(us.pos, st.alleles[a]);
            ua.quality = a ? 50 + (int)rng.below(900) : 0;
            ua.frequency = a ? st.af[a] : NAN;
            us.alleles.push_back(ua);
        }
        us.lost_allele_frequency = p.frac_lost_allele > 0 ? (float)(af_total * p.frac_lost_allele) : 0.0f;
        us.qual = 50 + (int)rng.below(900);
        us.fill_implicit_unification();
        usites.push_back(std::move(us));
    }

    auto* hb = new glxh_batch();
    PackedBatch& b = hb->b;
    s = pack_sites(cfg, usites, p.n_samples, b);
    if (s.bad()) {
        delete hb;
        return fail(s);
    }
    // column ids
    Cols c;
    auto find_col = [&](const std::string& nm) -> int {
        for (uint32_t i = 0; i < b.cfg.n_cols; i++)
            if (nm == b.cfg.cols[i].name && !b.cfg.cols[i].from_info && b.cfg.cols[i].type == GLX_TYPE_INT) return (int)i;
        return -1;
    };
    c.ref_dp = find_col(cfg.ref_dp_format);
    c.ad = find_col(cfg.allele_dp_format);
    c.pl = find_col("PL");
    c.gq = find_col("GQ");
    c.dp = find_col("DP");
    c.sb = find_col("SB");
    const uint32_t n_cols = b.cfg.n_cols;

    // ---- records, generated per sample on a few threads ----
    unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (p.n_samples < T) T = p.n_samples;
    if (p.sites_only) T = 0;  // unified sites only: every sample's record list stays empty
    std::vector<RecBuf> bufs(T);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; t++) {
        pool.emplace_back([&, t]() {
            RecBuf& rb = bufs[t];
            rb.cv.resize(n_cols);
            rb.cl.resize(n_cols);
            Gen g{p, b.cfg, syn, c, n_cols};
            const uint32_t lo = (uint32_t)((uint64_t)p.n_samples * t / T), hi = (uint32_t)((uint64_t)p.n_samples * (t + 1) / T);
            for (uint32_t sm = lo; sm < hi; sm++) g.sample(rb, sm);
        });
    }
    for (auto& th : pool) th.join();
    if (p.sites_only) {
        bufs.resize(1);
        bufs[0].ds_count.assign(p.n_samples, 0);
        bufs[0].cv.resize(n_cols);
        bufs[0].cl.resize(n_cols);
    }

    size_t R = 0;
    for (auto& rb : bufs) R += rb.beg.size();
    b.ds_rec_off.assign(1, 0);
    b.ds_n_sample.assign(p.n_samples, 1);
    b.ds_sample_off.resize(p.n_samples + 1);
    b.sample_out.resize(p.n_samples);
    b.sample_names.resize(p.n_samples);
    for (uint32_t d = 0; d <= p.n_samples; d++) b.ds_sample_off[d] = d;
    for (uint32_t d = 0; d < p.n_samples; d++) {
        b.sample_out[d] = (int32_t)d;
        char nm[24];
        snprintf(nm, sizeof nm, "S%07u", d);
        b.sample_names[d] = nm;
    }
    b.contigs = {{"chr20", 64444167}};
    b.col_val.assign((size_t)n_cols * R, 0);
    b.col_len.assign((size_t)n_cols * R, GLX_LEN_ABSENT);
    b.r_beg.reserve(R); b.r_end.reserve(R); b.r_maxend.reserve(R); b.r_qual.reserve(R); b.r_n_allele.reserve(R);
    b.r_flags.reserve(R); b.r_gt.reserve(R); b.r_allele_off.reserve(R);
    b.r_filt.assign(R, 0);
    size_t r_base = 0;
    for (auto& rb : bufs) {
        const size_t n = rb.beg.size();
        for (uint64_t cnt : rb.ds_count) b.ds_rec_off.push_back(b.ds_rec_off.back() + cnt);
        b.r_beg.insert(b.r_beg.end(), rb.beg.begin(), rb.beg.end());
        b.r_end.insert(b.r_end.end(), rb.end.begin(), rb.end.end());
        b.r_maxend.insert(b.r_maxend.end(), rb.maxend.begin(), rb.maxend.end());
        b.r_qual.insert(b.r_qual.end(), rb.qual.begin(), rb.qual.end());
        b.r_n_allele.insert(b.r_n_allele.end(), rb.n_allele.begin(), rb.n_allele.end());
        b.r_flags.insert(b.r_flags.end(), rb.flags.begin(), rb.flags.end());
        b.r_gt.insert(b.r_gt.end(), rb.gt.begin(), rb.gt.end());
        const uint32_t al_base = (uint32_t)b.al_len.size(), str_base = (uint32_t)b.al_pool.size();
        const int32_t pool_base = (int32_t)b.pool.size();
        for (uint32_t v : rb.allele_off) b.r_allele_off.push_back(v + al_base);
        b.al_len.insert(b.al_len.end(), rb.al_len.begin(), rb.al_len.end());
        b.al_flags.insert(b.al_flags.end(), rb.al_flags.begin(), rb.al_flags.end());
        b.al_prefix.insert(b.al_prefix.end(), rb.al_prefix.begin(), rb.al_prefix.end());
        for (uint32_t v : rb.al_str) b.al_str.push_back(v + str_base);
        b.al_pool.insert(b.al_pool.end(), rb.al_pool.begin(), rb.al_pool.end());
        b.pool.insert(b.pool.end(), rb.pool.begin(), rb.pool.end());
        for (uint32_t k = 0; k < n_cols; k++) {
            for (size_t i = 0; i < n; i++) {
                const uint16_t len = rb.cl[k][i];
                b.col_len[(size_t)k * R + r_base + i] = len;
                b.col_val[(size_t)k * R + r_base + i] = (len != GLX_LEN_ABSENT && len != 1) ? rb.cv[k][i] + pool_base : rb.cv[k][i];
            }
        }
        r_base += n;
        rb = RecBuf();
    }
    if (b.pool.size() > 0x7fffffffull) {
        delete hb;
        return fail(Status::NotImplemented("glxh_synth_batch: value pool exceeds 2^31 entries; generate a smaller batch"));
    }
    b.recs.n_datasets = p.n_samples;
    b.recs.n_samples_out = p.n_samples;
    b.recs.n_cols = n_cols;
    b.recs.n_records = R;
    b.rebind();
    *ans = hb;
    return GLX_OK;
}
