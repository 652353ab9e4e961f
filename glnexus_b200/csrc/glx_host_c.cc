// glx_host_c.cc — C entry points of include/glx_host.h over the C++ host pipeline.
#include <thread>
#include <atomic>
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/glx_host.h"
#include "glx_host.hpp"

// provided by the CUDA translation unit (glx_device.cu)
extern "C" void* glx_pinned_alloc(size_t bytes);
extern "C" void glx_pinned_free(void* p);
extern "C" int glx_host_register(void* p, size_t bytes);
extern "C" void glx_host_unregister(void* p);

struct glxh_out {
    glx_out out{};
    bool pinned = false;
    uint32_t n_fields = 0;
    uint64_t n_gt = 0, n_sites = 0;
    uint64_t n_fld[GLX_MAX_FIELDS] = {};
};

namespace {

int fail(const glx::Status& s, char* err, size_t errlen) {
    if (err && errlen) snprintf(err, errlen, "%s", s.str().c_str());
    return s.c_code();
}

void* out_alloc(size_t bytes, bool pinned) {
    if (bytes == 0) bytes = 16;
    void* p = pinned ? glx_pinned_alloc(bytes) : malloc(bytes);
    if (p) memset(p, 0, bytes);
    return p;
}
void out_release(void* p, bool pinned) {
    if (!p) return;
    if (pinned) glx_pinned_free(p);
    else free(p);
}

}  // namespace

extern "C" {

// read every dataset's gVCF text (independent: one of up to 16 threads each); statuses come back in dataset order
static void read_datasets(int n_datasets, const char* const* dataset_names, const char* const* vcf_texts, bool ingest_rules, std::vector<glx::GvcfDataset>& datasets,
                          std::vector<glx::VcfHeaderInfo>& headers, std::vector<glx::Status>& status) {
    datasets.assign(n_datasets, glx::GvcfDataset());
    headers.assign(n_datasets, glx::VcfHeaderInfo());
    status.assign(n_datasets, glx::Status::OK());
    size_t bytes = 0;
    for (int i = 0; i < n_datasets; i++) bytes += vcf_texts[i] ? strlen(vcf_texts[i]) : 0;
    const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>({(size_t)16, (size_t)std::thread::hardware_concurrency(), (size_t)n_datasets, bytes / (1u << 20) + 1}));
    auto one = [&](int i) { status[i] = glx::read_vcf_text(vcf_texts[i], dataset_names[i], ingest_rules, headers[i], datasets[i]); };
    if (T <= 1) {
        for (int i = 0; i < n_datasets; i++) one(i);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; t++)
        pool.emplace_back([&]() {
            for (int i = next++; i < n_datasets; i = next++) one(i);
        });
    for (auto& th : pool) th.join();
}

int glxh_batch_from_text(const char* preset, const char* config_text, const char* sites_text, int n_datasets, const char* const* dataset_names,
                         const char* const* vcf_texts, const char* contig, int test_ingest_rules, glxh_batch** ans, char* err, size_t errlen) {
    using namespace glx;
    if (!ans || !sites_text || n_datasets < 0) return GLX_E_INVALID;
    *ans = nullptr;
    Status s;
    genotyper_config cfg;
    if (preset && *preset) {
        s = load_config_preset(preset, cfg);
        if (s.bad()) return fail(s, err, errlen);
    }
    if (config_text && *config_text) {
        s = genotyper_config::of_text(config_text, cfg);
        if (s.bad()) return fail(s, err, errlen);
    }
    std::vector<GvcfDataset> datasets;
    std::vector<VcfHeaderInfo> headers;
    std::vector<Status> read_status;
    read_datasets(n_datasets, dataset_names, vcf_texts, test_ingest_rules != 0, datasets, headers, read_status);
    VcfHeaderInfo hdr0, hdr_all;
    for (int i = 0; i < n_datasets; i++) {
        VcfHeaderInfo& h = headers[i];
        s = read_status[i];
        if (s.bad()) return fail(s, err, errlen);
        if (i == 0) hdr0 = h;  // VCFData::contigs serves the first dataset's contigs (test/utils.cc:128-131)
        else {
            s = reconcile_contigs(hdr0, h, datasets[i]);  // gvcf_compatible, src/BCF_utils.cc:9-33
            if (s.bad()) return fail(s, err, errlen);
        }
        for (const auto& p : h.format_types) hdr_all.format_types.insert(p);
        for (const auto& p : h.info_types) hdr_all.info_types.insert(p);
    }
    hdr_all.contigs = hdr0.contigs;
    hdr_all.n_declared_contigs = hdr0.n_declared_contigs;
    std::vector<unified_site> sites;
    s = sites_of_text(sites_text, hdr_all.contigs, sites);
    if (s.bad()) return fail(s, err, errlen);
    int rid = -1;
    if (contig) {
        for (size_t i = 0; i < hdr_all.contigs.size(); i++)
            if (hdr_all.contigs[i].first == contig) rid = (int)i;
        if (rid < 0) return fail(Status::Invalid("glxh_batch_from_text: unknown contig", contig), err, errlen);
        std::vector<unified_site> keep;
        for (auto& us : sites)
            if (us.pos.rid == rid) keep.push_back(us);
        sites.swap(keep);
    } else if (!sites.empty()) {
        rid = sites[0].pos.rid;
        for (auto& us : sites)
            if (us.pos.rid != rid) return fail(Status::Invalid("glxh_batch_from_text: sites span several contigs; pass `contig`"), err, errlen);
    }
    for (int i = 0; i < n_datasets; i++) {
        // records of other contigs are outside this range batch; contig names resolve per dataset header
        std::vector<GvcfRecord> keep;
        for (auto& r : datasets[i].records)
            if (rid < 0 || r.rid == rid) keep.push_back(std::move(r));
        datasets[i].records.swap(keep);
    }
    std::set<std::string> sample_set;
    for (const auto& ds : datasets) {
        for (const auto& sm : ds.samples) {
            if (!sample_set.insert(sm).second) return fail(Status::Invalid("sample is duplicated", sm), err, errlen);
        }
    }
    std::vector<std::string> samples(sample_set.begin(), sample_set.end());
    auto* b = new glxh_batch();
    s = pack_sites(cfg, sites, (uint32_t)samples.size(), b->b);
    if (s.ok()) s = pack_records(datasets, hdr_all, samples, b->b);
    if (s.bad()) {
        delete b;
        return fail(s, err, errlen);
    }
    *ans = b;
    return GLX_OK;
}

int glxh_partition_sites(const glxh_batch* b, int n_parts, uint32_t* bounds) {
    if (!b || n_parts < 1 || !bounds) return GLX_E_INVALID;
    std::vector<uint32_t> v;
    glx::partition_sites(b->b, n_parts, v);
    for (int i = 0; i <= n_parts; i++) bounds[i] = v[i];
    return GLX_OK;
}

int glxh_partition_cohort(const glxh_batch* const* segments, uint32_t n_segments, int n_parts, uint32_t* cut_seg, uint32_t* cut_site) {
    if (!segments || !n_segments || n_parts < 1 || !cut_seg || !cut_site) return GLX_E_INVALID;
    std::vector<const glx::PackedBatch*> v;
    for (uint32_t g = 0; g < n_segments; g++) v.push_back(&segments[g]->b);
    std::vector<std::pair<uint32_t, uint32_t>> cuts;
    glx::partition_cohort(v, n_parts, cuts);
    for (int p = 0; p <= n_parts; p++) {
        cut_seg[p] = cuts[p].first;
        cut_site[p] = cuts[p].second;
    }
    return GLX_OK;
}

int glxh_batch_slice(const glxh_batch* b, uint32_t site_lo, uint32_t site_hi, glxh_batch** ans, char* err, size_t errlen) {
    if (!b || !ans) return GLX_E_INVALID;
    auto* o = new glxh_batch();
    glx::Status s = glx::slice_batch(b->b, site_lo, site_hi, o->b);
    if (s.bad()) {
        delete o;
        return fail(s, err, errlen);
    }
    *ans = o;
    return GLX_OK;
}

// Page-lock every array of the batch so that glx_upload's copies run at PCIe rate and asynchronously
// (the north star's "columnar pinned buffers"). Undone by glxh_batch_free.
int glxh_batch_pin(glxh_batch* hb) {
    if (!hb) return GLX_E_INVALID;
    if (hb->pinned) return GLX_OK;
    glx::PackedBatch& b = hb->b;
    int rc = GLX_OK;
    auto pin = [&](auto& v) {
        if (v.empty() || rc != GLX_OK) return;
        void* p = (void*)v.data();
        if (glx_host_register(p, v.size() * sizeof(v[0])) != GLX_OK) rc = GLX_E_CUDA;
        else hb->pinned_ptrs.push_back(p);
    };
    pin(b.ds_rec_off); pin(b.r_beg); pin(b.r_end); pin(b.r_maxend); pin(b.r_qual); pin(b.r_n_allele); pin(b.r_flags); pin(b.r_filt); pin(b.r_gt);
    pin(b.r_allele_off); pin(b.col_val); pin(b.col_len); pin(b.pool); pin(b.al_len); pin(b.al_flags); pin(b.al_prefix); pin(b.al_str); pin(b.al_pool);
    pin(b.s_beg); pin(b.s_end); pin(b.s_qbeg); pin(b.s_qend); pin(b.s_n_allele); pin(b.s_mono); pin(b.s_allele_off); pin(b.s_is_indel);
    pin(b.s_hom_prior); pin(b.s_lost_prior); pin(b.s_unif_off); pin(b.u_beg); pin(b.u_end); pin(b.u_to); pin(b.u_len); pin(b.u_prefix); pin(b.u_str);
    pin(b.u_pool);
    for (uint32_t k = 0; k < b.cfg.n_fields; k++) {
        pin(b.fld_cnt[k]);
        pin(b.fld_off[k]);
    }
    pin(b.s_af);
    {  // what the BCF encoder uploads per batch (glx_bcf_attach): page-locked too, or its copies would stall the stream
        glx::build_bcf_meta(b);
        glx::BcfMetaStore& m = b.bcf;
        pin(m.rid); pin(m.qual); pin(m.dna_off); pin(m.dna_pool); pin(m.norm_beg); pin(m.norm_end); pin(m.norm_off); pin(m.norm_pool); pin(m.aq);
        pin(m.contig_off); pin(m.contig_pool); pin(m.filter_off); pin(m.filter_pool);
    }
    hb->pinned = rc == GLX_OK;
    return rc;
}

void glxh_debug_wire_shape_cap(uint32_t cap) { glx::set_wire_shape_cap(cap); }
void glxh_debug_pack_threads(int n) { glx::set_pack_threads(n); }

int glxh_batch_wire(glxh_batch* hb, const glx_wire** wire) {
    if (!hb || !wire) return GLX_E_INVALID;
    if (hb->wire.empty()) {
        glx::Status s = glx::build_wire(hb->b, hb->wire);
        if (s.bad()) {
            hb->wire.clear();
            return s.c_code();
        }
    }
    if (hb->pinned && !hb->wire_pinned && glx_host_register(hb->wire.data(), hb->wire.size()) == GLX_OK) {
        hb->wire_pinned = true;
        hb->pinned_ptrs.push_back(hb->wire.data());
    }
    *wire = reinterpret_cast<const glx_wire*>(hb->wire.data());
    return GLX_OK;
}

// Service::genotype_sites on the device path, text in / pVCF file out (glx::genotype_sites)
int glxh_genotype_sites(const char* preset, const char* config_text, const char* sites_text, int n_datasets, const char* const* dataset_names,
                        const char* const* vcf_texts, const char* filename, int device, uint32_t batch_sites, int test_ingest_rules,
                        const volatile int* abort, uint64_t stats[4], char* err, size_t errlen) {
    using namespace glx;
    if (!sites_text || !filename || n_datasets < 0) return GLX_E_INVALID;
    Status s;
    genotyper_config cfg;
    if (preset && *preset) {
        s = load_config_preset(preset, cfg);
        if (s.bad()) return fail(s, err, errlen);
    }
    if (config_text && *config_text) {
        s = genotyper_config::of_text(config_text, cfg);
        if (s.bad()) return fail(s, err, errlen);
    }
    std::vector<GvcfDataset> datasets;
    std::vector<VcfHeaderInfo> headers;
    std::vector<Status> read_status;
    read_datasets(n_datasets, dataset_names, vcf_texts, test_ingest_rules != 0, datasets, headers, read_status);
    VcfHeaderInfo hdr_all;
    for (int i = 0; i < n_datasets; i++) {
        VcfHeaderInfo& h = headers[i];
        s = read_status[i];
        if (s.bad()) return fail(s, err, errlen);
        if (i == 0) {
            hdr_all.contigs = h.contigs;
            hdr_all.n_declared_contigs = h.n_declared_contigs;
        } else {
            s = reconcile_contigs(hdr_all, h, datasets[i]);  // gvcf_compatible, src/BCF_utils.cc:9-33
            if (s.bad()) return fail(s, err, errlen);
        }
        for (const auto& p : h.format_types) hdr_all.format_types.insert(p);
        for (const auto& p : h.info_types) hdr_all.info_types.insert(p);
    }
    std::vector<unified_site> sites;
    s = sites_of_text(sites_text, hdr_all.contigs, sites);
    if (s.bad()) return fail(s, err, errlen);
    ServiceStats st;
    s = genotype_sites(cfg, datasets, hdr_all, sites, filename, abort, device, batch_sites, &st);
    if (stats) {
        stats[0] = st.sites;
        stats[1] = st.sites_written;
        stats[2] = st.batches;
        stats[3] = st.cells;
    }
    if (s.bad()) return fail(s, err, errlen);
    return GLX_OK;
}

void glxh_batch_free(glxh_batch* b) {
    if (!b) return;
    for (void* p : b->pinned_ptrs) glx_host_unregister(p);
    delete b;
}
const glx_config* glxh_config(const glxh_batch* b) { return &b->b.cfg; }
const glx_records* glxh_records(const glxh_batch* b) { return &b->b.recs; }
const glx_sites* glxh_sites(const glxh_batch* b) { return &b->b.sites; }
uint32_t glxh_n_fields(const glxh_batch* b) { return b->b.cfg.n_fields; }

int glxh_device_supported(const glxh_batch* b) {
    // glx_upload_on's preconditions: every sample of every dataset is in the sample set, once (multi-sample datasets are fine:
    // they take the general multi-sample kernel; only single-sample datasets in output order have the fast paths and a wire form)
    const glx_records& r = b->b.recs;
    std::vector<char> seen(r.n_samples_out, 0);
    for (uint32_t d = 0; d < r.n_datasets; d++) {
        if (r.ds_n_sample[d] == 0 || r.ds_n_sample[d] > 0xFFFFu) return 0;
        for (uint32_t k = 0; k < r.ds_n_sample[d]; k++) {
            const int32_t o = r.sample_out[r.ds_sample_off[d] + k];
            if (o < 0 || (uint32_t)o >= r.n_samples_out || seen[o]) return 0;
            seen[o] = 1;
        }
    }
    return 1;
}
int glxh_has_fast_path(const glxh_batch* b) {
    const glx_records& r = b->b.recs;
    if (r.n_datasets != r.n_samples_out) return 0;
    for (uint32_t d = 0; d < r.n_datasets; d++)
        if (r.ds_n_sample[d] != 1 || r.sample_out[r.ds_sample_off[d]] != (int32_t)d) return 0;
    return 1;
}

void glxh_batch_stats(const glxh_batch* hb, uint64_t st[8]) {
    const glx::PackedBatch& b = hb->b;
    auto bytes = [](const auto& v) { return (uint64_t)v.size() * sizeof(v[0]); };
    st[0] = b.sites.n_sites;
    st[1] = b.sites.n_samples;
    st[2] = b.recs.n_records;
    st[3] = b.recs.n_datasets;
    uint64_t in = bytes(b.ds_rec_off) + bytes(b.r_beg) + bytes(b.r_end) + bytes(b.r_maxend) + bytes(b.r_qual) + bytes(b.r_n_allele) + bytes(b.r_flags) +
                  bytes(b.r_filt) + bytes(b.r_gt) + bytes(b.r_allele_off) + bytes(b.col_val) + bytes(b.col_len) + bytes(b.pool) + bytes(b.gt_pool) +
                  bytes(b.al_len) + bytes(b.al_str) + bytes(b.al_flags) + bytes(b.al_prefix) + bytes(b.al_pool) + bytes(b.s_beg) + bytes(b.s_end) +
                  bytes(b.s_qbeg) + bytes(b.s_qend) + bytes(b.s_n_allele) + bytes(b.s_mono) + bytes(b.s_is_indel) + bytes(b.u_to) + bytes(b.s_allele_off) +
                  bytes(b.s_unif_off) + bytes(b.u_len) + bytes(b.u_str) + bytes(b.s_hom_prior) + bytes(b.s_lost_prior) + bytes(b.u_beg) + bytes(b.u_end) +
                  bytes(b.u_prefix) + bytes(b.u_pool);
    for (uint32_t k = 0; k < b.cfg.n_fields; k++) in += bytes(b.fld_cnt[k]) + bytes(b.fld_off[k]);
    st[4] = in;
    uint64_t nv = 0;
    for (auto f : b.r_flags) nv += !(f & GLX_RF_IS_REF);
    st[5] = nv;
    st[6] = st[7] = 0;
}

glxh_out* glxh_out_alloc(const glxh_batch* b, int pinned) {
    auto* o = new glxh_out();
    o->pinned = pinned != 0;
    const uint64_t S = b->b.sites.n_sites, N = b->b.sites.n_samples;
    o->n_sites = S;
    o->n_gt = S * N * 2;
    o->n_fields = b->b.cfg.n_fields;
    o->out.gt = (int8_t*)out_alloc(o->n_gt, o->pinned);
    o->out.rnc = (uint8_t*)out_alloc(o->n_gt, o->pinned);
    o->out.allele_used = (uint32_t*)out_alloc(S * 4, o->pinned);
    o->out.site_flags = (uint8_t*)out_alloc(S, o->pinned);
    for (uint32_t k = 0; k < o->n_fields; k++) {
        o->n_fld[k] = b->b.fld_off[k].empty() ? 0 : b->b.fld_off[k].back();
        o->out.fld[k] = (int32_t*)out_alloc(o->n_fld[k] * 4, o->pinned);
    }
    return o;
}

void glxh_out_free(glxh_out* o) {
    if (!o) return;
    out_release(o->out.gt, o->pinned);
    out_release(o->out.rnc, o->pinned);
    out_release(o->out.allele_used, o->pinned);
    out_release(o->out.site_flags, o->pinned);
    for (uint32_t k = 0; k < GLX_MAX_FIELDS; k++) out_release(o->out.fld[k], o->pinned);
    delete o;
}

glx_out* glxh_out_struct(glxh_out* o) { return &o->out; }

int glxh_out_array(const glxh_out* o, int which, const void** ptr, uint64_t* nbytes) {
    if (!o || !ptr || !nbytes) return GLX_E_INVALID;
    switch (which) {
        case 0: *ptr = o->out.gt; *nbytes = o->n_gt; return GLX_OK;
        case 1: *ptr = o->out.rnc; *nbytes = o->n_gt; return GLX_OK;
        case 2: *ptr = o->out.allele_used; *nbytes = o->n_sites * 4; return GLX_OK;
        case 3: *ptr = o->out.site_flags; *nbytes = o->n_sites; return GLX_OK;
        default: break;
    }
    int k = which - 16;
    if (k < 0 || k >= (int)o->n_fields) return GLX_E_INVALID;
    *ptr = o->out.fld[k];
    *nbytes = o->n_fld[k] * 4;
    return GLX_OK;
}

struct glxh_out16 {
    glx_out16 out{};
    bool pinned = false;
    uint32_t n_fields = 0;
    uint64_t n_gt = 0, n_sites = 0;
    uint64_t n_fld[GLX_MAX_FIELDS] = {};
};

glxh_out16* glxh_out16_alloc(const glxh_batch* b, int pinned) {
    auto* o = new glxh_out16();
    o->pinned = pinned != 0;
    const uint64_t S = b->b.sites.n_sites, N = b->b.sites.n_samples;
    o->n_sites = S;
    o->n_gt = S * N * 2;
    o->n_fields = b->b.cfg.n_fields;
    o->out.gt = (int8_t*)out_alloc(o->n_gt, o->pinned);
    o->out.rnc = (uint8_t*)out_alloc(o->n_gt, o->pinned);
    o->out.allele_used = (uint32_t*)out_alloc(S * 4, o->pinned);
    o->out.site_flags = (uint8_t*)out_alloc(S, o->pinned);
    for (uint32_t k = 0; k < o->n_fields; k++) {
        o->n_fld[k] = b->b.fld_off[k].empty() ? 0 : b->b.fld_off[k].back();
        const glx_field& f = b->b.cfg.fields[k];
        const bool narrowable = f.type == GLX_TYPE_INT && f.kind != GLX_FK_FT && f.kind != GLX_FK_STRING;
        o->out.fld16[k] = narrowable ? (int16_t*)out_alloc(o->n_fld[k] * 2, o->pinned) : nullptr;
        o->out.fld8[k] = narrowable ? (int8_t*)out_alloc(o->n_fld[k], o->pinned) : nullptr;
        // the 32-bit form is the fallback for fields that do not fit (pageable: rare) and the home of float / FT fields
        o->out.fld32[k] = narrowable ? (int32_t*)malloc(o->n_fld[k] * 4 + 16) : (int32_t*)out_alloc(o->n_fld[k] * 4, o->pinned);  // untouched unless needed
    }
    return o;
}

void glxh_out16_free(glxh_out16* o) {
    if (!o) return;
    out_release(o->out.gt, o->pinned);
    out_release(o->out.rnc, o->pinned);
    out_release(o->out.allele_used, o->pinned);
    out_release(o->out.site_flags, o->pinned);
    for (uint32_t k = 0; k < GLX_MAX_FIELDS; k++) {
        const bool narrowable = o->out.fld16[k] != nullptr;
        out_release(o->out.fld16[k], o->pinned);
        out_release(o->out.fld8[k], o->pinned);
        out_release(o->out.fld32[k], narrowable ? false : o->pinned);
    }
    delete o;
}

glx_out16* glxh_out16_struct(glxh_out16* o) { return &o->out; }

int glxh_out16_array(const glxh_out16* o, int which, const void** ptr, uint64_t* nbytes) {
    if (!o || !ptr || !nbytes) return GLX_E_INVALID;
    switch (which) {
        case 0: *ptr = o->out.gt; *nbytes = o->n_gt; return GLX_OK;
        case 1: *ptr = o->out.rnc; *nbytes = o->n_gt; return GLX_OK;
        case 2: *ptr = o->out.allele_used; *nbytes = o->n_sites * 4; return GLX_OK;
        case 3: *ptr = o->out.site_flags; *nbytes = o->n_sites; return GLX_OK;
        default: break;
    }
    if (which >= 48) {
        const int k = which - 48;
        if (k >= (int)o->n_fields) return GLX_E_INVALID;
        *ptr = o->out.fld32[k];
        *nbytes = o->n_fld[k] * 4;
        return GLX_OK;
    }
    const int k = which - 16;  // the narrow form in use: int8 when glx_narrow_finish reported width 1, else int16
    if (k < 0 || k >= (int)o->n_fields) return GLX_E_INVALID;
    if (o->out.width[k] == 1) {
        *ptr = o->out.fld8[k];
        *nbytes = o->n_fld[k];
        return GLX_OK;
    }
    *ptr = o->out.fld16[k];
    *nbytes = o->out.fld16[k] ? o->n_fld[k] * 2 : 0;
    return GLX_OK;
}

int glxh_out16_widen(const glxh_out16* o, glx_out* wide) {
    if (!o || !wide) return GLX_E_INVALID;
    memcpy(wide->gt, o->out.gt, o->n_gt);
    memcpy(wide->rnc, o->out.rnc, o->n_gt);
    memcpy(wide->allele_used, o->out.allele_used, o->n_sites * 4);
    memcpy(wide->site_flags, o->out.site_flags, o->n_sites);
    for (uint32_t k = 0; k < o->n_fields; k++) {
        if (o->out.width[k] == 1) {
            const int8_t* s = o->out.fld8[k];
            int32_t* d = wide->fld[k];
            for (uint64_t i = 0; i < o->n_fld[k]; i++) {
                const int8_t v = s[i];
                d[i] = v == GLX_INT8_MISSING ? GLX_INT_MISSING : (v == GLX_INT8_VECTOR_END ? GLX_INT_VECTOR_END : (int32_t)v);
            }
        } else if (o->out.width[k] == 2) {
            const int16_t* s = o->out.fld16[k];
            int32_t* d = wide->fld[k];
            for (uint64_t i = 0; i < o->n_fld[k]; i++) {
                const int16_t v = s[i];
                d[i] = v == GLX_INT16_MISSING ? GLX_INT_MISSING : (v == GLX_INT16_VECTOR_END ? GLX_INT_VECTOR_END : (int32_t)v);
            }
        } else
            memcpy(wide->fld[k], o->out.fld32[k], o->n_fld[k] * 4);
    }
    return GLX_OK;
}

int glxh_assemble_vcf(const glxh_batch* b, const glx_out* out, char** text, char* err, size_t errlen) {
    if (!b || !out || !text) return GLX_E_INVALID;
    std::string o;
    glx::Status s = glx::assemble_vcf_text(b->b, *out, o);
    if (s.bad()) return fail(s, err, errlen);
    *text = (char*)malloc(o.size() + 1);
    memcpy(*text, o.c_str(), o.size() + 1);
    return GLX_OK;
}

void glxh_free_text(char* text) { free(text); }

static int give_bytes(const std::string& o, char** bytes, uint64_t* n_bytes) {
    if (!bytes || !n_bytes) return GLX_E_INVALID;
    *bytes = (char*)malloc(o.size() + 1);
    if (!*bytes) return GLX_E_FAILURE;
    memcpy(*bytes, o.data(), o.size());
    (*bytes)[o.size()] = 0;
    *n_bytes = o.size();
    return GLX_OK;
}

const glx_bcf_meta* glxh_bcf_meta(glxh_batch* b) { return b ? &glx::build_bcf_meta(b->b) : nullptr; }

int glxh_bcf_prologue(const glxh_batch* b, char** bytes, uint64_t* n_bytes) {
    if (!b) return GLX_E_INVALID;
    return give_bytes(glx::bcf_file_prologue(b->b), bytes, n_bytes);
}

int glxh_bgzf_stored(const void* p, uint64_t n, char** bytes, uint64_t* n_bytes) {
    if (!p && n) return GLX_E_INVALID;
    std::string o;
    for (uint64_t off = 0; off < n; off += 65280) o += glx::bgzf_block_stored(static_cast<const uint8_t*>(p) + off, (size_t)std::min<uint64_t>(65280, n - off));
    return give_bytes(o, bytes, n_bytes);
}

int glxh_bgzf_eof(char** bytes, uint64_t* n_bytes) { return give_bytes(glx::bgzf_eof_block(), bytes, n_bytes); }

int glxh_load_config(const char* name, const char* config_text, int more_PL, int squeeze, int trim_uncalled_alleles, char** text) {
    if (!text) return GLX_E_INVALID;
    *text = nullptr;
    glx::genotyper_config cfg;
    glx::Status s;
    if (name && *name) {
        s = glx::load_config_preset(name, cfg);
        if (s.bad()) return s.c_code();
    }
    if (config_text && *config_text) {
        s = glx::genotyper_config::of_text(config_text, cfg);
        if (s.bad()) return s.c_code();
    }
    cfg.more_PL = cfg.more_PL || more_PL;  // src/cli_utils.cc:1206-1208
    cfg.squeeze = cfg.squeeze || squeeze;
    cfg.trim_uncalled_alleles = cfg.trim_uncalled_alleles || trim_uncalled_alleles;
    std::string o = cfg.text();
    *text = (char*)malloc(o.size() + 1);
    memcpy(*text, o.c_str(), o.size() + 1);
    return GLX_OK;
}

struct glxh_stream {
    glx::RangeStream rs;
    bool capture = false, use_wire = false;
    std::string captured;
};

int glxh_stream_open(int device, int mode, int depth, int capture, glxh_stream** out, char* err, size_t errlen) {
    if (!out || (mode != 0 && mode != 2 && mode != 3)) return GLX_E_INVALID;
    *out = nullptr;
    auto* h = new glxh_stream();
    h->capture = capture != 0;
    glx::Status s = h->rs.open(device, (glx::RangeStream::Mode)mode, depth, [h](const glx::RangeStream::Result& r) -> glx::Status {
        if (h->capture && r.bytes) h->captured.append(reinterpret_cast<const char*>(r.bytes), r.n_bytes);
        return glx::Status::OK();
    });
    if (s.bad()) {
        delete h;
        return fail(s, err, errlen);
    }
    *out = h;
    return GLX_OK;
}

void glxh_stream_use_wire(glxh_stream* h, int on) {
    if (h) h->use_wire = on != 0;
}

int glxh_stream_submit(glxh_stream* h, glxh_batch* b, char* err, size_t errlen) {
    if (!h || !b) return GLX_E_INVALID;
    const glx_wire* wire = nullptr;
    if (h->use_wire) {
        int rc = glxh_batch_wire(b, &wire);
        if (rc != GLX_OK) return fail(glx::Status::NotImplemented("glxh_stream_submit: the batch has no wire form"), err, errlen);
    }
    glx::Status s = h->rs.submit(&b->b, nullptr, wire);
    return s.ok() ? GLX_OK : fail(s, err, errlen);
}

int glxh_stream_drain(glxh_stream* h, char* err, size_t errlen) {
    if (!h) return GLX_E_INVALID;
    glx::Status s = h->rs.drain();
    return s.ok() ? GLX_OK : fail(s, err, errlen);
}

void glxh_stream_stats(const glxh_stream* h, uint64_t st[8]) {
    const glx::StreamStats& t = h->rs.stats();
    st[0] = t.batches;
    st[1] = t.batches_retired;
    st[2] = t.sites;
    st[3] = t.cells;
    st[4] = t.records;
    st[5] = t.h2d_bytes;
    st[6] = t.d2h_bytes;
    st[7] = t.raw_bcf_bytes;
}

int glxh_stream_output(const glxh_stream* h, const char** bytes, uint64_t* n_bytes) {
    if (!h || !bytes || !n_bytes) return GLX_E_INVALID;
    *bytes = h->captured.data();
    *n_bytes = h->captured.size();
    return GLX_OK;
}

void glxh_stream_close(glxh_stream* h) { delete h; }
void glxh_stream_trace(glxh_stream* h, int on) {
    if (h) h->rs.set_trace(on != 0);
}
uint32_t glxh_stream_trace_rows(const glxh_stream* h, float* rows, uint32_t cap) {
    if (!h) return 0;
    const auto& t = h->rs.trace();
    for (uint32_t i = 0; i < t.size() && i < cap && rows; i++) memcpy(rows + 8 * i, t[i].data(), 8 * sizeof(float));
    return (uint32_t)t.size();
}

int glxh_preset_text(const char* name, char** text) {
    glx::genotyper_config cfg;
    glx::Status s = glx::load_config_preset(name ? name : "", cfg);
    if (s.bad()) return s.c_code();
    std::string o = cfg.text();
    *text = (char*)malloc(o.size() + 1);
    memcpy(*text, o.c_str(), o.size() + 1);
    return GLX_OK;
}
}
