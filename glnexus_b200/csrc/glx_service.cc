// glx_service.cc — glx::genotype_sites: the drop-in for Service::genotype_sites (src/service.cc:302-428) on the device
// path. See glx_host.hpp for the contract. Everything device-side goes through the C ABI of include/glx.h.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "glx_host.hpp"

namespace glx {

// sorted sample names of the "<ALL>" sample set (src/service.cc:307-309); a sample may appear in one dataset only
Status sample_set_of(const std::vector<GvcfDataset>& datasets, std::vector<std::string>& samples) {
    std::set<std::string> sample_set;
    for (const auto& ds : datasets)
        for (const auto& sm : ds.samples)
            if (!sample_set.insert(sm).second) return Status::Invalid("sample is duplicated", sm);
    samples.assign(sample_set.begin(), sample_set.end());
    return Status::OK();
}

namespace {

Status status_of(int rc, glx_ctx* ctx) {
    if (rc == GLX_OK) return Status::OK();
    const std::string msg = ctx ? glx_last_error(ctx) : "";
    switch (rc) {
        case GLX_E_INVALID: return Status::Invalid("genotyper", msg);
        case GLX_E_NOT_FOUND: return Status::NotFound("genotyper", msg);
        case GLX_E_EXISTS: return Status::Exists("genotyper", msg);
        case GLX_E_IO_ERROR: return Status::IOError("genotyper", msg);
        case GLX_E_NOT_IMPLEMENTED: return Status::NotImplemented("genotyper", msg);
        case GLX_E_UNSUPPORTED: return Status::NotImplemented("genotyper: input outside the device path", msg);
        case GLX_E_ABORTED: return Status::Aborted();
        case GLX_E_CUDA: return Status::Failure("genotyper: CUDA", msg);
        default: return Status::Failure("genotyper", msg);
    }
}

// pinned host columns of one batch
struct PinnedOut {
    glx_out out{};
    std::vector<void*> ptrs;
    ~PinnedOut() {
        for (void* p : ptrs) glx_pinned_free(p);
    }
    template <class T>
    T* take(size_t n) {
        void* p = glx_pinned_alloc((n ? n : 1) * sizeof(T));
        if (p) ptrs.push_back(p);
        return static_cast<T*>(p);
    }
    bool allocate(const PackedBatch& b) {
        const size_t S = b.sites.n_sites, N = b.sites.n_samples;
        out.gt = take<int8_t>(S * N * 2);
        out.rnc = take<uint8_t>(S * N * 2);
        out.allele_used = take<uint32_t>(S);
        out.site_flags = take<uint8_t>(S);
        bool ok = out.gt && out.rnc && out.allele_used && out.site_flags;
        for (uint32_t k = 0; k < b.cfg.n_fields; k++) {
            out.fld[k] = take<int32_t>(b.fld_off[k].empty() ? 0 : b.fld_off[k].back());
            ok = ok && out.fld[k];
        }
        return ok;
    }
};

struct Slot {
    std::unique_ptr<PackedBatch> batch;
    std::unique_ptr<PinnedOut> out;
    glx_dev_batch* dev = nullptr;
    void* stream = nullptr;
};

}  // namespace

Status genotype_sites(const genotyper_config& cfg, const std::vector<GvcfDataset>& datasets, const VcfHeaderInfo& hdr,
                      const std::vector<unified_site>& sites, const std::string& filename, const volatile int* abort, int device,
                      size_t batch_sites, ServiceStats* stats) {
    Status s;
    if (batch_sites == 0) batch_sites = 4096;
    std::vector<std::string> samples;
    GLX_S(sample_set_of(datasets, samples));

    glx_ctx* ctx = nullptr;
    int rc = glx_create(&ctx, device);
    if (rc != GLX_OK) return Status::Failure("genotype_sites: no usable sm_100 CUDA device; this path has no host fallback");
    Slot slots[2];
    FILE* fp = nullptr;
    ServiceStats st;
    auto cleanup = [&]() {
        for (auto& sl : slots) {
            if (sl.stream) glx_stream_sync(ctx, sl.stream);
            if (sl.dev) glx_free_batch(ctx, sl.dev);
            sl.dev = nullptr;
            if (sl.stream) glx_stream_destroy(ctx, sl.stream);
            sl.stream = nullptr;
        }
        if (fp && fp != stdout) fclose(fp);
        glx_destroy(ctx);
        if (stats) *stats = st;
    };
    for (auto& sl : slots) {
        sl.stream = glx_stream_create(ctx);
        if (!sl.stream) {
            cleanup();
            return Status::Failure("genotype_sites: cudaStreamCreate");
        }
    }
    fp = filename == "-" ? stdout : fopen(filename.c_str(), "w");
    if (!fp) {
        cleanup();
        return Status::IOError("failed to open output file for write", filename);
    }

    // retire one slot: wait, check the batch's status, write its rows in order
    bool header_written = false;
    // `write` false: an earlier batch failed — wait and free only, no rows after the gap (the reference stops writing at
    // the first failing site, src/service.cc:388-418)
    auto retire = [&](Slot& sl, bool write = true) -> Status {
        if (!sl.dev) return Status::OK();
        Status rs = status_of(glx_stream_sync(ctx, sl.stream), ctx);
        if (rs.ok()) rs = status_of(glx_status(ctx, sl.dev), ctx);
        std::string rows;
        if (rs.ok() && write) {
            if (!header_written) {
                vcf_header_text(*sl.batch, rows);
                header_written = true;
            }
            const size_t before = std::count(rows.begin(), rows.end(), '\n');
            rs = vcf_rows_text(*sl.batch, sl.out->out, rows);
            st.sites_written += std::count(rows.begin(), rows.end(), '\n') - before;
        }
        if (rs.ok() && fwrite(rows.data(), 1, rows.size(), fp) != rows.size()) rs = Status::IOError("bcf_write", filename);
        glx_free_batch(ctx, sl.dev);
        sl.dev = nullptr;
        sl.batch.reset();
        sl.out.reset();
        return rs;
    };

    // contigs in order of first appearance; sites of one contig are packed once, then cut into range batches
    size_t i = 0, n_submitted = 0;
    PackedBatch header_only;  // used when there are no sites at all
    while (i < sites.size() && s.ok()) {
        const int rid = sites[i].pos.rid;
        size_t j = i;
        while (j < sites.size() && sites[j].pos.rid == rid) j++;
        std::vector<unified_site> contig_sites(sites.begin() + i, sites.begin() + j);
        std::vector<GvcfDataset> contig_ds(datasets.size());
        for (size_t d = 0; d < datasets.size(); d++) {
            contig_ds[d].name = datasets[d].name;
            contig_ds[d].samples = datasets[d].samples;
            for (const auto& r : datasets[d].records)
                if (r.rid == rid) contig_ds[d].records.push_back(r);
        }
        PackedBatch whole;
        s = pack_sites(cfg, contig_sites, (uint32_t)samples.size(), whole);
        if (s.ok()) s = pack_records(contig_ds, hdr, samples, whole);
        for (size_t lo = 0; lo < contig_sites.size() && s.ok(); lo += batch_sites) {
            if (abort && *abort) {
                s = Status::Aborted();
                break;
            }
            const size_t hi = std::min(contig_sites.size(), lo + batch_sites);
            Slot& sl = slots[n_submitted & 1];
            s = retire(sl);  // the batch submitted two steps ago
            if (s.bad()) break;
            sl.batch.reset(new PackedBatch());
            s = slice_batch(whole, (uint32_t)lo, (uint32_t)hi, *sl.batch);
            if (s.bad()) break;
            sl.out.reset(new PinnedOut());
            if (!sl.out->allocate(*sl.batch)) {
                s = Status::Failure("genotype_sites: pinned allocation failed");
                break;
            }
            s = status_of(glx_upload_on(ctx, &sl.batch->cfg, &sl.batch->recs, &sl.batch->sites, sl.stream, 0, &sl.dev), ctx);
            if (s.ok()) s = status_of(glx_launch(ctx, sl.dev, sl.stream), ctx);
            if (s.ok()) s = status_of(glx_download_on(ctx, sl.dev, &sl.out->out, sl.stream, 0), ctx);
            if (s.bad()) break;
            n_submitted++;
            st.batches++;
            st.sites += hi - lo;
            st.cells += (uint64_t)(hi - lo) * samples.size();
        }
        i = j;
    }
    // drain in submission order; the first error wins (src/service.cc:388-418)
    for (int k = 0; k < 2; k++) {
        Slot& sl = slots[(n_submitted + k) & 1];
        Status rs = retire(sl, s.ok());
        if (s.ok() && rs.bad()) s = rs;
    }
    if (s.ok() && !header_written) {  // no sites: header only, like an empty pVCF
        s = pack_sites(cfg, {}, (uint32_t)samples.size(), header_only);
        if (s.ok()) {
            header_only.sample_names = samples;
            header_only.contigs = hdr.contigs;
            std::string h;
            vcf_header_text(header_only, h);
            if (fwrite(h.data(), 1, h.size(), fp) != h.size()) s = Status::IOError("bcf_write", filename);
        }
    }
    cleanup();
    return s;
}

}  // namespace glx
