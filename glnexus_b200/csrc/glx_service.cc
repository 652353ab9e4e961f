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

Status genotype_sites(const genotyper_config& cfg, const std::vector<GvcfDataset>& datasets, const VcfHeaderInfo& hdr,
                      const std::vector<unified_site>& sites, const std::string& filename, const volatile int* abort, int device,
                      size_t batch_sites, ServiceStats* stats) {
    Status s;
    if (batch_sites == 0) batch_sites = 4096;
    std::vector<std::string> samples;
    GLX_S(sample_set_of(datasets, samples));
    const bool bcf = cfg.output_format == GLnexusOutputFormat::BCF;

    FILE* fp = nullptr;
    ServiceStats st;
    bool header_written = false;
    auto write_all = [&](const void* p, size_t n) -> Status {
        return fwrite(p, 1, n, fp) == n ? Status::OK() : Status::IOError("bcf_write", filename);
    };
    auto write_header = [&](const PackedBatch& b) -> Status {
        if (header_written) return Status::OK();
        header_written = true;
        if (!bcf) {
            std::string h;
            vcf_header_text(b, h);
            return write_all(h.data(), h.size());
        }
        const std::string pro = bcf_file_prologue(b);  // bcf_hdr_write: its own BGZF block(s)
        std::string blocks;
        for (size_t off = 0; off < pro.size(); off += 65280)
            blocks += bgzf_block_stored(reinterpret_cast<const uint8_t*>(pro.data()) + off, std::min<size_t>(65280, pro.size() - off));
        return write_all(blocks.data(), blocks.size());
    };
    RangeStream rs;
    s = rs.open(device, bcf ? RangeStream::Mode::BGZF : RangeStream::Mode::COLUMNS, 3, [&](const RangeStream::Result& r) -> Status {
        Status ws = write_header(*r.batch);
        if (ws.bad()) return ws;
        if (bcf) {
            st.sites_written += r.n_records;
            return write_all(r.bytes, r.n_bytes);
        }
        std::string rows;
        ws = vcf_rows_text(*r.batch, *r.cols, rows);
        if (ws.bad()) return ws;
        st.sites_written += std::count(rows.begin(), rows.end(), '\n');
        return write_all(rows.data(), rows.size());
    });
    if (s.bad()) return s;
    fp = filename == "-" ? stdout : fopen(filename.c_str(), "wb");
    if (!fp) return Status::IOError("failed to open output file for write", filename);

    // contigs in order of first appearance; sites of one contig are packed once, then cut into range batches
    size_t i = 0;
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
            std::unique_ptr<PackedBatch> b(new PackedBatch());
            s = slice_batch(whole, (uint32_t)lo, (uint32_t)hi, *b);
            if (s.bad()) break;
            PackedBatch* raw = b.get();
            s = rs.submit(raw, std::move(b));
        }
        i = j;
    }
    // drain in submission order; the first error wins (src/service.cc:388-418)
    Status ds = rs.drain();
    if (s.ok()) s = ds;
    st.batches = rs.stats().batches;
    st.sites = rs.stats().sites;
    st.cells = rs.stats().cells;
    if (s.ok() && !header_written) {  // no sites: header only, like an empty pVCF
        PackedBatch header_only;
        s = pack_sites(cfg, {}, (uint32_t)samples.size(), header_only);
        if (s.ok()) {
            header_only.sample_names = samples;
            header_only.contigs = hdr.contigs;
            s = write_header(header_only);
        }
    }
    if (s.ok() && bcf) {
        const std::string eof = bgzf_eof_block();  // bcf_close
        s = write_all(eof.data(), eof.size());
    }
    if (fp != stdout) fclose(fp);
    else fflush(fp);
    rs.close();
    if (stats) *stats = st;
    return s;
}

}  // namespace glx
