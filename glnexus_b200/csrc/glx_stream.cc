// glx_stream.cc — glx::RangeStream: the streamed range-batch driver of the device path. One device, `depth` batches in
// flight on `depth` streams, results retired strictly in submission order — the device-side analogue of the reference's
// thread pool + ordered drain in Service::genotype_sites (src/service.cc:344-418): while batch i's kernels run, batch
// i+1 uploads and batch i-1's encoded records cross PCIe. Everything device-side goes through the C ABI of include/glx.h.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "glx_host.hpp"

namespace glx {

static Status status_of_rc(int rc, glx_ctx* ctx) {
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

struct RangeStream::Slot {
    PackedBatch* batch = nullptr;
    std::unique_ptr<PackedBatch> owned;
    glx_dev_batch* dev = nullptr;
    void* stream = nullptr;
    int state = 0;  // 0 free, 1 kernels (+ encoder) queued, 2 download queued
    // encoded output (modes BCF / BGZF): pinned, grown on demand, reused across batches
    uint8_t* pinned = nullptr;
    size_t pinned_cap = 0;
    uint64_t n_out = 0, n_raw = 0, n_records = 0;
    uint32_t n_blocks = 0;
    // column output (mode COLUMNS): pinned arrays sized for the largest batch seen
    glx_out cols{};
    size_t cap_cells = 0, cap_sites = 0, cap_fld[GLX_MAX_FIELDS] = {};
    void* dl_mark = nullptr;  // BCF modes: the download stream has finished this slot's copy
    void* mark[5] = {};       // tracing
    float host_ms[3] = {};
};

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void RangeStream::set_trace(bool on) {
    tracing_ = on && ctx_;
    trace_.clear();
    if (trace_base_) glx_mark_free(ctx_, trace_base_);
    trace_base_ = nullptr;
    if (tracing_) {
        glx_synchronize(ctx_);
        trace_base_ = glx_mark(ctx_, nullptr);
        glx_synchronize(ctx_);
        trace_t0_ = now_ms();
    }
}

RangeStream::RangeStream() = default;

RangeStream::~RangeStream() { close(); }

Status RangeStream::open(int device, Mode mode, int depth, Sink sink) {
    close();
    if (depth < 2) depth = 2;
    if (depth > 8) depth = 8;
    mode_ = mode;
    sink_ = std::move(sink);
    if (glx_create(&ctx_, device) != GLX_OK) return Status::Failure("genotype_sites: no usable sm_100 CUDA device; this path has no host fallback");
    slots_.clear();
    for (int i = 0; i < depth; i++) {
        slots_.emplace_back(new Slot());
        slots_.back()->stream = glx_stream_create(ctx_);
        if (!slots_.back()->stream) return Status::Failure("RangeStream: cudaStreamCreate");
    }
    // Encoded results leave on a stream of their own. A slot's stream that ended on a D2H copy would hand its next operation
    // (the next batch's upload) to the copy engine's queue behind whatever download is running there; kept apart, a slot's
    // stream only ever uploads and computes and the download stream only ever downloads.
    dl_stream_ = glx_stream_create(ctx_);
    meta_stream_ = glx_stream_create(ctx_);
    if (!dl_stream_ || !meta_stream_) return Status::Failure("RangeStream: cudaStreamCreate");
    n_submitted_ = n_retired_ = 0;
    stats_ = StreamStats();
    failed_ = Status::OK();
    return Status::OK();
}

void RangeStream::close() {
    if (!ctx_) return;
    for (auto& sp : slots_) {
        Slot& sl = *sp;
        if (sl.stream) glx_stream_sync(ctx_, sl.stream);
        if (dl_stream_) glx_stream_sync(ctx_, dl_stream_);
        if (meta_stream_) glx_stream_sync(ctx_, meta_stream_);
        if (sl.dl_mark) glx_mark_free(ctx_, sl.dl_mark);
        sl.dl_mark = nullptr;
        if (sl.dev) glx_free_batch(ctx_, sl.dev);
        if (sl.stream) glx_stream_destroy(ctx_, sl.stream);
        if (sl.pinned) glx_pinned_free(sl.pinned);
        if (sl.cols.gt) glx_pinned_free(sl.cols.gt);
        if (sl.cols.rnc) glx_pinned_free(sl.cols.rnc);
        if (sl.cols.allele_used) glx_pinned_free(sl.cols.allele_used);
        if (sl.cols.site_flags) glx_pinned_free(sl.cols.site_flags);
        for (auto& f : sl.cols.fld)
            if (f) glx_pinned_free(f);
        for (auto& m : sl.mark)
            if (m) glx_mark_free(ctx_, m);
    }
    slots_.clear();
    if (dl_stream_) glx_stream_destroy(ctx_, dl_stream_);
    if (meta_stream_) glx_stream_destroy(ctx_, meta_stream_);
    dl_stream_ = meta_stream_ = nullptr;
    if (trace_base_) glx_mark_free(ctx_, trace_base_);
    trace_base_ = nullptr;
    tracing_ = false;
    glx_destroy(ctx_);
    ctx_ = nullptr;
}

static bool grow(void** p, size_t* cap, size_t need) {
    if (*p && *cap >= need) return true;
    if (*p) glx_pinned_free(*p);
    *cap = need + need / 4 + 4096;
    *p = glx_pinned_alloc(*cap);
    return *p != nullptr;
}

Status RangeStream::ensure_cols(Slot& sl, const PackedBatch& b) {
    const size_t cells = (size_t)b.sites.n_sites * b.sites.n_samples, S = b.sites.n_sites;
    size_t c = sl.cap_cells;
    bool ok = grow((void**)&sl.cols.gt, &c, cells * 2);
    c = sl.cap_cells;
    ok = ok && grow((void**)&sl.cols.rnc, &c, cells * 2);
    sl.cap_cells = c;
    size_t s4 = sl.cap_sites;
    ok = ok && grow((void**)&sl.cols.allele_used, &s4, S * 4 + 16);
    s4 = sl.cap_sites;
    ok = ok && grow((void**)&sl.cols.site_flags, &s4, S * 4 + 16);
    sl.cap_sites = s4;
    for (uint32_t k = 0; k < b.cfg.n_fields && ok; k++) ok = grow((void**)&sl.cols.fld[k], &sl.cap_fld[k], (b.fld_off[k].empty() ? 0 : b.fld_off[k].back()) * 4 + 16);
    return ok ? Status::OK() : Status::Failure("RangeStream: pinned allocation failed");
}

// queue the device->host copy of a slot whose kernels are queued (BCF modes: needs the encoded size, known early in the encode)
Status RangeStream::start_download(Slot& sl) {
    if (sl.state != 1) return Status::OK();
    if (mode_ == Mode::COLUMNS) {
        sl.state = 2;  // queued at submit
        return Status::OK();
    }
    uint64_t n = 0, raw = 0;
    uint32_t blocks = 0;
    Status s = status_of_rc(glx_bcf_length(ctx_, sl.dev, &n, &raw, &blocks), ctx_);
    if (s.bad()) return s;
    if (!grow((void**)&sl.pinned, &sl.pinned_cap, (size_t)n + 64)) return Status::Failure("RangeStream: pinned allocation failed");
    sl.n_out = n;
    sl.n_raw = raw;
    sl.n_blocks = blocks;
    sl.n_records = glx_bcf_records(sl.dev);
    s = status_of_rc(glx_download_bcf_on(ctx_, sl.dev, sl.pinned, sl.pinned_cap, dl_stream_, 0), ctx_);
    if (s.ok()) {
        sl.state = 2;
        sl.dl_mark = glx_mark(ctx_, dl_stream_);
        if (!sl.dl_mark) s = Status::Failure("RangeStream: cudaEventRecord");
    }
    if (tracing_) {
        sl.mark[4] = glx_mark(ctx_, dl_stream_);
        sl.host_ms[2] = (float)(now_ms() - trace_t0_);
    }
    return s;
}

// wait for the slot, check its status word, hand the result to the sink; after a failure later slots are freed without output
Status RangeStream::retire(Slot& sl) {
    if (sl.state == 0) return Status::OK();
    Status s = start_download(sl);
    if (s.ok() && sl.dl_mark) s = status_of_rc(glx_mark_wait(ctx_, sl.dl_mark), ctx_);  // (its stream's work ended before the copy began)
    else if (s.ok()) s = status_of_rc(glx_stream_sync(ctx_, sl.stream), ctx_);
    else {
        glx_stream_sync(ctx_, sl.stream);
        glx_stream_sync(ctx_, dl_stream_);
    }
    if (sl.dl_mark) glx_mark_free(ctx_, sl.dl_mark);
    sl.dl_mark = nullptr;
    if (s.ok()) s = status_of_rc(glx_status(ctx_, sl.dev), ctx_);
    if (s.ok() && failed_.ok() && sink_) {
        Result r;
        r.batch = sl.batch;
        r.index = n_retired_;
        if (mode_ == Mode::COLUMNS) r.cols = &sl.cols;
        else {
            r.bytes = sl.pinned;
            r.n_bytes = sl.n_out;
            r.n_raw_bytes = sl.n_raw;
            r.n_blocks = sl.n_blocks;
            r.n_records = sl.n_records;
        }
        s = sink_(r);
    }
    if (tracing_ && sl.mark[0]) {
        std::array<float, 8> row{};
        for (int k = 0; k < 5; k++) row[k] = sl.mark[k] ? glx_mark_ms(ctx_, trace_base_, sl.mark[k]) : -1.f;
        for (int k = 0; k < 3; k++) row[5 + k] = sl.host_ms[k];
        trace_.push_back(row);
    }
    for (auto& m : sl.mark) {
        if (m) glx_mark_free(ctx_, m);
        m = nullptr;
    }
    if (s.ok() && failed_.ok()) {
        stats_.batches_retired++;
        stats_.d2h_bytes += mode_ == Mode::COLUMNS ? glx_out_bytes(sl.dev) : sl.n_out;
        stats_.raw_bcf_bytes += sl.n_raw;
    }
    glx_free_batch(ctx_, sl.dev);
    sl.dev = nullptr;
    sl.batch = nullptr;
    sl.owned.reset();
    sl.state = 0;
    n_retired_++;
    if (s.bad() && failed_.ok()) failed_ = s;
    return s;
}

Status RangeStream::submit(PackedBatch* b, std::unique_ptr<PackedBatch> owned, const glx_wire* wire) {
    if (!ctx_) return Status::Invalid("RangeStream: not open");
    if (failed_.bad()) return failed_;
    const size_t depth = slots_.size();
    Slot& sl = *slots_[n_submitted_ % depth];
    const double t_enter = tracing_ ? now_ms() - trace_t0_ : 0;
    Status s = retire(sl);  // the batch submitted `depth` steps ago
    if (s.bad()) return s;
    if (tracing_) {
        sl.host_ms[0] = (float)t_enter;
        sl.mark[0] = glx_mark(ctx_, sl.stream);
    }
    sl.batch = b;
    sl.owned = std::move(owned);
    s = status_of_rc(wire ? glx_upload_wire_on(ctx_, &b->cfg, wire, &b->sites, sl.stream, 0, &sl.dev)
                          : glx_upload_on(ctx_, &b->cfg, &b->recs, &b->sites, sl.stream, 0, &sl.dev),
                     ctx_);
    // The encoder's site metadata goes up right behind the batch, on a copy-only stream: on the slot's stream these small
    // copies would wait for the upload's expansion kernels first and then sit in the copy engine's queue behind the NEXT
    // batch's upload (the host is that far ahead), and the encoder with them.
    if (s.ok() && mode_ != Mode::COLUMNS) {
        const glx_bcf_meta& meta = build_bcf_meta(*b);
        s = status_of_rc(glx_bcf_attach(ctx_, sl.dev, &b->sites, &meta, &b->cfg, meta_stream_), ctx_);
    }
    if (tracing_ && s.ok()) sl.mark[1] = glx_mark(ctx_, sl.stream);
    if (s.ok()) s = status_of_rc(glx_launch(ctx_, sl.dev, sl.stream), ctx_);
    if (tracing_ && s.ok()) sl.mark[2] = glx_mark(ctx_, sl.stream);
    if (s.ok()) {
        if (mode_ == Mode::COLUMNS) {
            s = ensure_cols(sl, *b);
            if (s.ok()) s = status_of_rc(glx_download_on(ctx_, sl.dev, &sl.cols, sl.stream, 0), ctx_);
            if (tracing_ && s.ok()) sl.mark[4] = glx_mark(ctx_, sl.stream);
        } else {
            s = status_of_rc(glx_encode_bcf_on(ctx_, sl.dev, mode_ == Mode::BGZF ? 1 : 0, sl.stream), ctx_);
            if (tracing_ && s.ok()) sl.mark[3] = glx_mark(ctx_, sl.stream);
        }
    }
    if (s.bad()) {
        if (sl.dev) {
            glx_stream_sync(ctx_, sl.stream);
            glx_stream_sync(ctx_, meta_stream_);
            glx_free_batch(ctx_, sl.dev);
            sl.dev = nullptr;
        }
        sl.batch = nullptr;
        sl.owned.reset();
        failed_ = s;
        return s;
    }
    sl.state = 1;
    stats_.batches++;
    stats_.sites += b->sites.n_sites;
    stats_.cells += (uint64_t)b->sites.n_sites * b->sites.n_samples;
    stats_.records += b->recs.n_records;
    stats_.h2d_bytes += glx_in_bytes(sl.dev);
    n_submitted_++;
    // Queue the downloads of the batches whose encoders have finished, oldest first, without waiting for any: the host runs
    // `depth` batches ahead of the device (it only ever blocks in retire(), on the oldest batch), so the device always has the
    // next batches' uploads and kernels queued behind the encoder it is running.
    if (mode_ != Mode::COLUMNS) {
        const uint64_t oldest = n_submitted_ > depth ? n_submitted_ - depth : 0;
        for (uint64_t k = oldest; k < n_submitted_; k++) {
            Slot& o = *slots_[k % depth];
            if (o.state != 1) continue;
            const int ready = glx_bcf_ready(ctx_, o.dev);
            if (ready <= 0) break;  // keep the downloads in submission order
            s = start_download(o);
            if (s.bad()) {
                if (failed_.ok()) failed_ = s;
                break;
            }
        }
    }
    if (tracing_) sl.host_ms[1] = (float)(now_ms() - trace_t0_);
    return s;
}

Status RangeStream::drain() {
    if (!ctx_) return Status::OK();
    Status s = failed_;
    const size_t depth = slots_.size();
    for (size_t k = 0; k < depth; k++) {  // oldest first
        Slot& sl = *slots_[(n_submitted_ + k) % depth];
        Status rs = retire(sl);
        if (s.ok() && rs.bad()) s = rs;
    }
    return s.ok() ? failed_ : s;
}

}  // namespace glx
