// glx_host.hpp — host side of the B200 genotyper: the reference-facing types and the
// decode -> pack -> (device) -> assemble pipeline around the C ABI in include/glx.h.
//
// Mirrors (names, argument meaning, error behaviour) the pieces of GLnexus the hot path touches:
//   Status / StatusCode            include/types.h:31-128
//   range, allele                  include/types.h:139-264
//   unified_allele, unified_site   include/types.h:434-506
//   retained_format_field          include/types.h:652-698
//   genotyper_config               include/types.h:700-776
//   Service::genotype_sites        src/service.cc:302-428   (glx::genotype_sites below)
// RocksDB/htslib are not available in this image; the decoded-record type GvcfRecord
// stands in for an unpacked bcf1_t and read_vcf_text() for htslib's VCF reader.
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/glx.h"

namespace glx {

enum class StatusCode { OK, FAILURE, INVALID, NOT_FOUND, EXISTS, IO_ERROR, NOT_IMPLEMENTED, ABORTED };

class Status {
    StatusCode code_;
    std::string msg_;

public:
    Status(StatusCode code = StatusCode::OK, const std::string& msg = "", const std::string& noun = "")
        : code_(code), msg_(noun.empty() ? msg : msg + " (" + noun + ")") {}
    bool ok() const { return code_ == StatusCode::OK; }
    bool bad() const { return !ok(); }
    operator StatusCode() const { return code_; }
    StatusCode code() const { return code_; }
    int c_code() const { return -static_cast<int>(code_); }
    static Status OK() { return Status(); }
#define GLX_STATUS_SUGAR(nm, code) \
    static Status nm(const std::string& msg = "", const std::string& noun = "") { return Status(code, msg, noun); }
    GLX_STATUS_SUGAR(Failure, StatusCode::FAILURE)
    GLX_STATUS_SUGAR(Invalid, StatusCode::INVALID)
    GLX_STATUS_SUGAR(NotFound, StatusCode::NOT_FOUND)
    GLX_STATUS_SUGAR(Exists, StatusCode::EXISTS)
    GLX_STATUS_SUGAR(IOError, StatusCode::IO_ERROR)
    GLX_STATUS_SUGAR(NotImplemented, StatusCode::NOT_IMPLEMENTED)
    GLX_STATUS_SUGAR(Aborted, StatusCode::ABORTED)
    std::string str() const;
};
#define GLX_S(st)              \
    do {                       \
        s = (st);              \
        if (s.bad()) return s; \
    } while (0)

bool is_dna(const std::string& s);
bool is_iupac_nucleotides(const std::string& s);
bool is_symbolic_allele(const std::string& s);

struct range {
    int rid = -1, beg = -1, end = -1;
    range() = default;
    range(int rid_, int beg_, int end_) : rid(rid_), beg(beg_), end(end_) {}
    size_t size() const { return end - beg; }
    bool operator==(const range& r) const { return rid == r.rid && beg == r.beg && end == r.end; }
    bool operator!=(const range& r) const { return !(*this == r); }
    bool operator<(const range& r) const {
        return rid < r.rid || (rid == r.rid && (beg < r.beg || (beg == r.beg && end < r.end)));
    }
    bool overlaps(const range& r) const { return rid == r.rid && end > r.beg && beg < r.end; }
    bool within(const range& r) const { return rid == r.rid && beg >= r.beg && end <= r.end; }
    bool contains(const range& r) const { return r.within(*this); }
};

struct allele {
    range pos;
    std::string dna;
    allele(const range& pos_, const std::string& dna_) : pos(pos_), dna(dna_) {}
    bool operator==(const allele& rhs) const { return pos == rhs.pos && dna == rhs.dna; }
    bool operator<(const allele& rhs) const { return pos < rhs.pos || (pos == rhs.pos && dna < rhs.dna); }
};

struct unified_allele {
    std::string dna;
    allele normalized;
    int quality = 0;
    float frequency = NAN;
    unified_allele(const range& pos, const std::string& dna_) : dna(dna_), normalized(pos, dna_) {}
};

struct unified_site {
    range pos;
    range in_target;
    std::vector<unified_allele> alleles;
    std::map<allele, int> unification;
    float lost_allele_frequency = 0.0f;
    int qual = 0;
    bool monoallelic = false;
    explicit unified_site(const range& pos_) : pos(pos_), in_target(-1, -1, -1) {}
    void fill_implicit_unification();
};

enum class GLnexusOutputFormat { BCF, VCF };
enum class RetainedFieldFrom { FORMAT, INFO };
enum class RetainedFieldType { INT, FLOAT, STRING };
enum class FieldCombinationMethod { MIN, MAX, MISSING, SEMICOLON };
enum class RetainedFieldNumber { BASIC, ALT, GENOTYPE, ALLELES };
enum class DefaultValueFiller { MISSING = 0, ZERO, NONE };

struct retained_format_field {
    std::vector<std::string> orig_names;
    std::string name;
    std::string description;
    RetainedFieldFrom from = RetainedFieldFrom::FORMAT;
    RetainedFieldType type = RetainedFieldType::INT;
    RetainedFieldNumber number = RetainedFieldNumber::BASIC;
    int count = 0;
    DefaultValueFiller default_type = DefaultValueFiller::MISSING;
    FieldCombinationMethod combi_method = FieldCombinationMethod::MISSING;
    bool ignore_non_variants = false;
};

struct genotyper_config {
    bool revise_genotypes = false;
    float min_assumed_allele_frequency = 0.0001f;
    float snv_prior_calibration = 1.0f, indel_prior_calibration = 1.0f;
    size_t required_dp = 0;
    bool allow_partial_data = false;
    std::string allele_dp_format = "AD";
    std::string ref_dp_format = "MIN_DP";
    bool output_residuals = false;
    GLnexusOutputFormat output_format = GLnexusOutputFormat::BCF;
    std::vector<retained_format_field> liftover_fields;
    bool more_PL = false;
    bool squeeze = false;
    bool trim_uncalled_alleles = false;
    bool top_two_half_calls = false;

    // line-oriented text form used by the fixtures and the CLI (see config_of_text)
    static Status of_text(const std::string& txt, genotyper_config& ans);
    std::string text() const;
};

// `glnexus_cli --config NAME` presets (src/cli_utils.cc:465-1188), genotyper half.
Status load_config_preset(const std::string& name, genotyper_config& ans);
std::vector<std::string> config_preset_names();

// ---------------------------------------------------------------------------------------------
// Decoded gVCF record: what an unpacked bcf1_t + its header give the reference through
// bcf_get_genotypes / bcf_get_format_int32 / bcf_get_info_* (htslib 1.9).
// ---------------------------------------------------------------------------------------------
struct FieldValues {
    std::string id;
    uint8_t type = GLX_TYPE_INT;  // header Type: Integer / Float / other (STRING)
    int per_sample = 0;           // values per sample (FORMAT) or total (INFO)
    std::vector<int32_t> v;       // int32 or float bits; n_sample*per_sample
};

struct GvcfRecord {
    int rid = 0, pos = 0, rlen = 0;
    float qual = NAN;
    bool qual_missing = true;
    std::vector<std::string> alleles;
    std::vector<std::string> filters;  // FILTER ids as written ("PASS" included)
    int n_sample = 0;
    int n_gt = 0;                      // GT values per sample (1 haploid, 2 diploid), 0 = no GT
    std::vector<int32_t> gt;           // htslib GT ints, n_sample*n_gt
    std::vector<FieldValues> format;
    std::vector<FieldValues> info;
    range rng() const { return range(rid, pos, pos + rlen); }
};

struct GvcfDataset {
    std::string name;
    std::vector<std::string> samples;
    std::vector<GvcfRecord> records;
};

struct VcfHeaderInfo {
    std::vector<std::pair<std::string, size_t>> contigs;
    size_t n_declared_contigs = 0;  // contigs[0..n) come from ##contig lines; the rest were added on first appearance
    std::map<std::string, uint8_t> format_types, info_types;  // GLX_TYPE_*
};

// Minimal VCF text reader (stand-in for htslib's). `test_ingest_rules` applies what VCFData /
// validate_bcf do at ingest in the reference's tests (test/utils.cc:75-84, src/BCF_utils.cc:54-66).
Status read_vcf_text(const std::string& text, const std::string& dataset_name, bool test_ingest_rules,
                     VcfHeaderInfo& hdr, GvcfDataset& ans);

// gvcf_compatible (src/BCF_utils.cc:9-33, :198): every gVCF must declare exactly the database's contigs, order included.
// `db` is the shared contig table (the first dataset's). A dataset whose ##contig lines differ is rejected with
// Status::Invalid; a dataset without ##contig lines has its records' rid remapped by contig name into `db`.
Status reconcile_contigs(const VcfHeaderInfo& db, const VcfHeaderInfo& h, GvcfDataset& ds);

Status sites_of_text(const std::string& txt, const std::vector<std::pair<std::string, size_t>>& contigs,
                     std::vector<unified_site>& ans);

// Owner of the arrays behind glx_bcf_meta (include/glx.h): built from the batch by build_bcf_meta().
struct BcfMetaStore {
    bool built = false;
    glx_bcf_meta meta{};
    std::vector<int32_t> rid, norm_beg, norm_end, aq;
    std::vector<float> qual;
    std::vector<uint32_t> dna_off, norm_off, contig_off, filter_off;
    std::vector<char> dna_pool, norm_pool, contig_pool, filter_pool;
    std::vector<std::string> dict;  // BCF_DT_ID dictionary in header order (PASS first)
};

// ---------------------------------------------------------------------------------------------
// Packed batch: owns every array the C ABI structs point at.
// ---------------------------------------------------------------------------------------------
struct PackedBatch {
    glx_config cfg{};
    glx_records recs{};
    glx_sites sites{};
    // metadata for output assembly
    genotyper_config gcfg;
    std::vector<unified_site> usites;
    std::vector<std::string> sample_names;
    std::vector<std::pair<std::string, size_t>> contigs;
    std::vector<std::string> filter_names;
    std::vector<const char*> filter_name_ptrs;
    // storage
    std::vector<uint64_t> ds_rec_off;
    std::vector<uint32_t> ds_n_sample, ds_sample_off;
    std::vector<int32_t> sample_out;
    std::vector<int32_t> r_beg, r_end, r_maxend;
    std::vector<float> r_qual;
    std::vector<uint8_t> r_n_allele, r_flags;
    std::vector<uint32_t> r_filt, r_gt, r_allele_off;
    std::vector<int32_t> col_val;
    std::vector<uint16_t> col_len;
    std::vector<int32_t> pool, gt_pool;
    std::vector<uint32_t> al_len, al_str;
    std::vector<uint8_t> al_flags;
    std::vector<uint64_t> al_prefix;
    std::vector<char> al_pool;
    std::vector<int32_t> s_beg, s_end, s_qbeg, s_qend;
    std::vector<uint8_t> s_n_allele, s_mono, s_is_indel, u_to;
    std::vector<uint32_t> s_allele_off, s_unif_off, u_len, u_str;
    std::vector<float> s_af, s_hom_prior, s_lost_prior, s_lost_af;
    std::vector<int32_t> u_beg, u_end;
    std::vector<uint64_t> u_prefix;
    std::vector<char> u_pool;
    std::vector<uint16_t> fld_cnt[GLX_MAX_FIELDS];
    std::vector<uint64_t> fld_off[GLX_MAX_FIELDS];
    BcfMetaStore bcf;  // N1: what the BCF record encoder needs (lazily built; see build_bcf_meta)
    void rebind();  // refresh the raw pointers after the vectors are final
};

struct OutBuffers {
    glx_out out{};
    std::vector<int8_t> gt;
    std::vector<uint8_t> rnc, site_flags;
    std::vector<int32_t> fld[GLX_MAX_FIELDS];
    std::vector<uint32_t> allele_used;
    void allocate(const PackedBatch& b);
};

Status make_glx_config(const genotyper_config& cfg, glx_config& ans);
Status pack_sites(const genotyper_config& cfg, const std::vector<unified_site>& sites, uint32_t n_samples,
                  PackedBatch& b);
// samples = sorted sample names of the sample set (Service::genotype_sites, src/service.cc:307-309)
Status pack_records(const std::vector<GvcfDataset>& datasets, const VcfHeaderInfo& hdr,
                    const std::vector<std::string>& samples, PackedBatch& b);

// Output assembly = src/genotyper.cc:862-1003 after the per-dataset loop + the pVCF header of
// src/service.cc:190-241, written as VCF text (htslib vcf_format value semantics).
Status assemble_vcf_text(const PackedBatch& b, const glx_out& out, std::string& vcf_text);
void vcf_header_text(const PackedBatch& b, std::string& o);
Status vcf_rows_text(const PackedBatch& b, const glx_out& out, std::string& o);

// ---------------------------------------------------------------------------------------------
// N1: BCF2 output (glx_bcf_host.cc). The records themselves are encoded on the device (glx_bcf.cu); the host owns the
// header (src/service.cc:190-241), the BCF file prologue and the BGZF framing of those few bytes.
// ---------------------------------------------------------------------------------------------
const glx_bcf_meta& build_bcf_meta(PackedBatch& b);
// pVCF header as htslib writes it into a BCF file: every FILTER/INFO/FORMAT/contig line carries its IDX
std::string bcf_header_text(const PackedBatch& b);
// "BCF\2\2", l_text, header text, NUL — the bytes before the first record (uncompressed)
std::string bcf_file_prologue(const PackedBatch& b);
uint32_t crc32_ieee(const uint8_t* p, size_t n, uint32_t crc = 0);
// one BGZF block (RFC 1952 member with the BC extra field) holding p[0..n) as a stored deflate block; n <= 65280
std::string bgzf_block_stored(const uint8_t* p, size_t n);
std::string bgzf_eof_block();

// N2 (start): the compact wire form of the batch's record store (glx_wire, include/glx.h) — glx_wire_host.cc
// bytes whose resize() leaves them uninitialised: the wire buffer's pages are first touched by the threads that fill them
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = default_init_allocator<U>;
    };
    template <class U, class... A>
    void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new (static_cast<void*>(p)) U;
        else ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
};
using WireBytes = std::vector<uint8_t, default_init_allocator<uint8_t>>;
Status build_wire(const PackedBatch& b, WireBytes& wire);
void set_pack_threads(int n);  // test aid: threads of pack_records (0 = chosen from the input's size)
void set_wire_shape_cap(uint32_t cap);  // test aid: rows of the wire form's shape table (default and maximum 255)

// Multi-GPU sharding by contiguous site range (no collective): cut points of equal output weight, and the shard
// [site_lo, site_hi) with every record its sites can touch.
void partition_sites(const PackedBatch& b, int n_parts, std::vector<uint32_t>& bounds);
void partition_cohort(const std::vector<const PackedBatch*>& segments, int n_parts, std::vector<std::pair<uint32_t, uint32_t>>& cuts);
Status slice_batch(const PackedBatch& b, uint32_t site_lo, uint32_t site_hi, PackedBatch& ans);

}  // namespace glx

// ---------------------------------------------------------------------------------------------
// The upper seam: Service::genotype_sites (include/service.h:67-70, src/service.cc:302-428) on the device path.
// Same contract: sites are genotyped and written to `filename` ("-" = stdout) in input order; `abort` is polled
// between range batches (the reference polls it per site task and per dataset); the first failing site's Status is
// returned. Instead of one task per site on a thread pool, sites go to the GPU in contiguous range batches of
// `batch_sites`, two batches in flight on two streams (retrieval/packing of the next overlaps the kernels and the
// download of the previous). Output follows cfg.output_format like BCFFileSink::Open (src/service.cc:254-265): BCF = a
// BGZF-compressed BCF2 file whose records are encoded and compressed on the device (N1), VCF = uncompressed pVCF text.
// ---------------------------------------------------------------------------------------------
#include <array>
#include <functional>
namespace glx {
// ---------------------------------------------------------------------------------------------
// RangeStream (glx_stream.cc): the streamed range-batch driver. `depth` batches in flight on one device, retired in
// submission order through `sink` — the device-side analogue of the thread pool + ordered drain of
// Service::genotype_sites (src/service.cc:344-418). Modes: COLUMNS (packed int32 FORMAT columns, include/glx.h glx_out),
// BCF (finished uncompressed BCF2 records), BGZF (those records as BGZF blocks, ready to append to the output file).
// ---------------------------------------------------------------------------------------------
struct StreamStats {
    uint64_t batches = 0, batches_retired = 0, sites = 0, cells = 0, records = 0, h2d_bytes = 0, d2h_bytes = 0, raw_bcf_bytes = 0;
};
class RangeStream {
public:
    enum class Mode { COLUMNS = 0, BCF = 2, BGZF = 3 };
    struct Result {
        const PackedBatch* batch = nullptr;
        uint64_t index = 0;             // submission index
        const glx_out* cols = nullptr;  // COLUMNS
        const uint8_t* bytes = nullptr; // BCF / BGZF
        uint64_t n_bytes = 0, n_raw_bytes = 0, n_records = 0;
        uint32_t n_blocks = 0;
    };
    typedef std::function<Status(const Result&)> Sink;
    RangeStream();
    ~RangeStream();
    Status open(int device, Mode mode, int depth, Sink sink);
    // Queue one batch (upload, kernels, encoder). `b` must stay valid until it has been retired; pass `owned` to hand it over.
    // `wire`: the batch's compact wire form (build_wire) to upload instead of its full-width record store
    Status submit(PackedBatch* b, std::unique_ptr<PackedBatch> owned = nullptr, const glx_wire* wire = nullptr);
    Status drain();  // retire everything in flight, in order; the first error wins and later batches produce no output
    void close();
    const StreamStats& stats() const { return stats_; }
    glx_ctx* ctx() const { return ctx_; }
    // Measurement aid: with tracing on, every retired batch adds one row of times in ms since set_trace(true):
    // device {upload begins, upload done, genotyping kernels done, encoder done, download done}, host {submit entered,
    // submit returned, download queued}
    void set_trace(bool on);
    const std::vector<std::array<float, 8>>& trace() const { return trace_; }

private:
    struct Slot;
    Status ensure_cols(Slot& sl, const PackedBatch& b);
    Status start_download(Slot& sl);
    Status retire(Slot& sl);
    glx_ctx* ctx_ = nullptr;
    void* dl_stream_ = nullptr;    // downloads of encoded results (modes BCF / BGZF)
    void* meta_stream_ = nullptr;  // uploads of the encoder's site metadata
    Mode mode_ = Mode::COLUMNS;
    Sink sink_;
    std::vector<std::unique_ptr<Slot>> slots_;
    uint64_t n_submitted_ = 0, n_retired_ = 0;
    StreamStats stats_;
    Status failed_;
    bool tracing_ = false;
    void* trace_base_ = nullptr;
    double trace_t0_ = 0;
    std::vector<std::array<float, 8>> trace_;
};

struct ServiceStats {
    uint64_t sites = 0, sites_written = 0, batches = 0, cells = 0;
};
Status sample_set_of(const std::vector<GvcfDataset>& datasets, std::vector<std::string>& samples);
Status genotype_sites(const genotyper_config& cfg, const std::vector<GvcfDataset>& datasets, const VcfHeaderInfo& hdr,
                      const std::vector<unified_site>& sites, const std::string& filename, const volatile int* abort, int device,
                      size_t batch_sites, ServiceStats* stats);
}  // namespace glx

// opaque handle of include/glx_host.h
struct glxh_batch {
    glx::PackedBatch b;
    bool pinned = false;
    std::vector<void*> pinned_ptrs;
    glx::WireBytes wire;  // built on demand (glxh_batch_wire)
    bool wire_pinned = false;
};
