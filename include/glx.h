/* glx.h — C ABI of the B200-native joint genotyper (drop-in for GLnexus' genotype_site path).
 *
 * The reference (dnanexus-rnd/GLnexus) has no FFI layer; its seam for this path is the pair of
 * C++ signatures
 *     Service::genotype_sites(cfg, sampleset, sites, filename, abort)   include/service.h:67-70
 *     genotype_site(cfg, cache, data, site, sampleset, samples, hdr, ans, ...)  include/genotyper.h:18-24
 * fed by BCFData::sampleset_range / RangeBCFIterator::next (include/data.h:86-94,149-153).
 * This header is what a GLnexus maintainer binds instead: the host decodes each contig range's
 * gVCF records ONCE into the columnar record store below (glx_records), flattens genotyper_config
 * (include/types.h:700-776) and the range's unified_site vector (include/types.h:454-506) into
 * glx_config / glx_sites, and calls glx_genotype_batch. The kernels do everything genotype_site
 * does per (site, sample) cell: record selection, allele translation, revise_genotypes, calling,
 * FORMAT lifting, censoring, RNC codes; the host turns the packed FORMAT columns into bcf1_t.
 *
 * Plain pointers and sizes only. All arrays are caller-owned. 0 = OK, negative = error
 * (codes mirror GLnexus::StatusCode, include/types.h:31-40).
 */
#ifndef GLX_H
#define GLX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLX_ABI_VERSION 1u
#define GLX_MAX_ALLELES 32 /* unified alleles per site and alleles per record (unifier max_alleles_per_site) */
#define GLX_MAX_FIELDS 12  /* genotyper_config::liftover_fields */
#define GLX_MAX_ORIG 4     /* retained_format_field::orig_names */
#define GLX_MAX_COLS 24    /* distinct source columns of the record store */
#define GLX_MAX_FILTERS 32 /* distinct non-PASS FILTER ids per batch (FT field) */

/* ---- status codes: -(GLnexus::StatusCode), include/types.h:31-40 ---- */
enum {
    GLX_OK = 0,
    GLX_E_FAILURE = -1,
    GLX_E_INVALID = -2,
    GLX_E_NOT_FOUND = -3,
    GLX_E_EXISTS = -4,
    GLX_E_IO_ERROR = -5,
    GLX_E_NOT_IMPLEMENTED = -6,
    GLX_E_ABORTED = -7,
    GLX_E_CUDA = -100,       /* CUDA runtime error (message in glx_last_error) */
    GLX_E_UNSUPPORTED = -101 /* input outside what the device path accepts (a dataset only partly in the sample set, > 2^32 cells) */
};

/* ---- sentinels, exactly htslib's (htslib/vcf.h) ---- */
#define GLX_INT_MISSING ((int32_t)0x80000000)    /* bcf_int32_missing    */
#define GLX_INT_VECTOR_END ((int32_t)0x80000001) /* bcf_int32_vector_end */
#define GLX_FLOAT_MISSING_BITS 0x7F800001u       /* bcf_float_missing    */
#define GLX_FLOAT_VECTOR_END_BITS 0x7F800002u    /* bcf_float_vector_end */
#define GLX_LEN_ABSENT 0xFFFFu                   /* column not present on this record (htslib rv -1/-3) */
#define GLX_GT16_VECTOR_END ((int16_t)-32767)    /* inline GT slot = bcf_int32_vector_end */

/* ---- retained_format_field enums (include/types.h:604-650) ---- */
enum { GLX_TYPE_INT = 0, GLX_TYPE_FLOAT = 1, GLX_TYPE_STRING = 2 };
enum { GLX_NUM_BASIC = 0, GLX_NUM_ALT = 1, GLX_NUM_GENOTYPE = 2, GLX_NUM_ALLELES = 3 };
enum { GLX_COMBI_MIN = 0, GLX_COMBI_MAX = 1, GLX_COMBI_MISSING = 2, GLX_COMBI_SEMICOLON = 3 };
/* helper class chosen by setup_format_helpers (src/genotyper_utils.h:865-931) */
enum { GLX_FK_NUMERIC = 0, GLX_FK_DP = 1, GLX_FK_AD = 2, GLX_FK_PL = 3, GLX_FK_PL2 = 4, GLX_FK_FT = 5, GLX_FK_STRING = 6 };

/* NoCallReason (include/genotyper.h:28-39); ASCII RNC codes of src/genotyper.cc:930-947 */
enum {
    GLX_RNC_NA = 0, GLX_RNC_MISSING = 1, GLX_RNC_PARTIAL = 2, GLX_RNC_DEPTH = 3, GLX_RNC_LOST_DELETION = 4,
    GLX_RNC_LOST_ALLELE = 5, GLX_RNC_UNPHASED = 6, GLX_RNC_OVERLAPPING = 7, GLX_RNC_MONOALLELIC = 8,
    GLX_RNC_INPUT_NONCALLED = 9
};

typedef struct glx_field {
    char name[16];               /* output FORMAT id */
    uint8_t kind;                /* GLX_FK_* */
    uint8_t type;                /* GLX_TYPE_* */
    uint8_t number;              /* GLX_NUM_* */
    uint8_t combi;               /* GLX_COMBI_* */
    uint8_t default_zero;        /* DefaultValueFiller::ZERO (else MISSING) */
    uint8_t ignore_non_variants;
    uint8_t n_cols;              /* #orig_names */
    uint8_t _pad;
    int32_t count;               /* RetainedFieldNumber::BASIC only */
    int16_t cols[GLX_MAX_ORIG];  /* source column ids, orig_names order (first found wins) */
} glx_field;

typedef struct glx_column {
    char name[24];     /* FORMAT/INFO id in the input gVCFs */
    uint8_t from_info; /* RetainedFieldFrom::INFO */
    uint8_t type;      /* GLX_TYPE_INT / GLX_TYPE_FLOAT */
    uint8_t _pad[6];
} glx_column;

/* flattened genotyper_config (include/types.h:700-776) */
typedef struct glx_config {
    uint32_t abi_version;
    uint8_t revise_genotypes, allow_partial_data, more_PL, squeeze;
    uint8_t trim_uncalled_alleles, top_two_half_calls, xatlas_depth_mode, _pad;
    uint32_t required_dp;
    float min_assumed_allele_frequency, snv_prior_calibration, indel_prior_calibration;
    /* libm constants evaluated on the host so device arithmetic is bit-identical
       (src/diploid.cc:105-106, src/genotyper.cc:213) */
    double neg10_log10e; /* -10.0*log10(exp(1.0)) */
    double ln10;         /* log(10.0) */
    int16_t col_ref_dp, col_allele_dp, col_pl, col_gq; /* ref_dp_format, allele_dp_format, "PL", "GQ" */
    uint32_t n_cols;
    glx_column cols[GLX_MAX_COLS];
    uint32_t n_fields;
    glx_field fields[GLX_MAX_FIELDS];
} glx_config;

/* Columnar record store: every gVCF record of one contig range, decoded once.
 * Records of dataset d are [ds_rec_off[d], ds_rec_off[d+1]), sorted by (beg,end) (range::operator<).
 * Single-sample datasets with sample_out[d]==d take the fast kernels (and have a wire form); a store with multi-sample or
 * reordered datasets takes the general multi-sample kernel (every sample of every dataset must be in the sample set, once). */
typedef struct glx_records {
    uint32_t n_datasets, n_samples_out, n_cols, _pad;
    uint64_t n_records;
    const uint64_t* ds_rec_off;    /* [D+1] */
    const uint32_t* ds_n_sample;   /* [D] */
    const uint32_t* ds_sample_off; /* [D+1] into sample_out */
    const int32_t* sample_out;     /* output sample index of each dataset sample, or -1 (not in sample set) */
    /* per record */
    const int32_t* beg;            /* bcf->pos */
    const int32_t* end;            /* pos + rlen (rlen from INFO/END when present) */
    const int32_t* maxend;         /* running max of end within the dataset (interval-join aid) */
    const float* qual;
    const uint8_t* n_allele;
    const uint8_t* flags;          /* GLX_RF_* */
    const uint32_t* filt;          /* bit i = FILTER id filter_names[i] set (non-PASS only) */
    const uint32_t* gt;            /* single-sample dataset: int16 gt0 | int16 gt1 << 16 (htslib GT ints);
                                      multi-sample: offset into gt_pool (2 int32 per sample) */
    const uint32_t* allele_off;    /* first allele descriptor (non-ref records only; n_allele entries) */
    /* per column c, record r: col_len[c*n_records+r] values per sample (GLX_LEN_ABSENT = not present);
       col_val holds the value itself when the dataset is single-sample and len==1, else an offset into pool */
    const int32_t* col_val;
    const uint16_t* col_len;
    const int32_t* pool;
    const int32_t* gt_pool;
    /* allele descriptors */
    const uint32_t* al_len;
    const uint8_t* al_flags;       /* GLX_AF_* */
    const uint64_t* al_prefix;     /* first 8 bytes of the allele, zero padded */
    const uint32_t* al_str;        /* offset into al_pool */
    const char* al_pool;
    uint64_t n_pool, n_gt_pool, n_al, n_al_pool;
    uint32_t n_filters, _pad2;
    const char* const* filter_names; /* sorted; host-side only */
} glx_records;

/* ---- N2 (start): compact wire form of glx_records for the host->device leg. The record store above is laid out for the
 * kernels (int32 columns, absolute offsets, running maxima); on the wire every column goes in the narrowest type the batch's
 * values fit, and everything the device can derive is left out: end = beg + len; maxend = end except for the listed
 * records; allele descriptor offsets, value-pool offsets and string offsets are prefix sums; allele flags / prefixes come from
 * the allele strings; QUAL (and FILTER, unless a lifted field reads it) only exist for non-reference records. glxh_batch_wire
 * builds it from a packed batch; glx_upload_wire_on expands it on the device into exactly the arrays glx_upload_on would copy.
 * One contiguous buffer: this header, then the sections at off[GLX_WS_*] (each 16-byte aligned). */
enum { GLX_WS_DS_REC_OFF = 0, GLX_WS_BEG, GLX_WS_LEN, GLX_WS_FLAGS, GLX_WS_NALLELE, GLX_WS_GT, GLX_WS_QUAL, GLX_WS_FILT, GLX_WS_COL_LEN, GLX_WS_COL_VAL,
       GLX_WS_POOL, GLX_WS_AL_LEN, GLX_WS_AL_POOL, GLX_WS_MAXEND_EXC, GLX_WS_SHAPE, GLX_WS_SHAPE_TABLE, GLX_WS_SHAPE_EXC_IDX, GLX_WS_SHAPE_EXC, GLX_WS_COUNT };
enum {
    GLX_WF_LEN16 = 1,          /* end - beg as uint16 (else uint32) */
    GLX_WF_GT8 = 2,            /* GT pair as two uint8: htslib GT int / 2, 0xFF = vector end (else the uint32 of glx_records.gt) */
    GLX_WF_COLLEN8 = 4,        /* col_len as uint8: 0xFF absent, 0xFE type clash, 0xFD length error (else uint16) */
    GLX_WF_POOL16 = 8,         /* value pool as int16 with BCF's 16-bit sentinels (else int32) */
    GLX_WF_FILT_VARIANT = 16,  /* FILTER masks only for non-reference records (no lifted field reads FILTER) */
    GLX_WF_ALLEN16 = 32,       /* allele lengths as uint16 (else uint32) */
    GLX_WF_ALLEN8 = 128,       /* allele lengths as uint8 (every allele of the batch is shorter than 256) */
    GLX_WF_FILT8 = 256,        /* FILTER masks as uint8 (the batch's FILTER dictionary has at most 8 ids) */
    GLX_WF_SHAPES = 64         /* a record's (flags, n_allele, GT pair, column lengths) travel as ONE byte: its row of the batch's shape
                                  table (GLX_WS_SHAPE_TABLE: n_shapes <= 255 rows of 4 + n_cols bytes {flags, n_allele, gt0, gt1, col_len8[]},
                                  the batch's commonest shapes), 0xFF = the record's own row is in GLX_WS_SHAPE_EXC, found by its index in
                                  the sorted GLX_WS_SHAPE_EXC_IDX. The FLAGS / NALLELE / GT / COL_LEN sections are then absent (needs
                                  GLX_WF_GT8 and GLX_WF_COLLEN8): reference bands and the few genotype patterns of a cohort are a few
                                  dozen shapes, so this takes 3 + n_cols bytes off every record */
};
#define GLX_WIRE_MAGIC 0x57584c47u /* "GLXW" */
typedef struct glx_wire {
    uint32_t magic, flags;
    uint64_t n_bytes; /* the whole buffer, header included */
    uint64_t n_records, n_variant, n_al, n_al_pool, n_pool, n_maxend_exc;
    uint64_t n_shape_exc; /* GLX_WF_SHAPES: records with code 0xFF */
    uint32_t n_shapes, _pad;
    uint32_t n_datasets, n_cols;
    uint8_t col_w[GLX_MAX_COLS]; /* bytes per scalar value of column c on the wire: 0 (the column has no scalar entry), 2 or 4 */
    uint64_t off[GLX_WS_COUNT];
} glx_wire;

enum { GLX_RF_IS_REF = 1, GLX_RF_HAPLOID = 2, GLX_RF_SYMBOLIC_LAST = 4, GLX_RF_NO_GT = 8 };
enum { GLX_AF_IS_DNA = 1, GLX_AF_DELETION = 2, GLX_AF_SYMBOLIC = 4 };

/* One batch of unified sites (include/types.h:454-506), sorted by position. */
typedef struct glx_sites {
    uint32_t n_sites, n_samples;
    const int32_t* beg;  /* site.pos */
    const int32_t* end;
    const int32_t* qbeg; /* hull of site.pos and all unification keys (src/genotyper.cc:700-706) */
    const int32_t* qend;
    const uint8_t* n_allele;
    const uint8_t* monoallelic;
    const uint32_t* allele_off;    /* [S+1] */
    const uint8_t* allele_is_indel;/* alleles[a].dna.size() != alleles[0].dna.size() */
    const float* allele_af;        /* unified_allele::frequency (NaN allowed); host/output only */
    const float* lost_af;          /* [S] unified_site::lost_allele_frequency; host only */
    const float* hom_log_prior;    /* logf(max(frequency, min_assumed_allele_frequency))  genotyper.cc:164 */
    const float* lost_log_prior;   /* [S] logf(max(lost_allele_frequency, min_AF))        genotyper.cc:154 */
    const uint32_t* unif_off;      /* [S+1] unification entries (std::map<allele,int>) */
    const int32_t* unif_beg;
    const int32_t* unif_end;
    const uint8_t* unif_to;
    const uint32_t* unif_len;
    const uint64_t* unif_prefix;
    const uint32_t* unif_str;
    const char* unif_pool;
    uint64_t n_unif, n_unif_pool, n_site_alleles;
    /* output layout: field k of site s holds fld_cnt[k][s] values per sample starting at value index fld_off[k][s] */
    const uint16_t* fld_cnt[GLX_MAX_FIELDS];
    const uint64_t* fld_off[GLX_MAX_FIELDS]; /* [S+1] */
} glx_sites;

/* Packed FORMAT columns, sample-major within a site. */
typedef struct glx_out {
    int8_t* gt;     /* [S][N][2] unified allele index, -1 = missing */
    uint8_t* rnc;   /* [S][N][2] ASCII RNC code */
    int32_t* fld[GLX_MAX_FIELDS]; /* int32 / float bits / FT filter mask; see glx_sites.fld_off */
    uint32_t* allele_used; /* [S] bit a = unified allele a appears in some GT (bcf_trim_alleles) */
    uint8_t* site_flags;   /* [S] bit0 = some call lost (residuals rule, src/genotyper.cc:837-849) */
} glx_out;

/* Narrow form of the lifted columns for the device->host leg. htslib stores every FORMAT vector of a BCF record in the
 * smallest integer type its values fit (bcf_enc_vint); an integer field whose every value in the batch fits int16 is
 * therefore delivered as int16 with BCF's own 16-bit sentinels. width[k] says which of fld16[k] / fld32[k] holds field k
 * after glx_narrow_finish (1, 2 or 4 bytes per value; float fields and FT masks are always 4). */
#define GLX_INT16_MISSING ((int16_t)0x8000)    /* bcf_int16_missing    */
#define GLX_INT16_VECTOR_END ((int16_t)0x8001) /* bcf_int16_vector_end */
#define GLX_INT8_MISSING ((int8_t)0x80)        /* bcf_int8_missing      */
#define GLX_INT8_VECTOR_END ((int8_t)0x81)     /* bcf_int8_vector_end   */
typedef struct glx_out16 {
    int8_t* gt;
    uint8_t* rnc;
    int16_t* fld16[GLX_MAX_FIELDS]; /* caller-allocated, one int16 per value */
    int32_t* fld32[GLX_MAX_FIELDS]; /* caller-allocated, one int32 per value (used when the field does not fit) */
    int8_t* fld8[GLX_MAX_FIELDS];   /* caller-allocated, one int8 per value (used when the previous batch of the context fitted int8) */
    uint8_t width[GLX_MAX_FIELDS];  /* out: 1, 2 or 4 = which of fld8 / fld16 / fld32 holds field k */
    uint32_t* allele_used;
    uint8_t* site_flags;
} glx_out16;

/* ---- N1: BCF2 record encoder (replaces the tail of genotype_site, src/genotyper.cc:869-1003, and the serialisation
 * htslib does for it: bcf_update_* / bcf_trim_alleles / bcf1_sync, written out by bcf_write in src/service.cc:266-275) ----
 * glx_bcf_meta is everything of a site's record that is not computed per cell, plus the header's dictionary indices
 * (header lines of src/service.cc:190-241 in htslib's BCF_DT_ID order). Per-allele arrays are indexed like
 * glx_sites.allele_off. Strings are not NUL-terminated; string i of a pool is [off[i], off[i+1]). */
typedef struct glx_bcf_meta {
    uint32_t n_sites, n_contigs, n_filters, _pad;
    const int32_t* rid;         /* [S] CHROM (index among the ##contig lines) */
    const float* qual;          /* [S] (float)unified_site::qual */
    const uint32_t* dna_off;    /* [nA+1] unified_allele::dna */
    const char* dna_pool;
    const int32_t* norm_beg;    /* [nA] unified_allele::normalized.pos */
    const int32_t* norm_end;
    const uint32_t* norm_off;   /* [nA+1] unified_allele::normalized.dna */
    const char* norm_pool;
    const int32_t* aq;          /* [nA] unified_allele::quality (INFO AQ); INFO AF comes from glx_sites.allele_af */
    const uint32_t* contig_off; /* [n_contigs+1] contig names (ID column) */
    const char* contig_pool;
    const uint32_t* filter_off; /* [n_filters+1] names of glx_records.filter_names (FT strings) */
    const char* filter_pool;
    uint64_t n_dna_pool, n_norm_pool, n_contig_pool, n_filter_pool;
    int32_t id_AF, id_AQ, id_MONOALLELIC, id_GT, id_RNC; /* BCF_DT_ID dictionary indices */
    int32_t id_field[GLX_MAX_FIELDS];
    uint8_t field_vl[GLX_MAX_FIELDS]; /* header Number of the field's description line: 'A', 'R', 'G', else 0 (what
                                         bcf_remove_allele_set looks at when it trims) */
} glx_bcf_meta;

typedef struct glx_ctx glx_ctx;

/* lifecycle */
int glx_create(glx_ctx** ctx, int device);
void glx_destroy(glx_ctx* ctx);
const char* glx_last_error(const glx_ctx* ctx);

/* Fill hom_log_prior / lost_log_prior (host logf, src/genotyper.cc:154,164). Arrays are caller-owned. */
int glx_compute_priors(const glx_config* cfg, uint32_t n_sites, const uint32_t* allele_off, const float* allele_af,
                       const float* lost_af, float* hom_log_prior, float* lost_log_prior);

/* Replaces the genotype_site loop of Service::genotype_sites (src/service.cc:344-381) for one batch.
 * Host buffers in, host buffers out: H2D, kernels, D2H on the context's stream; returns after completion. */
int glx_genotype_batch(glx_ctx* ctx, const glx_config* cfg, const glx_records* recs, const glx_sites* sites,
                       glx_out* out);

/* Device-resident variant: upload once, launch many (site batches of one range share the record store). */
typedef struct glx_dev_batch glx_dev_batch;
int glx_upload(glx_ctx* ctx, const glx_config* cfg, const glx_records* recs, const glx_sites* sites,
               glx_dev_batch** dev);
int glx_launch(glx_ctx* ctx, glx_dev_batch* dev, void* stream /* cudaStream_t or NULL = ctx stream */);
int glx_download(glx_ctx* ctx, glx_dev_batch* dev, glx_out* out);
int glx_check(glx_ctx* ctx, glx_dev_batch* dev); /* device-side error word -> status; waits for the batch's last glx_launch on the stream it was given */
uint64_t glx_out_bytes(const glx_dev_batch* dev);
uint64_t glx_in_bytes(const glx_dev_batch* dev);
int glx_launch_count(const glx_dev_batch* dev); /* kernels per glx_launch */
void glx_free_batch(glx_ctx* ctx, glx_dev_batch* dev);
/* Asynchronous forms for pipelining successive range batches (the reference overlaps retrieval, genotyping and
 * writing with a thread pool and futures, src/service.cc:344-418): all work of one batch goes onto `stream`
 * (cudaStream_t from glx_stream_create); with sync == 0 the calls return once the work is enqueued, host arrays
 * must stay valid until glx_stream_sync, and glx_status reports the batch's error word after its download landed.
 * Host arrays should be page-locked (glx_pinned_alloc / glx_host_register) for the copies to overlap. */
int glx_upload_on(glx_ctx* ctx, const glx_config* cfg, const glx_records* recs, const glx_sites* sites, void* stream, int sync,
                  glx_dev_batch** dev);
/* The same from the wire form (glx_wire header + sections in one buffer, page-locked for the copy to overlap): one H2D copy of
 * wire->n_bytes, then expansion kernels. The batch is indistinguishable from one uploaded with glx_upload_on. */
int glx_upload_wire_on(glx_ctx* ctx, const glx_config* cfg, const glx_wire* wire, const glx_sites* sites, void* stream, int sync, glx_dev_batch** dev);
/* Test aid: compare the device's record-store arrays with a host glx_records; bit i of *diff = array i differs (order of the
 * per-record arrays in glx_records: beg, end, maxend, qual, n_allele, flags, filt, gt, allele_off, col_val, col_len, pool,
 * al_len, al_flags, al_prefix, al_str, al_pool). qual / filt of reference records and col_val of absent entries are skipped. */
int glx_debug_compare_records(glx_ctx* ctx, glx_dev_batch* dev, const glx_records* recs, uint32_t* diff);
int glx_download_on(glx_ctx* ctx, glx_dev_batch* dev, glx_out* out, void* stream, int sync);
int glx_status(glx_ctx* ctx, glx_dev_batch* dev);
/* Narrow download: one extra pass packs each integer field to int16 — or to int8 when every value of that field in the
 * context's previous batch fitted BCF's int8 range, as htslib's typed-integer encoding would — on the device, recording
 * whether it fits; then GT, RNC and the narrow columns cross PCIe (18, or 14 with 8-bit DP/AD/GQ, instead of 32 bytes per
 * biallelic DeepVariant cell). After the stream is synchronised glx_narrow_finish sets width[] and fetches, synchronously,
 * the int32 form of any field that did not fit the width it was sent in. */
int glx_download_narrow_on(glx_ctx* ctx, glx_dev_batch* dev, glx_out16* out, void* stream, int sync);
int glx_narrow_finish(glx_ctx* ctx, glx_dev_batch* dev, glx_out16* out);
void* glx_stream_create(glx_ctx* ctx);
void glx_stream_destroy(glx_ctx* ctx, void* stream);
int glx_stream_sync(glx_ctx* ctx, void* stream);
/* Timing marks (cudaEvent_t): glx_mark records a new event on `stream`; glx_mark_ms waits for `b` and returns the device
 * time from `a` to `b` in ms (negative on error); marks are freed with glx_mark_free. */
void* glx_mark(glx_ctx* ctx, void* stream);
float glx_mark_ms(glx_ctx* ctx, void* a, void* b);
int glx_mark_wait(glx_ctx* ctx, void* mark); /* host waits until the stream has reached the mark */
void glx_mark_free(glx_ctx* ctx, void* mark);
int glx_host_register(void* p, size_t bytes); /* cudaHostRegister: page-lock caller-owned arrays */
void glx_host_unregister(void* p);
/* ---- N1: BCF2 records (and BGZF blocks) encoded on the device. glx_bcf_attach any time after the upload, on any stream
 * (the encoder waits for it): in a pipeline of several batches give it a stream that only ever copies host-to-device and
 * call it right after the upload, so that its small copies sit in the copy engine's queue behind this batch's upload and not
 * behind the next one's. glx_encode_bcf_on after glx_launch, on the stream of the launch:
 *   glx_bcf_attach     once per batch: uploads glx_bcf_meta (+ allele_af of `sites`, the glx_sites the batch was uploaded
 *                      with) and sizes the encoder's buffers; host arrays must stay valid until the stream reaches it
 *   glx_encode_bcf_on  mode 0: the batch's records as one uncompressed stream, exactly what bcf_write would hand to bgzf
 *                      (sites dropped after allele trimming have no record); mode 1: that stream cut into 65,280-byte BGZF
 *                      blocks (DEFLATE + CRC-32 on the device), contiguous, ready to append to the output file
 *   glx_bcf_length     waits until the encoder's sizes are known (an event early in the encode, not the whole stream)
 *   glx_download_bcf_on copies exactly the encoded bytes to `dst` (and the batch's status word, see glx_status); its stream may
 *                      be a different one (it waits for the encoder's end): a stream that only ever downloads keeps the
 *                      next upload on the encoder's stream from queueing behind this copy
 * glx_bcf_bound = bytes `dst` must hold in the worst case. */
int glx_bcf_attach(glx_ctx* ctx, glx_dev_batch* dev, const glx_sites* sites, const glx_bcf_meta* meta, const glx_config* cfg, void* stream);
int glx_encode_bcf_on(glx_ctx* ctx, glx_dev_batch* dev, int mode, void* stream);
int glx_bcf_length(glx_ctx* ctx, glx_dev_batch* dev, uint64_t* n_bytes, uint64_t* n_raw_bytes, uint32_t* n_blocks);
int glx_bcf_ready(glx_ctx* ctx, glx_dev_batch* dev); /* 1: the encoder has finished and glx_bcf_length will not block; 0: not yet; < 0: error */
int glx_download_bcf_on(glx_ctx* ctx, glx_dev_batch* dev, void* dst, uint64_t capacity, void* stream, int sync);
int glx_bcf_record_offsets(glx_ctx* ctx, glx_dev_batch* dev, uint64_t* rec_off /* [n_sites + 1] */);
int glx_debug_bgzf_phases(glx_ctx* ctx, uint64_t cycles[16]); /* profiling aid: cycles per phase of the BGZF kernel since the last call */
uint64_t glx_bcf_bound(const glx_dev_batch* dev);
uint64_t glx_bcf_records(const glx_dev_batch* dev); /* records in the encoded batch; valid after glx_bcf_length */

/* Measurement aids: record CUDA events around the stages of glx_launch on the launching stream; glx_kernel_ms waits
 * for the last launch and returns the durations of {record resolve kernel, tiled fast-path kernel (incl. the fused
 * lone-variant closed form), closed-form worklist kernels (bands, variant, single-record), general worklist kernel}; glx_work_count = cells the tiled kernel put on its worklist, glx_fused_count = those of them it then
 * finished itself (lone-variant-record cells its fused biallelic closed form took; the ones it declined are finished by
 * glx_cells_variant_list_kernel and are not counted). */
int glx_set_profile(glx_ctx* ctx, glx_dev_batch* dev, int on);
/* Kernel-path switches used by the parity tests and ablations (all 0 by default; they apply to batches uploaded afterwards):
 * "force_general" (the general kernel alone computes every cell), "force_dynamic_tiled", "no_resolve_sig", "no_fuse",
 * "chunks" (1..8), "profile_upload" (time the two upload-side kernels: glx_upload_kernel_ms). Unknown name: GLX_E_INVALID. */
int glx_set_option(glx_ctx* ctx, const char* name, int value);
/* {glx_prepare_inputs_kernel, glx_block_hints_kernel} ms of this batch's upload (option profile_upload) */
int glx_upload_kernel_ms(glx_ctx* ctx, glx_dev_batch* dev, float ms[2]);
int glx_kernel_ms(glx_ctx* ctx, glx_dev_batch* dev, float ms[4]);
int glx_work_count(glx_ctx* ctx, glx_dev_batch* dev, uint64_t* n);
int glx_fused_count(glx_ctx* ctx, glx_dev_batch* dev, uint64_t* n);
/* Known-answer aid: out[j(j+1)/2 + i] = alleles_gt(i, j) | alleles_gt(j, i) << 16 for i <= j < n_allele, computed by the
 * device's diploid index math (src/diploid.cc:12-50; test/diploid.cc:9-39). out holds n_allele(n_allele+1)/2 words. */
int glx_selftest_diploid(glx_ctx* ctx, uint32_t n_allele, uint32_t* out);
void* glx_stream(glx_ctx* ctx); /* the context's cudaStream_t */
/* pinned host memory for caller-owned buffers (cudaHostAlloc) */
void* glx_pinned_alloc(size_t bytes);
void glx_pinned_free(void* p);
int glx_synchronize(glx_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
