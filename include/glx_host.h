/* glx_host.h — C entry points of the host pipeline around include/glx.h: decode (VCF text stand-in for
 * htslib), pack into the columnar record store, and assemble pVCF text from the packed FORMAT columns.
 * These replace, for this path, what Service::genotype_sites does around the genotype_site loop
 * (src/service.cc:302-343 header/sample set-up, 388-418 ordered output) and the tail of genotype_site
 * (src/genotyper.cc:869-1003: alleles, INFO AF/AQ, bcf_trim_alleles, ID).
 */
#ifndef GLX_HOST_H
#define GLX_HOST_H
#include "glx.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct glxh_batch glxh_batch;
typedef struct glxh_out glxh_out;

/* Build one batch from text inputs.
 *  config_text : genotyper_config, line format of glx::genotyper_config::of_text
 *  preset      : if non-NULL, `glnexus_cli --config` preset name loaded first (src/cli_utils.cc:1236-1265);
 *                config_text (may be NULL/empty) then replaces it wholesale, as genotyper_config::of_yaml does
 *  sites_text  : unified sites, line format of glx::sites_of_text (0-based half-open)
 *  vcf_texts   : one single- or multi-sample gVCF per dataset (dataset name = file stem)
 *  contig      : keep only this contig's sites and records (NULL: all input must be on one contig)
 *  test_ingest_rules : apply the ingest-time record skips of the reference's test harness
 * The sample set is "<ALL>": every sample of every dataset, sorted by name (src/service.cc:307-309). */
int glxh_batch_from_text(const char* preset, const char* config_text, const char* sites_text, int n_datasets,
                         const char* const* dataset_names, const char* const* vcf_texts, const char* contig,
                         int test_ingest_rules, glxh_batch** ans, char* err, size_t errlen);
/* Multi-GPU: sites shard by contiguous range with no collective (they are independent; AF priors are inputs).
 * glxh_partition_sites writes n_parts+1 cut points of equal output weight; glxh_batch_slice builds the shard
 * [site_lo, site_hi) with every record its sites' query hulls can touch (records straddling a cut are replicated).
 * The shards' output columns concatenate, in order, to the unsharded batch's. */
int glxh_partition_sites(const glxh_batch* b, int n_parts, uint32_t* bounds);
/* The same cut over a cohort held as several consecutive segments (batches of consecutive genomic ranges): cut point p lies in
 * segment cut_seg[p] at site cut_site[p] of that segment (n_parts + 1 entries each; the last one is {n_segments, 0}). */
int glxh_partition_cohort(const glxh_batch* const* segments, uint32_t n_segments, int n_parts, uint32_t* cut_seg, uint32_t* cut_site);
int glxh_batch_slice(const glxh_batch* b, uint32_t site_lo, uint32_t site_hi, glxh_batch** ans, char* err, size_t errlen);
/* page-lock the batch's arrays (cudaHostRegister) so uploads run at PCIe rate and can overlap other batches */
int glxh_batch_pin(glxh_batch* b);
/* The compact wire form of the batch's record store (glx_wire, include/glx.h), built on first use and page-locked when the
 * batch is pinned; owned by the batch. Pass it to glx_upload_wire_on. GLX_E_NOT_IMPLEMENTED when the batch has no wire form
 * (multi-sample datasets). */
int glxh_batch_wire(glxh_batch* b, const glx_wire** wire);
/* test aid: rows of the shape table of wire forms built from now on (GLX_WF_SHAPES; default and maximum 255) — a small cap sends
 * the records of rarer shapes through the exception list */
void glxh_debug_wire_shape_cap(uint32_t cap);
/* test aid: threads pack_records uses for the batches built from now on (0 = chosen from the input's size: datasets are packed
 * independently and their pools laid end to end) */
void glxh_debug_pack_threads(int n);
void glxh_batch_free(glxh_batch* b);
const glx_config* glxh_config(const glxh_batch* b);
const glx_records* glxh_records(const glxh_batch* b);
const glx_sites* glxh_sites(const glxh_batch* b);
uint32_t glxh_n_fields(const glxh_batch* b);
int glxh_device_supported(const glxh_batch* b); /* 1 if the batch meets the device path's preconditions */
int glxh_has_fast_path(const glxh_batch* b);    /* 1 if it is single-sample datasets in output order (tiled kernels, wire form); else the multi-sample kernel */
/* st[0..5] = n_sites, n_samples, n_records, n_datasets, bytes of all input arrays, #variant (non-reference) records */
void glxh_batch_stats(const glxh_batch* b, uint64_t st[8]);

/* Host output buffers sized for the batch. `pinned` != 0 uses cudaHostAlloc. */
glxh_out* glxh_out_alloc(const glxh_batch* b, int pinned);
void glxh_out_free(glxh_out* o);
glx_out* glxh_out_struct(glxh_out* o);
/* raw access for comparisons: which = 0 gt, 1 rnc, 2 allele_used, 3 site_flags, 16+k field k */
int glxh_out_array(const glxh_out* o, int which, const void** ptr, uint64_t* nbytes);

/* Host buffers for the narrow download (glx_out16): 8-, 16- and 32-bit storage per field. which = 0 gt, 1 rnc,
 * 2 allele_used, 3 site_flags, 16+k field k in the narrow form in use (int8 if width[k] == 1, else int16), 48+k field k as int32. glxh_out16_widen expands into a glx_out
 * (int32 columns, htslib 32-bit sentinels) for consumers that want the wide form. */
typedef struct glxh_out16 glxh_out16;
glxh_out16* glxh_out16_alloc(const glxh_batch* b, int pinned);
void glxh_out16_free(glxh_out16* o);
glx_out16* glxh_out16_struct(glxh_out16* o);
int glxh_out16_array(const glxh_out16* o, int which, const void** ptr, uint64_t* nbytes);
int glxh_out16_widen(const glxh_out16* o, glx_out* wide);

/* pVCF text from packed columns. Caller frees with glxh_free_text. */
int glxh_assemble_vcf(const glxh_batch* b, const glx_out* out, char** text, char* err, size_t errlen);
void glxh_free_text(char* text);

/* ---- N1: BCF2 output. The records are encoded (and BGZF-compressed) on the device, see glx_encode_bcf_on in glx.h; the host
 * supplies the per-site strings / header dictionary (glx_bcf_meta) and writes the few bytes that are not records. ---- */
const glx_bcf_meta* glxh_bcf_meta(glxh_batch* b);
/* The uncompressed bytes that precede the first record of a BCF file: "BCF\2\2", l_text, header text, NUL; the header is
 * what prepare_bcf_header builds (src/service.cc:190-241) with htslib's IDX attributes. Caller frees with glxh_free_text. */
int glxh_bcf_prologue(const glxh_batch* b, char** bytes, uint64_t* n_bytes);
/* BGZF framing for those bytes: p[0..n) as consecutive BGZF blocks holding stored deflate blocks; and the 28-byte EOF block. */
int glxh_bgzf_stored(const void* p, uint64_t n, char** bytes, uint64_t* n_bytes);
int glxh_bgzf_eof(char** bytes, uint64_t* n_bytes);

/* genotyper_config preset -> text (for `--config NAME`); caller frees with glxh_free_text */
int glxh_preset_text(const char* name, char** text);

/* glnexus_cli's load_config (src/cli_utils.cc:1190-1265): the preset `name` (may be NULL) is loaded, `config_text` (may be
 * NULL/empty) then replaces it wholesale as genotyper_config::of_yaml does, and the command-line switches -P/--more-PL,
 * -S/--squeeze and -a/--trim-uncalled-alleles are OR-ed into the result (:1206-1208). *text = the resulting config in the
 * line format (pass it as config_text to the calls above); caller frees with glxh_free_text. Unknown preset: GLX_E_NOT_FOUND. */
int glxh_load_config(const char* name, const char* config_text, int more_PL, int squeeze, int trim_uncalled_alleles, char** text);

/* The upper seam: Service::genotype_sites (include/service.h:67-70, src/service.cc:302-428) on the device path. Sites are
 * genotyped in contiguous range batches of `batch_sites` (0 = 4096), two batches in flight, rows written to `filename`
 * ("-" = stdout) in input order as pVCF text; *abort != 0 stops between batches with GLX_E_ABORTED; the first failing
 * site's status is returned. stats = {sites submitted, rows written, batches, cells}. Inputs as glxh_batch_from_text. */
int glxh_genotype_sites(const char* preset, const char* config_text, const char* sites_text, int n_datasets, const char* const* dataset_names,
                        const char* const* vcf_texts, const char* filename, int device, uint32_t batch_sites, int test_ingest_rules,
                        const volatile int* abort, uint64_t stats[4], char* err, size_t errlen);

/* The streamed range-batch driver under glxh_genotype_sites, for callers that bring their own packed batches (one per contig
 * range, in genomic order): `depth` batches in flight on one device, results retired strictly in submission order.
 *   mode 0  packed int32 FORMAT columns (glx_out)            mode 2  finished, uncompressed BCF2 records
 *   mode 3  those records as BGZF blocks (DEFLATE + CRC-32 on the device), ready to append to the output file
 * A submitted batch must stay alive until `depth` later submits or glxh_stream_drain have retired it. With capture != 0 the
 * encoded bytes of every batch (modes 2, 3) are kept, concatenated in order, for glxh_stream_output; otherwise results are
 * dropped after their status check (measurement). stats = {batches, batches retired, sites, cells, records, H2D bytes, D2H
 * bytes, uncompressed BCF bytes}. */
typedef struct glxh_stream glxh_stream;
int glxh_stream_open(int device, int mode, int depth, int capture, glxh_stream** s, char* err, size_t errlen);
int glxh_stream_submit(glxh_stream* s, glxh_batch* b, char* err, size_t errlen);
/* upload through the wire form (glxh_batch_wire + glx_upload_wire_on) instead of the full-width columns */
void glxh_stream_use_wire(glxh_stream* s, int on);
int glxh_stream_drain(glxh_stream* s, char* err, size_t errlen);
void glxh_stream_stats(const glxh_stream* s, uint64_t stats[8]);
/* Measurement aid: glxh_stream_trace(s, 1) starts a timeline; every batch retired afterwards adds a row of 8 times in ms:
 * device {upload begins, upload done, genotyping kernels done, encoder done, download done}, host {submit entered, submit
 * returned, download queued}. glxh_stream_trace_rows copies up to `cap` rows and returns how many there are. */
void glxh_stream_trace(glxh_stream* s, int on);
uint32_t glxh_stream_trace_rows(const glxh_stream* s, float* rows, uint32_t cap);
int glxh_stream_output(const glxh_stream* s, const char** bytes, uint64_t* n_bytes);
void glxh_stream_close(glxh_stream* s);

/* Seeded synthetic cohort in record-store form (BASELINE.json configs[1..4] shapes). See DESIGN.md §synthetic. */
typedef struct glxh_synth_params {
    uint32_t n_samples, n_sites;
    uint64_t seed;
    int32_t region_beg, region_len; /* sites are drawn uniformly without replacement inside [beg, beg+len) */
    uint32_t max_alt;               /* ALT alleles per site: 1..max_alt */
    float frac_multiallelic;        /* fraction of sites with >1 ALT (when max_alt>1) */
    float frac_indel;               /* fraction of ALT alleles that are indels */
    float af_max;                   /* site AF ~ 1/x on [1/(2N), af_max] */
    float band_mean_len;            /* mean reference-band length (geometric) */
    float frac_uncovered;           /* fraction of band starts replaced by a coverage gap */
    float frac_two_record_cells;    /* extra overlapping/0-0 variant records (config 3 stress) */
    float frac_lost_allele;         /* carrier records whose ALT is absent from the unified site */
    float frac_homalt;              /* carriers called hom-ALT (forces revise_genotypes) */
    uint8_t gatk_style;             /* <NON_REF> + SB field, no PL on bands */
    uint8_t sites_only;             /* unified sites without any records (cheap: for partitioning a cohort of several segments) */
    uint8_t _pad[2];
} glxh_synth_params;
int glxh_synth_batch(const char* preset, const glxh_synth_params* p, glxh_batch** ans, char* err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
