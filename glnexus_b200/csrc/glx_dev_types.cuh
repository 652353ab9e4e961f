// glx_dev_types.cuh — device-side views of the C ABI structs (include/glx.h) and the flattened config.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/glx.h"

namespace glxd {

#define GLX_RF_DEV_LAST 0x80  // device copy of glx_records.flags only: last record of its dataset (set by glx_upload)

// name-based censoring exemptions of genotype_site (src/genotyper.cc:796-821) and the squeeze path
// (src/genotyper_utils.h:943-947): decided by FORMAT *name*, not helper class.
enum { NF_DP = 1, NF_GQ = 2, NF_FT = 4 };

struct DFld {
    uint8_t kind, type, number, combi;
    uint8_t default_zero, ignore_non_variants, n_cols, name_flags;
    int32_t count;
    int16_t cols[GLX_MAX_ORIG];
};

struct DCfg {
    uint8_t revise, allow_partial, more_PL, squeeze;
    uint8_t top_two, xatlas, calibration, has_string_field;
    uint32_t required_dp;
    float snv_cal, indel_cal;
    double neg10_log10e, ln10;
    int16_t col_ref_dp, col_allele_dp, col_pl, col_gq;
    uint32_t n_fields;
    uint32_t _pad;
    DFld f[GLX_MAX_FIELDS];
};

struct DRecs {
    uint64_t n_records;
    const uint64_t* ds_rec_off;
    const int32_t *beg, *end, *maxend;
    const float* qual;
    const uint8_t *n_allele, *flags;
    const uint32_t *filt, *gt, *allele_off;
    const int32_t* col_val;
    const uint16_t* col_len;
    const int32_t* pool;
    const uint32_t* al_len;
    const uint8_t* al_flags;
    const uint64_t* al_prefix;
    const uint32_t* al_str;
    const char* al_pool;
};

struct DSites {
    uint32_t n_sites, n_samples;
    const int32_t *beg, *end, *qbeg, *qend;
    const uint8_t *n_allele, *monoallelic;
    const uint32_t* allele_off;
    const uint8_t* allele_is_indel;
    const float *hom_log_prior, *lost_log_prior;
    const uint32_t* unif_off;
    const int32_t *unif_beg, *unif_end;
    const uint8_t* unif_to;
    const uint32_t* unif_len;
    const uint64_t* unif_prefix;
    const uint32_t* unif_str;
    const char* unif_pool;
    const uint16_t* fld_cnt[GLX_MAX_FIELDS];
    const uint64_t* fld_off[GLX_MAX_FIELDS];
    const uint32_t* slot_base;  // [S][n_fields+1] prefix sums of fld_cnt over fields (scratch layout of the general path)
    // packed copies built by glx_upload (build_site_rows) so that a thread gets a site / a unification entry in two 16-byte loads:
    const int4* site_rows;  // [S][2]: {beg, end, qbeg, qend} {n_allele | monoallelic << 8, allele_off, unif_off, unif_off[s+1]}
    const int4* unif_rows;  // [n_unif][2]: {beg, end, len, to} {prefix lo, prefix hi, str, 0}
};

struct DOut {
    int8_t* gt;
    uint8_t* rnc;
    int32_t* fld[GLX_MAX_FIELDS];
    uint32_t* allele_used;
    uint8_t* site_flags;
    unsigned long long* error;  // min over failing cells of (cell << 8 | -status), cell = site * n_samples + sample
};

// flatten glx_config for the kernels (host side)
inline DCfg make_dcfg(const glx_config& c) {
    DCfg d;
    memset(&d, 0, sizeof d);
    d.revise = c.revise_genotypes;
    d.allow_partial = c.allow_partial_data;
    d.more_PL = c.more_PL;
    d.squeeze = c.squeeze;
    d.top_two = c.top_two_half_calls;
    d.xatlas = c.xatlas_depth_mode;
    d.calibration = (c.snv_prior_calibration != 1.0 || c.indel_prior_calibration != 1.0) ? 1 : 0;  // genotyper.cc:156
    d.required_dp = c.required_dp;
    d.snv_cal = c.snv_prior_calibration;
    d.indel_cal = c.indel_prior_calibration;
    d.neg10_log10e = c.neg10_log10e;
    d.ln10 = c.ln10;
    d.col_ref_dp = c.col_ref_dp;
    d.col_allele_dp = c.col_allele_dp;
    d.col_pl = c.col_pl;
    d.col_gq = c.col_gq;
    d.n_fields = c.n_fields;
    for (uint32_t k = 0; k < c.n_fields; k++) {
        const glx_field& f = c.fields[k];
        DFld& o = d.f[k];
        o.kind = f.kind;
        o.type = f.type;
        o.number = f.number;
        o.combi = f.combi;
        o.default_zero = f.default_zero;
        o.ignore_non_variants = f.ignore_non_variants;
        o.n_cols = f.n_cols;
        o.count = f.count;
        o.name_flags = (strcmp(f.name, "DP") == 0 ? NF_DP : 0) | (strcmp(f.name, "GQ") == 0 ? NF_GQ : 0) | (strcmp(f.name, "FT") == 0 ? NF_FT : 0);
        for (int i = 0; i < GLX_MAX_ORIG; i++) o.cols[i] = f.cols[i];
        if (f.kind == GLX_FK_STRING) d.has_string_field = 1;
    }
    return d;
}

// host side: the packed site / unification rows of DSites
inline void build_site_rows(const glx_sites& s, std::vector<int4>& site_rows, std::vector<int4>& unif_rows) {
    site_rows.resize((size_t)s.n_sites * 2 + 2);
    unif_rows.resize((size_t)s.n_unif * 2 + 2);
    for (uint32_t i = 0; i < s.n_sites; i++) {
        site_rows[2 * (size_t)i] = int4{s.beg[i], s.end[i], s.qbeg[i], s.qend[i]};
        site_rows[2 * (size_t)i + 1] = int4{(int)s.n_allele[i] | (s.monoallelic[i] ? 256 : 0), (int)s.allele_off[i], (int)s.unif_off[i], (int)s.unif_off[i + 1]};
    }
    for (uint64_t u = 0; u < s.n_unif; u++) {
        unif_rows[2 * u] = int4{s.unif_beg[u], s.unif_end[u], (int)s.unif_len[u], (int)s.unif_to[u]};
        unif_rows[2 * u + 1] = int4{(int)(uint32_t)s.unif_prefix[u], (int)(uint32_t)(s.unif_prefix[u] >> 32), (int)s.unif_str[u], 0};
    }
}

}  // namespace glxd
