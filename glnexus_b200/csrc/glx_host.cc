// glx_host.cc — host pipeline around the device genotyper: text decode, columnar packing, output assembly.
// See glx_host.hpp for the reference interfaces mirrored here.
#include "glx_host.hpp"

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace glx {

std::string Status::str() const {
    const char* nm = "Failure";
    switch (code_) {
        case StatusCode::OK: nm = "OK"; break;
        case StatusCode::INVALID: nm = "Invalid"; break;
        case StatusCode::NOT_FOUND: nm = "NotFound"; break;
        case StatusCode::EXISTS: nm = "Exists"; break;
        case StatusCode::IO_ERROR: nm = "IOError"; break;
        case StatusCode::NOT_IMPLEMENTED: nm = "NotImplemented"; break;
        case StatusCode::ABORTED: nm = "Aborted"; break;
        default: break;
    }
    std::string ans(nm);
    if (!msg_.empty()) ans += ": " + msg_;
    return ans;
}

// src/types.cc:26-43
bool is_dna(const std::string& s) {
    if (s.empty()) return false;
    for (char c : s)
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return false;
    return true;
}

// src/types.cc:44-78
bool is_iupac_nucleotides(const std::string& s) {
    if (s.empty()) return false;
    for (char c : s)
        if (!strchr("ACGTURYSWKMBDHVN", c)) return false;
    return true;
}

// src/types.cc:970-973
bool is_symbolic_allele(const std::string& s) { return s.size() > 1 && s.front() == '<' && s.back() == '>'; }

// src/types.cc:402-407
void unified_site::fill_implicit_unification() {
    for (int i = 0; i < (int)alleles.size(); i++) {
        unification[allele(pos, alleles[i].dna)] = i;
        unification[alleles[i].normalized] = i;
    }
}

// ---------------------------------------------------------------------------------------------
// small text helpers
// ---------------------------------------------------------------------------------------------
static std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> ans;
    size_t i = 0;
    while (true) {
        size_t j = s.find(sep, i);
        if (j == std::string::npos) {
            ans.push_back(s.substr(i));
            break;
        }
        ans.push_back(s.substr(i, j - i));
        i = j + 1;
    }
    return ans;
}

static std::vector<std::string> lines_of(const std::string& txt) {
    std::vector<std::string> ans = split(txt, '\n');
    for (auto& l : ans)
        if (!l.empty() && l.back() == '\r') l.pop_back();
    return ans;
}

static bool parse_int(const std::string& s, long& v) {
    if (s.empty()) return false;
    char* e = nullptr;
    v = strtol(s.c_str(), &e, 10);
    return e && *e == 0;
}

static bool parse_float(const std::string& s, float& v) {
    if (s.empty()) return false;
    char* e = nullptr;
    v = strtof(s.c_str(), &e);
    return e && *e == 0;
}

static bool parse_bool(const std::string& s) { return s == "1" || s == "true" || s == "True" || s == "yes"; }

static inline int32_t float_bits(float f) {
    int32_t b;
    memcpy(&b, &f, 4);
    return b;
}
static inline float bits_float(int32_t b) {
    float f;
    memcpy(&f, &b, 4);
    return f;
}

// ---------------------------------------------------------------------------------------------
// genotyper_config <-> text. One `key value` per line; liftover fields are tab-separated:
// field <name> <orig,names> <format|info> <int|float|string> <basic|alt|genotype|alleles> <count>
//       <missing|zero> <min|max|missing|semicolon> <ignore_non_variants> <description>
// Field semantics and validation follow genotyper_config::of_yaml / retained_format_field::of_yaml
// (src/types.cc:693-812, 854-967).
// ---------------------------------------------------------------------------------------------
Status genotyper_config::of_text(const std::string& txt, genotyper_config& ans) {
    ans = genotyper_config();
    for (const auto& line : lines_of(txt)) {
        if (line.empty() || line[0] == '#') continue;
        if (line.compare(0, 6, "field\t") == 0) {
            auto t = split(line, '\t');
            if (t.size() < 11) return Status::Invalid("genotyper_config::of_text: malformed field line", line);
            retained_format_field f;
            f.name = t[1];
            if (f.name.empty()) return Status::Invalid("retained_format_field: missing name");
            f.orig_names = split(t[2], ',');
            if (t[2].empty() || f.orig_names.empty()) return Status::Invalid("retained_format_field: missing orig_names");
            if (t[3] == "info") f.from = RetainedFieldFrom::INFO;
            else if (t[3] == "format") f.from = RetainedFieldFrom::FORMAT;
            else return Status::Invalid("retained_format_field: invalid from");
            if (t[4] == "int") f.type = RetainedFieldType::INT;
            else if (t[4] == "float") f.type = RetainedFieldType::FLOAT;
            else if (t[4] == "string") f.type = RetainedFieldType::STRING;
            else return Status::Invalid("retained_format_field: invalid type");
            if (t[5] == "basic") f.number = RetainedFieldNumber::BASIC;
            else if (t[5] == "alt") f.number = RetainedFieldNumber::ALT;
            else if (t[5] == "genotype") f.number = RetainedFieldNumber::GENOTYPE;
            else if (t[5] == "alleles") f.number = RetainedFieldNumber::ALLELES;
            else return Status::Invalid("retained_format_field: invalid number");
            long c = 0;
            if (!parse_int(t[6], c)) return Status::Invalid("retained_format_field: missing count");
            f.count = f.number == RetainedFieldNumber::BASIC ? (int)c : 0;
            if (t[7] == "zero") f.default_type = DefaultValueFiller::ZERO;
            else if (t[7] == "missing") f.default_type = DefaultValueFiller::MISSING;
            else return Status::Invalid("retained_format_field: invalid default_type value");
            if (t[8] == "min") f.combi_method = FieldCombinationMethod::MIN;
            else if (t[8] == "max") f.combi_method = FieldCombinationMethod::MAX;
            else if (t[8] == "missing") f.combi_method = FieldCombinationMethod::MISSING;
            else if (t[8] == "semicolon") f.combi_method = FieldCombinationMethod::SEMICOLON;
            else return Status::Invalid("retained_format_field: invalid combi_method");
            f.ignore_non_variants = parse_bool(t[9]);
            f.description = t[10];
            if (f.description.empty()) return Status::Invalid("retained_format_field: missing description");
            ans.liftover_fields.push_back(f);
            continue;
        }
        size_t sp = line.find_first_of(" \t");
        if (sp == std::string::npos) return Status::Invalid("genotyper_config::of_text: malformed line", line);
        std::string k = line.substr(0, sp), v = line.substr(sp + 1);
        float fv = 0;
        long iv = 0;
        if (k == "revise_genotypes") ans.revise_genotypes = parse_bool(v);
        else if (k == "min_assumed_allele_frequency") {
            if (!parse_float(v, fv) || !(fv >= 0 && fv <= 1.0)) return Status::Invalid("genotyper_config: invalid min_assumed_allele_frequency");
            ans.min_assumed_allele_frequency = fv;
        } else if (k == "snv_prior_calibration") {
            if (!parse_float(v, fv) || !(fv >= 0)) return Status::Invalid("genotyper_config: invalid snv_prior_calibration");
            ans.snv_prior_calibration = fv;
        } else if (k == "indel_prior_calibration") {
            if (!parse_float(v, fv) || !(fv >= 0)) return Status::Invalid("genotyper_config: invalid indel_prior_calibration");
            ans.indel_prior_calibration = fv;
        } else if (k == "required_dp") {
            if (!parse_int(v, iv) || iv < 0) return Status::Invalid("genotyper_config: invalid required_dp");
            ans.required_dp = (size_t)iv;
        } else if (k == "allow_partial_data") ans.allow_partial_data = parse_bool(v);
        else if (k == "allele_dp_format") ans.allele_dp_format = v;
        else if (k == "ref_dp_format") ans.ref_dp_format = v;
        else if (k == "output_residuals") ans.output_residuals = parse_bool(v);
        else if (k == "more_PL") ans.more_PL = parse_bool(v);
        else if (k == "squeeze") ans.squeeze = parse_bool(v);
        else if (k == "trim_uncalled_alleles") ans.trim_uncalled_alleles = parse_bool(v);
        else if (k == "top_two_half_calls") ans.top_two_half_calls = parse_bool(v);
        else if (k == "output_format") {
            if (v == "BCF") ans.output_format = GLnexusOutputFormat::BCF;
            else if (v == "VCF") ans.output_format = GLnexusOutputFormat::VCF;
            else return Status::Invalid("genotyper_config: invalid output_format. Must be one of {BCF, VCF}.");
        } else
            return Status::Invalid("genotyper_config::of_text: unknown key", k);
    }
    return Status::OK();
}

std::string genotyper_config::text() const {
    std::ostringstream o;
    char buf[64];
    o << "revise_genotypes " << (revise_genotypes ? 1 : 0) << "\n";
    snprintf(buf, sizeof buf, "%.9g", min_assumed_allele_frequency);
    o << "min_assumed_allele_frequency " << buf << "\n";
    snprintf(buf, sizeof buf, "%.9g", snv_prior_calibration);
    o << "snv_prior_calibration " << buf << "\n";
    snprintf(buf, sizeof buf, "%.9g", indel_prior_calibration);
    o << "indel_prior_calibration " << buf << "\n";
    o << "required_dp " << required_dp << "\n";
    o << "allow_partial_data " << (allow_partial_data ? 1 : 0) << "\n";
    o << "allele_dp_format " << allele_dp_format << "\n";
    o << "ref_dp_format " << ref_dp_format << "\n";
    o << "output_residuals " << (output_residuals ? 1 : 0) << "\n";
    o << "more_PL " << (more_PL ? 1 : 0) << "\n";
    o << "squeeze " << (squeeze ? 1 : 0) << "\n";
    o << "trim_uncalled_alleles " << (trim_uncalled_alleles ? 1 : 0) << "\n";
    o << "top_two_half_calls " << (top_two_half_calls ? 1 : 0) << "\n";
    o << "output_format " << (output_format == GLnexusOutputFormat::VCF ? "VCF" : "BCF") << "\n";
    static const char* ty[] = {"int", "float", "string"};
    static const char* nu[] = {"basic", "alt", "genotype", "alleles"};
    static const char* co[] = {"min", "max", "missing", "semicolon"};
    for (const auto& f : liftover_fields) {
        o << "field\t" << f.name << "\t";
        for (size_t i = 0; i < f.orig_names.size(); i++) o << (i ? "," : "") << f.orig_names[i];
        o << "\t" << (f.from == RetainedFieldFrom::INFO ? "info" : "format") << "\t" << ty[(int)f.type] << "\t"
          << nu[(int)f.number] << "\t" << f.count << "\t"
          << (f.default_type == DefaultValueFiller::ZERO ? "zero" : "missing") << "\t" << co[(int)f.combi_method]
          << "\t" << (f.ignore_non_variants ? 1 : 0) << "\t" << f.description << "\n";
    }
    return o.str();
}

// ---------------------------------------------------------------------------------------------
// unified sites <-> text (0-based half-open coordinates):
//   site   <contig> <beg> <end> <qual> <monoallelic> <lost_allele_frequency>
//   allele <dna> <norm_beg> <norm_end> <norm_dna> <quality> <frequency|nan>
//   unif   <beg> <end> <dna> <to>
// Semantics of unified_site::of_yaml (src/types.cc:409-511): implicit unification entries first,
// explicit ones must not contradict them.
// ---------------------------------------------------------------------------------------------
Status sites_of_text(const std::string& txt, const std::vector<std::pair<std::string, size_t>>& contigs,
                     std::vector<unified_site>& ans) {
    ans.clear();
    std::vector<std::tuple<range, std::string, int>> pending;  // explicit unification entries of the current site
    auto close_site = [&]() -> Status {
        if (ans.empty()) return Status::OK();
        unified_site& us = ans.back();
        if (us.alleles.size() < 2) return Status::Invalid("unified_site_of_text: not enough alleles");
        us.unification.clear();
        us.fill_implicit_unification();
        for (const auto& p : pending) {
            allele al(std::get<0>(p), std::get<1>(p));
            int to = std::get<2>(p);
            if (to < 0 || to >= (int)us.alleles.size()) return Status::Invalid("unified_site_of_text: invalid 'to' field in unification entry");
            auto it = us.unification.find(al);
            if (it != us.unification.end() && it->second != to) return Status::Invalid("unified_site_of_text: inconsistent unification entries");
            us.unification[al] = to;
        }
        pending.clear();
        return Status::OK();
    };
    Status s;
    for (const auto& line : lines_of(txt)) {
        if (line.empty() || line[0] == '#') continue;
        auto t = split(line, '\t');
        if (t[0] == "site") {
            GLX_S(close_site());
            if (t.size() < 7) return Status::Invalid("unified_site_of_text: malformed site line", line);
            int rid = -1;
            for (size_t i = 0; i < contigs.size(); i++)
                if (contigs[i].first == t[1]) rid = (int)i;
            if (rid < 0) return Status::Invalid("range_of_text: unknown contig", t[1]);
            long b, e, q;
            float laf;
            if (!parse_int(t[2], b) || !parse_int(t[3], e) || b < 0 || e < b) return Status::Invalid("range_of_text: invalid beg/end coordinates");
            if (!parse_int(t[4], q) || q < 0) return Status::Invalid("unified_site_of_text: invalid 'qual' field");
            if (!parse_float(t[6], laf) || !(laf >= 0.0f && laf <= 1.0f)) return Status::Invalid("unified_site_of_text: invalid 'lost_allele_frequency' field");
            unified_site us(range(rid, (int)b, (int)e));
            us.qual = (int)q;
            us.monoallelic = parse_bool(t[5]);
            us.lost_allele_frequency = laf;
            ans.push_back(us);
        } else if (t[0] == "allele") {
            if (ans.empty() || t.size() < 7) return Status::Invalid("unified_site_of_text: malformed allele line", line);
            unified_site& us = ans.back();
            long nb, ne, q;
            if (!parse_int(t[2], nb) || !parse_int(t[3], ne) || !parse_int(t[5], q)) return Status::Invalid("unified_site_of_text: invalid allele");
            if (!is_iupac_nucleotides(t[1]) || !is_iupac_nucleotides(t[4])) return Status::Invalid("allele(): invalid DNA", t[1]);
            unified_allele ua(us.pos, t[1]);
            ua.normalized = allele(range(us.pos.rid, (int)nb, (int)ne), t[4]);
            ua.quality = (int)q;
            if (t[6] == "nan" || t[6] == "NaN" || t[6] == ".") ua.frequency = NAN;
            else {
                float f;
                if (!parse_float(t[6], f) || !(f >= 0.0f && f <= 1.0f)) return Status::Invalid("unified_site_of_text: invalid 'frequency' field in allele info");
                ua.frequency = f;
            }
            us.alleles.push_back(ua);
        } else if (t[0] == "unif") {
            if (ans.empty() || t.size() < 5) return Status::Invalid("unified_site_of_text: malformed unif line", line);
            long b, e, to;
            if (!parse_int(t[1], b) || !parse_int(t[2], e) || !parse_int(t[4], to) || b < 0 || e < b) return Status::Invalid("unified_site_of_text: invalid unification entry");
            if (t[3].empty()) return Status::Invalid("unified_site_of_text: empty 'dna' in unification entry");
            if (!is_iupac_nucleotides(t[3])) return Status::Invalid("allele(): invalid DNA", t[3]);
            pending.emplace_back(range(ans.back().pos.rid, (int)b, (int)e), t[3], (int)to);
        } else {
            return Status::Invalid("unified_site_of_text: unknown line", line);
        }
    }
    GLX_S(close_site());
    return Status::OK();
}

// ---------------------------------------------------------------------------------------------
// VCF text reader. Value semantics are those the reference sees through htslib 1.9 accessors:
// int32 with INT32_MIN missing / INT32_MIN+1 vector_end, float missing bits 0x7F800001,
// GT ints ((allele+1)<<1 | phased), rlen from INFO/END (include/types.h:152-157).
// ---------------------------------------------------------------------------------------------
static uint8_t header_type_of(const std::string& line) {
    size_t p = line.find("Type=");
    if (p == std::string::npos) return GLX_TYPE_STRING;
    std::string t = line.substr(p + 5, line.find_first_of(",>", p + 5) - (p + 5));
    if (t == "Integer") return GLX_TYPE_INT;
    if (t == "Float") return GLX_TYPE_FLOAT;
    return GLX_TYPE_STRING;
}

static std::string header_attr(const std::string& line, const std::string& key) {
    size_t p = line.find(key + "=");
    if (p == std::string::npos) return "";
    p += key.size() + 1;
    return line.substr(p, line.find_first_of(",>", p) - p);
}

static void parse_numeric_vector(const std::string& txt, uint8_t type, std::vector<int32_t>& out) {
    out.clear();
    for (const auto& tok : split(txt, ',')) {
        if (type == GLX_TYPE_INT) {
            long v;
            if (tok == "." || !parse_int(tok, v)) out.push_back(GLX_INT_MISSING);
            else out.push_back((int32_t)v);
        } else {
            float f;
            if (tok == "." || !parse_float(tok, f)) out.push_back((int32_t)GLX_FLOAT_MISSING_BITS);
            else out.push_back(float_bits(f));
        }
    }
}

Status reconcile_contigs(const VcfHeaderInfo& db, const VcfHeaderInfo& h, GvcfDataset& ds) {
    const char* msg = "Incompatible gVCF. The reference contigs must match the database configuration exactly.";
    if (db.n_declared_contigs && h.n_declared_contigs) {
        if (db.n_declared_contigs != h.n_declared_contigs) return Status::Invalid(msg, ds.name);
        for (size_t i = 0; i < h.n_declared_contigs; i++)
            if (h.contigs[i].first != db.contigs[i].first || (h.contigs[i].second > 0 && db.contigs[i].second > 0 && h.contigs[i].second != db.contigs[i].second))
                return Status::Invalid(msg, ds.name);
    }
    // rid of every record by contig NAME in the shared table (identity when the declared tables matched)
    std::vector<int> remap(h.contigs.size(), -1);
    bool identity = true;
    for (size_t i = 0; i < h.contigs.size(); i++) {
        for (size_t j = 0; j < db.contigs.size(); j++)
            if (db.contigs[j].first == h.contigs[i].first) remap[i] = (int)j;
        if (remap[i] != (int)i) identity = false;
    }
    if (identity) return Status::OK();
    for (auto& r : ds.records) {
        if (r.rid < 0 || r.rid >= (int)remap.size() || remap[r.rid] < 0) return Status::Invalid(msg, ds.name);
        r.rid = remap[r.rid];
    }
    return Status::OK();
}

Status read_vcf_text(const std::string& text, const std::string& dataset_name, bool test_ingest_rules,
                     VcfHeaderInfo& hdr, GvcfDataset& ans) {
    ans = GvcfDataset();
    ans.name = dataset_name;
    hdr = VcfHeaderInfo();
    bool have_cols = false;
    int prev_rid = -1, prev_pos = -1;
    for (const auto& line : lines_of(text)) {
        if (line.empty()) continue;
        if (line.compare(0, 2, "##") == 0) {
            if (line.compare(0, 9, "##FORMAT=") == 0) hdr.format_types[header_attr(line, "ID")] = header_type_of(line);
            else if (line.compare(0, 7, "##INFO=") == 0) hdr.info_types[header_attr(line, "ID")] = header_type_of(line);
            else if (line.compare(0, 9, "##contig=") == 0) {
                std::string len = header_attr(line, "length");
                long l = 0;
                parse_int(len, l);
                hdr.contigs.emplace_back(header_attr(line, "ID"), (size_t)l);
                hdr.n_declared_contigs = hdr.contigs.size();
            }
            continue;
        }
        if (line[0] == '#') {
            auto t = split(line, '\t');
            if (t.size() < 8) return Status::IOError("bcf_hdr_read failed", dataset_name);
            for (size_t i = 9; i < t.size(); i++)
                if (!t[i].empty()) ans.samples.push_back(t[i]);
            have_cols = true;
            continue;
        }
        if (!have_cols) return Status::IOError("bcf_hdr_read failed", dataset_name);
        auto t = split(line, '\t');
        if (t.size() < 8) return Status::IOError("bcf_read", dataset_name);
        GvcfRecord r;
        r.rid = -1;
        for (size_t i = 0; i < hdr.contigs.size(); i++)
            if (hdr.contigs[i].first == t[0]) r.rid = (int)i;
        if (r.rid < 0) {
            hdr.contigs.emplace_back(t[0], 0);
            r.rid = (int)hdr.contigs.size() - 1;
        }
        long pos1;
        if (!parse_int(t[1], pos1) || pos1 < 1) return Status::IOError("bcf_read", dataset_name);
        r.pos = (int)pos1 - 1;
        r.alleles.push_back(t[3]);
        if (t[4] != "." && !t[4].empty())
            for (const auto& a : split(t[4], ',')) r.alleles.push_back(a);
        if (t[5] != ".") {
            float q;
            if (parse_float(t[5], q)) {
                r.qual = q;
                r.qual_missing = false;
            }
        }
        if (r.qual_missing) r.qual = bits_float((int32_t)GLX_FLOAT_MISSING_BITS);
        if (t[6] != "." && !t[6].empty()) r.filters = split(t[6], ';');
        r.rlen = (int)t[3].size();
        if (t[7] != "." && !t[7].empty()) {
            for (const auto& kv : split(t[7], ';')) {
                size_t eq = kv.find('=');
                std::string k = eq == std::string::npos ? kv : kv.substr(0, eq);
                std::string v = eq == std::string::npos ? "" : kv.substr(eq + 1);
                auto it = hdr.info_types.find(k);
                uint8_t ty = it == hdr.info_types.end() ? GLX_TYPE_STRING : it->second;
                if (k == "END") {
                    long e;
                    if (parse_int(v, e)) r.rlen = (int)e - r.pos;
                    ty = GLX_TYPE_INT;
                }
                FieldValues fv;
                fv.id = k;
                fv.type = ty;
                if (ty != GLX_TYPE_STRING && eq != std::string::npos) {
                    parse_numeric_vector(v, ty, fv.v);
                    fv.per_sample = (int)fv.v.size();
                }
                r.info.push_back(std::move(fv));
            }
        }
        r.n_sample = (int)ans.samples.size();
        if (t.size() > 8 && r.n_sample > 0) {
            auto keys = split(t[8], ':');
            std::vector<std::vector<std::string>> sv(r.n_sample);
            for (int i = 0; i < r.n_sample; i++)
                if (9 + i < (int)t.size()) sv[i] = split(t[9 + i], ':');
            for (size_t k = 0; k < keys.size(); k++) {
                if (keys[k] == "GT") {
                    std::vector<std::vector<int32_t>> g(r.n_sample);
                    size_t mx = 0;
                    for (int i = 0; i < r.n_sample; i++) {
                        std::string gs = k < sv[i].size() ? sv[i][k] : ".";
                        size_t p = 0;
                        bool phased = false;
                        while (p <= gs.size()) {
                            size_t q = gs.find_first_of("/|", p);
                            std::string tok = gs.substr(p, q == std::string::npos ? std::string::npos : q - p);
                            long a;
                            int32_t val = (tok == "." || !parse_int(tok, a)) ? 0 : (int32_t)((a + 1) << 1);
                            g[i].push_back(val | (phased ? 1 : 0));
                            if (q == std::string::npos) break;
                            phased = gs[q] == '|';
                            p = q + 1;
                        }
                        mx = std::max(mx, g[i].size());
                    }
                    r.n_gt = (int)mx;
                    for (int i = 0; i < r.n_sample; i++) {
                        for (size_t j = 0; j < mx; j++) r.gt.push_back(j < g[i].size() ? g[i][j] : GLX_INT_VECTOR_END);
                    }
                    continue;
                }
                auto it = hdr.format_types.find(keys[k]);
                FieldValues fv;
                fv.id = keys[k];
                fv.type = it == hdr.format_types.end() ? GLX_TYPE_STRING : it->second;
                if (fv.type != GLX_TYPE_STRING) {
                    std::vector<std::vector<int32_t>> vals(r.n_sample);
                    size_t mx = 1;
                    for (int i = 0; i < r.n_sample; i++) {
                        parse_numeric_vector(k < sv[i].size() ? sv[i][k] : ".", fv.type, vals[i]);
                        mx = std::max(mx, vals[i].size());
                    }
                    fv.per_sample = (int)mx;
                    int32_t vend = fv.type == GLX_TYPE_INT ? GLX_INT_VECTOR_END : (int32_t)GLX_FLOAT_VECTOR_END_BITS;
                    for (int i = 0; i < r.n_sample; i++)
                        for (size_t j = 0; j < mx; j++) fv.v.push_back(j < vals[i].size() ? vals[i][j] : vend);
                }
                r.format.push_back(std::move(fv));
            }
        }
        if (test_ingest_rules) {
            // test/utils.cc:75-77 and src/BCF_utils.cc:26-66 (skip_ingestion)
            bool skip = std::find(r.filters.begin(), r.filters.end(), "VRFromDeletion") != r.filters.end();
            if (r.rlen >= 100000) skip = true;
            if (!skip && r.n_sample == 1) {
                int32_t vr = GLX_INT_MISSING, rr = GLX_INT_MISSING;
                for (const auto& f : r.format) {
                    if (f.id == "VR" && f.type == GLX_TYPE_INT && f.v.size() == 1) vr = f.v[0];
                    if (f.id == "RR" && f.type == GLX_TYPE_INT && f.v.size() == 1) rr = f.v[0];
                }
                if (vr != GLX_INT_MISSING && rr != GLX_INT_MISSING && vr + rr >= 65536) skip = true;
            }
            if (skip) continue;
            if (prev_rid == r.rid && prev_pos > r.pos) return Status::Invalid("gVCF records are out-of-order ", dataset_name);
        }
        prev_rid = r.rid;
        prev_pos = r.pos;
        ans.records.push_back(std::move(r));
    }
    if (!have_cols) return Status::IOError("bcf_hdr_read failed", dataset_name);
    return Status::OK();
}

// ---------------------------------------------------------------------------------------------
// genotyper_config -> glx_config (helper selection = setup_format_helpers, src/genotyper_utils.h:865-931)
// ---------------------------------------------------------------------------------------------
static int get_col(glx_config& c, const std::string& name, bool from_info, uint8_t type) {
    for (uint32_t i = 0; i < c.n_cols; i++)
        if (name == c.cols[i].name && c.cols[i].from_info == (from_info ? 1 : 0) && c.cols[i].type == type) return (int)i;
    if (c.n_cols >= GLX_MAX_COLS || name.size() >= sizeof(c.cols[0].name)) return -1;
    glx_column& col = c.cols[c.n_cols];
    memset(&col, 0, sizeof col);
    strncpy(col.name, name.c_str(), sizeof(col.name) - 1);
    col.from_info = from_info ? 1 : 0;
    col.type = type;
    return (int)c.n_cols++;
}

Status make_glx_config(const genotyper_config& cfg, glx_config& c) {
    memset(&c, 0, sizeof c);
    c.abi_version = GLX_ABI_VERSION;
    c.revise_genotypes = cfg.revise_genotypes;
    c.allow_partial_data = cfg.allow_partial_data;
    c.more_PL = cfg.more_PL;
    c.squeeze = cfg.squeeze;
    c.trim_uncalled_alleles = cfg.trim_uncalled_alleles;
    c.top_two_half_calls = cfg.top_two_half_calls;
    c.xatlas_depth_mode = (cfg.ref_dp_format == "RR" && cfg.allele_dp_format == "VR");  // genotyper_utils.h:1118-1123
    c.required_dp = cfg.required_dp > 0x7fffffffu ? 0x7fffffffu : (uint32_t)cfg.required_dp;
    c.min_assumed_allele_frequency = cfg.min_assumed_allele_frequency;
    c.snv_prior_calibration = cfg.snv_prior_calibration;
    c.indel_prior_calibration = cfg.indel_prior_calibration;
    c.neg10_log10e = -10.0 * log10(exp(1.0));
    c.ln10 = log(10.0);
    int a = get_col(c, cfg.ref_dp_format, false, GLX_TYPE_INT), b = get_col(c, cfg.allele_dp_format, false, GLX_TYPE_INT),
        p = get_col(c, "PL", false, GLX_TYPE_INT), g = get_col(c, "GQ", false, GLX_TYPE_INT);
    if (a < 0 || b < 0 || p < 0 || g < 0) return Status::NotImplemented("make_glx_config: too many source columns");
    c.col_ref_dp = (int16_t)a;
    c.col_allele_dp = (int16_t)b;
    c.col_pl = (int16_t)p;
    c.col_gq = (int16_t)g;
    if (cfg.liftover_fields.size() > GLX_MAX_FIELDS) return Status::NotImplemented("make_glx_config: too many liftover_fields");
    for (const auto& f : cfg.liftover_fields) {
        glx_field& o = c.fields[c.n_fields++];
        if (f.name.size() >= sizeof(o.name)) return Status::NotImplemented("make_glx_config: field name too long", f.name);
        strncpy(o.name, f.name.c_str(), sizeof(o.name) - 1);
        o.type = (uint8_t)f.type;
        o.number = (uint8_t)f.number;
        o.combi = (uint8_t)f.combi_method;
        o.default_zero = f.default_type == DefaultValueFiller::ZERO;
        o.ignore_non_variants = f.ignore_non_variants;
        o.count = f.count;
        if (f.name == "AD") {
            if (f.type != RetainedFieldType::INT || f.number != RetainedFieldNumber::ALLELES)
                return Status::Invalid("genotyper misconfiguration: AD format field should have type=int, number=alleles");
            o.kind = GLX_FK_AD;
        } else if (f.name == "DP") {
            if (f.type != RetainedFieldType::INT || f.number != RetainedFieldNumber::BASIC || f.count != 1)
                return Status::Invalid("genotyper misconfiguration: DP format field should have type=int, number=basic, count=1");
            o.kind = GLX_FK_DP;
        } else if (f.name == "FT") {
            if (f.type != RetainedFieldType::STRING || f.number != RetainedFieldNumber::BASIC || f.count != 1)
                return Status::Invalid("genotyper misconfiguration: FT format field should have type=string, number=basic, count=1");
            o.kind = GLX_FK_FT;
        } else if (f.name == "PL") {
            if (f.type != RetainedFieldType::INT || f.number != RetainedFieldNumber::GENOTYPE || f.combi_method != FieldCombinationMethod::MISSING)
                return Status::Invalid("genotyper misconfiguration: PL format field should have type=int, number=genotype, combi_method=missing");
            if (cfg.more_PL && !cfg.squeeze) {
                // PLFieldHelper2's constructor asserts (genotyper_utils.h:545-553)
                if (!f.ignore_non_variants || f.from != RetainedFieldFrom::FORMAT || f.orig_names != std::vector<std::string>{"PL"})
                    return Status::Invalid("genotyper misconfiguration: more_PL requires PL from FORMAT/PL with ignore_non_variants");
                o.kind = GLX_FK_PL2;
            } else
                o.kind = GLX_FK_PL;
        } else if (f.type == RetainedFieldType::STRING) {
            o.kind = GLX_FK_STRING;
        } else {
            o.kind = GLX_FK_NUMERIC;
        }
        if (f.number == RetainedFieldNumber::BASIC && f.count < 0)
            return Status::Failure("setup_format_helpers: failed to identify count for format field");
        if (o.kind == GLX_FK_FT || o.kind == GLX_FK_STRING) continue;
        if (f.orig_names.size() > GLX_MAX_ORIG) return Status::NotImplemented("make_glx_config: too many orig_names", f.name);
        for (const auto& nm : f.orig_names) {
            int ci = get_col(c, nm, f.from == RetainedFieldFrom::INFO, f.type == RetainedFieldType::FLOAT ? GLX_TYPE_FLOAT : GLX_TYPE_INT);
            if (ci < 0) return Status::NotImplemented("make_glx_config: too many source columns");
            o.cols[o.n_cols++] = (int16_t)ci;
        }
    }
    return Status::OK();
}

extern "C" int glx_compute_priors(const glx_config* cfg, uint32_t n_sites, const uint32_t* allele_off, const float* allele_af,
                                  const float* lost_af, float* hom_log_prior, float* lost_log_prior) {
    if (!cfg || !allele_off || !allele_af || !lost_af || !hom_log_prior || !lost_log_prior) return GLX_E_INVALID;
    const float min_af = cfg->min_assumed_allele_frequency;
    for (uint32_t s = 0; s < n_sites; s++) {
        // const float lost_log_prior = log(std::max(us.lost_allele_frequency, cfg.min_assumed_allele_frequency));  genotyper.cc:154
        lost_log_prior[s] = logf(std::max(lost_af[s], min_af));
        for (uint32_t a = allele_off[s]; a < allele_off[s + 1]; a++) {
            // log(std::max(us.alleles[..].frequency, cfg.min_assumed_allele_frequency))  genotyper.cc:164 (float overload)
            hom_log_prior[a] = logf(std::max(allele_af[a], min_af));
        }
    }
    return GLX_OK;
}

static uint64_t prefix8(const std::string& s) {
    uint64_t p = 0;
    for (size_t i = 0; i < s.size() && i < 8; i++) p |= (uint64_t)(uint8_t)s[i] << (8 * i);
    return p;
}

static unsigned diploid_genotypes(unsigned n) { return (n + 1) * n / 2; }

Status pack_sites(const genotyper_config& cfg, const std::vector<unified_site>& sites, uint32_t n_samples, PackedBatch& b) {
    Status s;
    GLX_S(make_glx_config(cfg, b.cfg));
    b.gcfg = cfg;
    b.usites = sites;
    const size_t S = sites.size();
    b.s_beg.clear(); b.s_end.clear(); b.s_qbeg.clear(); b.s_qend.clear();
    b.s_n_allele.clear(); b.s_mono.clear(); b.s_is_indel.clear(); b.u_to.clear();
    b.s_allele_off.assign(1, 0); b.s_unif_off.assign(1, 0);
    b.u_len.clear(); b.u_str.clear(); b.s_af.clear(); b.s_lost_af.clear();
    b.u_beg.clear(); b.u_end.clear(); b.u_prefix.clear(); b.u_pool.clear();
    for (uint32_t k = 0; k < GLX_MAX_FIELDS; k++) {
        b.fld_cnt[k].clear();
        b.fld_off[k].assign(1, 0);
    }
    for (const auto& us : sites) {
        const size_t A = us.alleles.size();
        if (A < 2) return Status::Invalid("pack_sites: unified site has fewer than two alleles");
        if (A > GLX_MAX_ALLELES) return Status::NotImplemented("pack_sites: more unified alleles than GLX_MAX_ALLELES");
        b.s_beg.push_back(us.pos.beg);
        b.s_end.push_back(us.pos.end);
        int qb = us.pos.beg, qe = us.pos.end;  // genotyper.cc:700-706
        for (const auto& p : us.unification) {
            qb = std::min(qb, p.first.pos.beg);
            qe = std::max(qe, p.first.pos.end);
        }
        b.s_qbeg.push_back(qb);
        b.s_qend.push_back(qe);
        b.s_n_allele.push_back((uint8_t)A);
        b.s_mono.push_back(us.monoallelic ? 1 : 0);
        b.s_lost_af.push_back(us.lost_allele_frequency);
        for (size_t a = 0; a < A; a++) {
            b.s_is_indel.push_back(us.alleles[a].dna.size() != us.alleles[0].dna.size());
            b.s_af.push_back(us.alleles[a].frequency);
        }
        b.s_allele_off.push_back((uint32_t)b.s_af.size());
        for (const auto& p : us.unification) {
            b.u_beg.push_back(p.first.pos.beg);
            b.u_end.push_back(p.first.pos.end);
            b.u_to.push_back((uint8_t)p.second);
            b.u_len.push_back((uint32_t)p.first.dna.size());
            b.u_prefix.push_back(prefix8(p.first.dna));
            b.u_str.push_back((uint32_t)b.u_pool.size());
            b.u_pool.insert(b.u_pool.end(), p.first.dna.begin(), p.first.dna.end());
        }
        b.s_unif_off.push_back((uint32_t)b.u_to.size());
        for (uint32_t k = 0; k < b.cfg.n_fields; k++) {
            const glx_field& f = b.cfg.fields[k];
            unsigned cnt = 1;
            if (f.kind == GLX_FK_FT || f.kind == GLX_FK_STRING) cnt = 1;
            else if (f.number == GLX_NUM_BASIC) cnt = (unsigned)f.count;
            else if (f.number == GLX_NUM_ALT) cnt = (unsigned)A - 1;
            else if (f.number == GLX_NUM_ALLELES) cnt = (unsigned)A;
            else cnt = diploid_genotypes((unsigned)A);
            b.fld_cnt[k].push_back((uint16_t)cnt);
            b.fld_off[k].push_back(b.fld_off[k].back() + (uint64_t)cnt * n_samples);
        }
    }
    b.s_hom_prior.assign(b.s_af.size(), 0.0f);
    b.s_lost_prior.assign(S, 0.0f);
    if (S) glx_compute_priors(&b.cfg, (uint32_t)S, b.s_allele_off.data(), b.s_af.data(), b.s_lost_af.data(), b.s_hom_prior.data(), b.s_lost_prior.data());
    b.sites.n_sites = (uint32_t)S;
    b.sites.n_samples = n_samples;
    b.rebind();
    return Status::OK();
}

// src/genotyper.cc:20-32
static bool is_deletion(const std::string& ref, const std::string& alt) {
    if (alt.size() >= ref.size()) return false;
    for (size_t j = 0; j < alt.size(); j++)
        if (alt[j] != ref[j]) return false;
    return true;
}

// threads of pack_records: 0 = by the size of the input (tests force 1 and several to compare the two)
static std::atomic<int> g_pack_threads{0};
void set_pack_threads(int n) { g_pack_threads = n < 0 ? 0 : n; }

Status pack_records(const std::vector<GvcfDataset>& datasets, const VcfHeaderInfo& hdr, const std::vector<std::string>& samples,
                    PackedBatch& b) {
    const glx_config& c = b.cfg;
    b.sample_names = samples;
    b.contigs = hdr.contigs;
    std::map<std::string, int> samples_index;
    for (size_t i = 0; i < samples.size(); i++) samples_index[samples[i]] = (int)i;

    // datasets contributing to the sample set, ordered by output sample index
    std::vector<std::pair<int, const GvcfDataset*>> order;
    for (const auto& ds : datasets) {
        int first = -1;
        for (const auto& sm : ds.samples) {
            auto it = samples_index.find(sm);
            if (it != samples_index.end() && (first < 0 || it->second < first)) first = it->second;
        }
        if (first >= 0) order.emplace_back(first, &ds);  // sample_mapping.empty() => skipped, genotyper.cc:760-762
    }
    std::stable_sort(order.begin(), order.end(), [](const auto& x, const auto& y) { return x.first < y.first; });

    // FILTER dictionary (sorted => bit order == std::set<string> order of StringFormatFieldHelper)
    std::set<std::string> filt;
    for (const auto& o : order)
        for (const auto& r : o.second->records)
            for (const auto& f : r.filters)
                if (!f.empty() && f != "PASS") filt.insert(f);
    bool want_ft = false;
    for (uint32_t k = 0; k < c.n_fields; k++) want_ft |= c.fields[k].kind == GLX_FK_FT;
    if (filt.size() > GLX_MAX_FILTERS) {
        if (want_ft) return Status::NotImplemented("pack_records: more distinct FILTER ids than GLX_MAX_FILTERS");
        filt.clear();
    }
    b.filter_names.assign(filt.begin(), filt.end());

    // Datasets are independent: each is packed by one of T threads into the per-record arrays (at its record base, known up front) and
    // into pools of its own (values, GT pairs, allele descriptors, allele strings); the pools are then laid end to end in dataset order
    // and the offsets that point into them shifted by the dataset's base — the same bytes a serial pass writes.
    const size_t nD = order.size();
    std::vector<size_t> rec_base(nD + 1, 0);
    for (size_t d = 0; d < nD; d++) rec_base[d + 1] = rec_base[d] + order[d].second->records.size();
    const size_t R = rec_base[nD];
    b.ds_rec_off.assign(rec_base.begin(), rec_base.end());
    b.ds_n_sample.clear();
    b.ds_sample_off.assign(1, 0);
    b.sample_out.clear();
    for (const auto& o : order) {
        b.ds_n_sample.push_back((uint32_t)o.second->samples.size());
        for (const auto& sm : o.second->samples) {
            auto it = samples_index.find(sm);
            b.sample_out.push_back(it == samples_index.end() ? -1 : it->second);
        }
        b.ds_sample_off.push_back((uint32_t)b.sample_out.size());
    }
    b.r_beg.assign(R, 0); b.r_end.assign(R, 0); b.r_maxend.assign(R, 0); b.r_qual.assign(R, 0.0f); b.r_n_allele.assign(R, 0); b.r_flags.assign(R, 0);
    b.r_filt.assign(R, 0); b.r_gt.assign(R, 0); b.r_allele_off.assign(R, 0);
    b.col_val.assign((size_t)c.n_cols * R, 0);
    b.col_len.assign((size_t)c.n_cols * R, GLX_LEN_ABSENT);
    struct Local {
        std::vector<int32_t> pool, gt_pool;
        std::vector<uint32_t> al_len, al_str;
        std::vector<uint8_t> al_flags;
        std::vector<uint64_t> al_prefix;
        std::vector<char> al_pool;
        Status st;
    };
    std::vector<Local> loc(nD);
    auto pack_one = [&](size_t d) -> Status {
        const GvcfDataset& ds = *order[d].second;
        Local& L = loc[d];
        const uint32_t ns = (uint32_t)ds.samples.size();
        int32_t maxend = INT32_MIN;
        int prev_rid = -1, prev_beg = -1;
        size_t r_idx = rec_base[d];
        for (const auto& r : ds.records) {
            if (r.n_sample != (int)ns) return Status::Invalid("gVCF record doesn't have expected # of samples", ds.name);
            if (r.alleles.size() > GLX_MAX_ALLELES) return Status::NotImplemented("pack_records: more record alleles than GLX_MAX_ALLELES", ds.name);
            if (r.rid != prev_rid) maxend = INT32_MIN;
            else if (r.pos < prev_beg) return Status::Invalid("gVCF records are out-of-order ", ds.name);
            prev_rid = r.rid;
            prev_beg = r.pos;
            maxend = std::max(maxend, r.pos + r.rlen);
            b.r_beg[r_idx] = r.pos;
            b.r_end[r_idx] = r.pos + r.rlen;
            b.r_maxend[r_idx] = maxend;
            b.r_qual[r_idx] = r.qual;
            const size_t na = r.alleles.size();
            b.r_n_allele[r_idx] = (uint8_t)na;
            // is_gvcf_ref_record, src/types.cc:978-980
            const bool is_ref = na == 1 || (na == 2 && is_symbolic_allele(r.alleles[1]));
            uint8_t fl = is_ref ? GLX_RF_IS_REF : 0;
            if (is_symbolic_allele(r.alleles[na - 1])) fl |= GLX_RF_SYMBOLIC_LAST;
            uint32_t fm = 0;
            for (const auto& f : r.filters) {
                auto it = std::lower_bound(b.filter_names.begin(), b.filter_names.end(), f);
                if (it != b.filter_names.end() && *it == f) fm |= 1u << (it - b.filter_names.begin());
            }
            b.r_filt[r_idx] = fm;
            // GT (preprocess_record's view of bcf_get_genotypes, src/genotyper.cc:44-57)
            uint32_t gtv = 0;
            if (ns == 1) {
                if (r.n_gt == 1) {
                    fl |= GLX_RF_HAPLOID;
                    gtv = (uint16_t)(int16_t)(r.gt[0] & ~1) | ((uint32_t)(uint16_t)GLX_GT16_VECTOR_END << 16);
                } else if (r.n_gt == 2) {
                    auto enc = [](int32_t v) -> uint16_t { return v == GLX_INT_VECTOR_END ? (uint16_t)GLX_GT16_VECTOR_END : (uint16_t)(int16_t)(v & ~1); };
                    gtv = enc(r.gt[0]) | ((uint32_t)enc(r.gt[1]) << 16);
                } else
                    fl |= GLX_RF_NO_GT;
            } else {
                if (r.n_gt == 2) {
                    gtv = (uint32_t)L.gt_pool.size();  // (dataset-local: shifted below)
                    for (auto v : r.gt) L.gt_pool.push_back(v == GLX_INT_VECTOR_END ? v : (v & ~1));
                } else
                    fl |= GLX_RF_NO_GT;
            }
            b.r_gt[r_idx] = gtv;
            b.r_flags[r_idx] = fl;
            // allele descriptors (preprocess_record's is_dna / deletion_allele inputs, src/genotyper.cc:59-77)
            b.r_allele_off[r_idx] = (uint32_t)L.al_len.size();  // (dataset-local)
            if (!is_ref) {
                const std::string& ref = r.alleles[0];
                for (size_t i = 0; i < na; i++) {
                    const std::string& al = r.alleles[i];
                    uint8_t af = 0;
                    if (is_dna(al)) af |= GLX_AF_IS_DNA;
                    if (is_symbolic_allele(al)) af |= GLX_AF_SYMBOLIC;
                    if (i > 0 && al.size() < (size_t)r.rlen && (size_t)r.rlen == ref.size() && is_deletion(ref, al)) af |= GLX_AF_DELETION;
                    L.al_len.push_back((uint32_t)al.size());
                    L.al_flags.push_back(af);
                    L.al_prefix.push_back(prefix8(al));
                    L.al_str.push_back((uint32_t)L.al_pool.size());  // (dataset-local)
                    L.al_pool.insert(L.al_pool.end(), al.begin(), al.end());
                }
            }
            // source columns
            for (uint32_t ci = 0; ci < c.n_cols; ci++) {
                const glx_column& col = c.cols[ci];
                const auto& types = col.from_info ? hdr.info_types : hdr.format_types;
                auto ht = types.find(col.name);
                if (ht == types.end()) continue;  // htslib rv -1
                if (ht->second != col.type) {     // htslib rv -2
                    b.col_len[(size_t)ci * R + r_idx] = 0xFFFE;
                    continue;
                }
                const FieldValues* fv = nullptr;
                for (const auto& f : (col.from_info ? r.info : r.format))
                    if (f.id == col.name) fv = &f;
                if (!fv || fv->v.empty()) continue;  // htslib rv -3
                const uint32_t len = col.from_info ? (uint32_t)fv->v.size() / std::max(1u, ns) : (uint32_t)fv->per_sample;
                if (col.from_info && (uint32_t)fv->v.size() != len * ns) {
                    // INFO vector not divisible among samples: rv != n_sample*n_val => length error downstream
                    b.col_len[(size_t)ci * R + r_idx] = 0xFFFD;
                    continue;
                }
                if (len >= 0xFFF0) return Status::NotImplemented("pack_records: field vector too long", col.name);
                b.col_len[(size_t)ci * R + r_idx] = (uint16_t)len;
                if (ns == 1 && len == 1) b.col_val[(size_t)ci * R + r_idx] = fv->v[0];
                else {
                    b.col_val[(size_t)ci * R + r_idx] = (int32_t)L.pool.size();  // (dataset-local)
                    L.pool.insert(L.pool.end(), fv->v.begin(), fv->v.end());
                }
            }
            r_idx++;
        }
        return Status::OK();
    };
    unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>({(size_t)16, (size_t)std::thread::hardware_concurrency(), nD, R / 4096 + 1}));
    if (g_pack_threads.load() > 0) T = (unsigned)std::min<size_t>((size_t)g_pack_threads.load(), std::max<size_t>(nD, 1));
    auto over_datasets = [&](auto&& fn) {  // fn(d) for every dataset, dynamically shared among T threads
        if (T <= 1) {
            for (size_t d = 0; d < nD; d++) fn(d);
            return;
        }
        std::atomic<size_t> next{0};
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < T; t++)
            pool.emplace_back([&]() {
                for (size_t d = next++; d < nD; d = next++) fn(d);
            });
        for (auto& th : pool) th.join();
    };
    over_datasets([&](size_t d) { loc[d].st = pack_one(d); });
    for (size_t d = 0; d < nD; d++)
        if (loc[d].st.bad()) return loc[d].st;  // the first failing dataset, as a serial pass would report
    std::vector<size_t> pool_base(nD + 1, 0), gt_base(nD + 1, 0), al_base(nD + 1, 0), str_base(nD + 1, 0);
    for (size_t d = 0; d < nD; d++) {
        pool_base[d + 1] = pool_base[d] + loc[d].pool.size();
        gt_base[d + 1] = gt_base[d] + loc[d].gt_pool.size();
        al_base[d + 1] = al_base[d] + loc[d].al_len.size();
        str_base[d + 1] = str_base[d] + loc[d].al_pool.size();
    }
    if (pool_base[nD] > (size_t)INT32_MAX || gt_base[nD] > UINT32_MAX || al_base[nD] > UINT32_MAX || str_base[nD] > UINT32_MAX)
        return Status::NotImplemented("pack_records: the range's pools exceed 32-bit offsets; use smaller range batches");
    b.pool.resize(pool_base[nD]);
    b.gt_pool.resize(gt_base[nD]);
    b.al_len.resize(al_base[nD]); b.al_str.resize(al_base[nD]); b.al_flags.resize(al_base[nD]); b.al_prefix.resize(al_base[nD]);
    b.al_pool.resize(str_base[nD]);
    over_datasets([&](size_t d) {
        Local& L = loc[d];
        const uint32_t ns = b.ds_n_sample[d];
        std::copy(L.pool.begin(), L.pool.end(), b.pool.begin() + pool_base[d]);
        std::copy(L.gt_pool.begin(), L.gt_pool.end(), b.gt_pool.begin() + gt_base[d]);
        std::copy(L.al_len.begin(), L.al_len.end(), b.al_len.begin() + al_base[d]);
        std::copy(L.al_flags.begin(), L.al_flags.end(), b.al_flags.begin() + al_base[d]);
        std::copy(L.al_prefix.begin(), L.al_prefix.end(), b.al_prefix.begin() + al_base[d]);
        std::copy(L.al_pool.begin(), L.al_pool.end(), b.al_pool.begin() + str_base[d]);
        for (size_t a = 0; a < L.al_str.size(); a++) b.al_str[al_base[d] + a] = L.al_str[a] + (uint32_t)str_base[d];
        for (size_t r = rec_base[d]; r < rec_base[d + 1]; r++) {
            b.r_allele_off[r] += (uint32_t)al_base[d];
            if (ns != 1 && !(b.r_flags[r] & GLX_RF_NO_GT)) b.r_gt[r] += (uint32_t)gt_base[d];
            for (uint32_t ci = 0; ci < c.n_cols; ci++) {
                const uint16_t len = b.col_len[(size_t)ci * R + r];
                if (len >= 0xFFF0 || (ns == 1 && len == 1)) continue;
                b.col_val[(size_t)ci * R + r] += (int32_t)pool_base[d];
            }
        }
        L = Local();
    });
    b.recs.n_datasets = (uint32_t)order.size();
    b.recs.n_samples_out = (uint32_t)samples.size();
    b.recs.n_cols = c.n_cols;
    b.recs.n_records = R;
    b.rebind();
    return Status::OK();
}

template <class T>
static const T* ptr_or_dummy(const std::vector<T>& v) {
    static const T dummy[2] = {};
    return v.empty() ? dummy : v.data();
}

void PackedBatch::rebind() {
    recs.ds_rec_off = ptr_or_dummy(ds_rec_off);
    recs.ds_n_sample = ptr_or_dummy(ds_n_sample);
    recs.ds_sample_off = ptr_or_dummy(ds_sample_off);
    recs.sample_out = ptr_or_dummy(sample_out);
    recs.beg = ptr_or_dummy(r_beg);
    recs.end = ptr_or_dummy(r_end);
    recs.maxend = ptr_or_dummy(r_maxend);
    recs.qual = ptr_or_dummy(r_qual);
    recs.n_allele = ptr_or_dummy(r_n_allele);
    recs.flags = ptr_or_dummy(r_flags);
    recs.filt = ptr_or_dummy(r_filt);
    recs.gt = ptr_or_dummy(r_gt);
    recs.allele_off = ptr_or_dummy(r_allele_off);
    recs.col_val = ptr_or_dummy(col_val);
    recs.col_len = ptr_or_dummy(col_len);
    recs.pool = ptr_or_dummy(pool);
    recs.gt_pool = ptr_or_dummy(gt_pool);
    recs.al_len = ptr_or_dummy(al_len);
    recs.al_flags = ptr_or_dummy(al_flags);
    recs.al_prefix = ptr_or_dummy(al_prefix);
    recs.al_str = ptr_or_dummy(al_str);
    recs.al_pool = ptr_or_dummy(al_pool);
    recs.n_pool = pool.size();
    recs.n_gt_pool = gt_pool.size();
    recs.n_al = al_len.size();
    recs.n_al_pool = al_pool.size();
    filter_name_ptrs.clear();
    for (const auto& f : filter_names) filter_name_ptrs.push_back(f.c_str());
    recs.n_filters = (uint32_t)filter_names.size();
    recs.filter_names = filter_name_ptrs.empty() ? nullptr : filter_name_ptrs.data();

    sites.beg = ptr_or_dummy(s_beg);
    sites.end = ptr_or_dummy(s_end);
    sites.qbeg = ptr_or_dummy(s_qbeg);
    sites.qend = ptr_or_dummy(s_qend);
    sites.n_allele = ptr_or_dummy(s_n_allele);
    sites.monoallelic = ptr_or_dummy(s_mono);
    sites.allele_off = ptr_or_dummy(s_allele_off);
    sites.allele_is_indel = ptr_or_dummy(s_is_indel);
    sites.allele_af = ptr_or_dummy(s_af);
    sites.lost_af = ptr_or_dummy(s_lost_af);
    sites.hom_log_prior = ptr_or_dummy(s_hom_prior);
    sites.lost_log_prior = ptr_or_dummy(s_lost_prior);
    sites.unif_off = ptr_or_dummy(s_unif_off);
    sites.unif_beg = ptr_or_dummy(u_beg);
    sites.unif_end = ptr_or_dummy(u_end);
    sites.unif_to = ptr_or_dummy(u_to);
    sites.unif_len = ptr_or_dummy(u_len);
    sites.unif_prefix = ptr_or_dummy(u_prefix);
    sites.unif_str = ptr_or_dummy(u_str);
    sites.unif_pool = ptr_or_dummy(u_pool);
    sites.n_unif = u_to.size();
    sites.n_unif_pool = u_pool.size();
    sites.n_site_alleles = s_af.size();
    for (uint32_t k = 0; k < GLX_MAX_FIELDS; k++) {
        sites.fld_cnt[k] = ptr_or_dummy(fld_cnt[k]);
        sites.fld_off[k] = ptr_or_dummy(fld_off[k]);
    }
}

void OutBuffers::allocate(const PackedBatch& b) {
    const size_t S = b.sites.n_sites, N = b.sites.n_samples;
    gt.assign(S * N * 2, -1);
    rnc.assign(S * N * 2, 'M');
    site_flags.assign(S, 0);
    allele_used.assign(S, 0);
    memset(&out, 0, sizeof out);
    for (uint32_t k = 0; k < b.cfg.n_fields; k++) {
        fld[k].assign(b.fld_off[k].empty() ? 0 : b.fld_off[k].back(), 0);
        out.fld[k] = fld[k].empty() ? nullptr : fld[k].data();
    }
    out.gt = gt.data();
    out.rnc = rnc.data();
    out.site_flags = site_flags.data();
    out.allele_used = allele_used.data();
}

// ---------------------------------------------------------------------------------------------
// Output assembly
// ---------------------------------------------------------------------------------------------
static void put_float(std::string& o, float f) {
    char buf[48];
    snprintf(buf, sizeof buf, "%g", (double)f);
    o += buf;
}

static char header_number_class(const std::string& description) {
    std::string n = header_attr(description, "Number");
    return n.size() == 1 ? n[0] : '1';
}

// pVCF header, src/service.cc:190-241
void vcf_header_text(const PackedBatch& b, std::string& o) {
    const genotyper_config& cfg = b.gcfg;
    o += "##fileformat=VCFv4.2\n";
    o += "##FILTER=<ID=PASS,Description=\"All filters passed\">\n";
    o += "##GLnexusVersion=glnexus_b200\n";
    o += "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Allele Frequency estimate for each alternate allele\">\n";
    o += "##INFO=<ID=AQ,Number=A,Type=Integer,Description=\"Allele Quality score reflecting evidence for each alternate allele (Phred scale)\">\n";
    o += "##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Allele count in genotypes\">\n";
    o += "##INFO=<ID=AN,Number=1,Type=Integer,Description=\"Total number of alleles in called genotypes\">\n";
    o += "##FILTER=<ID=MONOALLELIC,Description=\"Site represents one ALT allele in a region with multiple variants that could not be unified into non-overlapping multi-allelic sites\">\n";
    o += "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
    o += "##FORMAT=<ID=RNC,Number=2,Type=Character,Description=\"Reason for No Call in GT: . = n/a, M = Missing data, P = Partial data, I = gVCF input site is non-called, D = insufficient Depth of coverage, - = unrepresentable overlapping deletion, L = Lost/unrepresentable allele (other than deletion), U = multiple Unphased variants present, O = multiple Overlapping variants present, 1 = site is Monoallelic, no assertion about presence of REF or ALT allele\">\n";
    for (const auto& f : cfg.liftover_fields) o += f.description + "\n";
    for (const auto& ctg : b.contigs) {
        o += "##contig=<ID=" + ctg.first;
        if (ctg.second != 1000000042 && ctg.second != 0) o += ",length=" + std::to_string(ctg.second);
        o += ">\n";
    }
    o += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
    for (const auto& sm : b.sample_names) o += "\t" + sm;
    o += "\n";
}

Status assemble_vcf_text(const PackedBatch& b, const glx_out& out, std::string& o) {
    o.clear();
    vcf_header_text(b, o);
    return vcf_rows_text(b, out, o);
}

// one pVCF row per surviving site of the batch, appended to `o`
Status vcf_rows_text(const PackedBatch& b, const glx_out& out, std::string& o) {
    const genotyper_config& cfg = b.gcfg;
    const size_t N = b.sites.n_samples;
    std::vector<char> number_class;
    for (const auto& f : cfg.liftover_fields) number_class.push_back(header_number_class(f.description));

    for (size_t si = 0; si < b.usites.size(); si++) {
        const unified_site& site = b.usites[si];
        const int A = (int)site.alleles.size();
        // bcf_trim_alleles (htslib vcfutils.c) restated: drop ALT alleles no GT entry uses, genotyper.cc:958-966
        std::vector<int> keep;  // surviving allele indices
        keep.push_back(0);
        for (int a = 1; a < A; a++)
            if (!cfg.trim_uncalled_alleles || (out.allele_used[si] >> a & 1)) keep.push_back(a);
        if ((int)keep.size() < 2) continue;  // ans.reset(): site dropped
        std::vector<int> renum(A, -1);
        for (size_t i = 0; i < keep.size(); i++) renum[keep[i]] = (int)i;
        const bool trimmed = (int)keep.size() != A;

        o += b.contigs.at(site.pos.rid).first;
        o += "\t" + std::to_string(site.pos.beg + 1) + "\t";
        // ID, genotyper.cc:968-993
        for (size_t i = 1; i < keep.size(); i++) {
            const auto& norm = site.alleles[keep[i]].normalized;
            if (i > 1) o += ";";
            o += b.contigs.at(site.pos.rid).first + "_" + std::to_string(norm.pos.beg + 1) + "_" +
                 site.alleles[0].dna.substr(norm.pos.beg - site.pos.beg, norm.pos.size()) + "_" + norm.dna;
        }
        o += "\t" + site.alleles[0].dna + "\t";
        for (size_t i = 1; i < keep.size(); i++) o += (i > 1 ? "," : "") + site.alleles[keep[i]].dna;
        o += "\t";
        put_float(o, (float)site.qual);
        o += site.monoallelic ? "\tMONOALLELIC\t" : "\t.\t";
        // INFO AF / AQ, genotyper.cc:884-912
        bool output_af = true, any_aq = false;
        for (int a = 1; a < A; a++) {
            if (site.alleles[a].frequency != site.alleles[a].frequency) output_af = false;
            if (site.alleles[a].quality) any_aq = true;
        }
        std::string info;
        if (output_af) {
            info += "AF=";
            for (size_t i = 1; i < keep.size(); i++) {
                if (i > 1) info += ",";
                put_float(info, site.alleles[keep[i]].frequency);
            }
        }
        if (any_aq) {
            if (!info.empty()) info += ";";
            info += "AQ=";
            for (size_t i = 1; i < keep.size(); i++) info += (i > 1 ? "," : "") + std::to_string(site.alleles[keep[i]].quality);
        }
        o += info.empty() ? "." : info;
        o += "\tGT";
        for (const auto& f : cfg.liftover_fields) o += ":" + f.name;
        o += ":RNC";
        for (size_t n = 0; n < N; n++) {
            o += "\t";
            for (int h = 0; h < 2; h++) {
                int g = out.gt[(si * N + n) * 2 + h];
                if (h) o += "/";
                if (g < 0) o += ".";
                else o += std::to_string(trimmed ? renum[g] : g);
            }
            for (uint32_t k = 0; k < b.cfg.n_fields; k++) {
                const glx_field& f = b.cfg.fields[k];
                const unsigned cnt = b.fld_cnt[k][si];
                const int32_t* v = out.fld[k] + b.fld_off[k][si] + (uint64_t)n * cnt;
                o += ":";
                if (f.kind == GLX_FK_FT || f.kind == GLX_FK_STRING) {
                    uint32_t m = (uint32_t)v[0];
                    if (!m) o += ".";
                    bool first = true;
                    for (uint32_t i = 0; i < b.filter_names.size(); i++)
                        if (m >> i & 1) {
                            if (!first) o += ";";
                            o += b.filter_names[i];
                            first = false;
                        }
                    continue;
                }
                // select the entries surviving bcf_remove_allele_set for Number=A/R/G fields
                std::vector<int32_t> vals;
                const char nc = number_class[k];
                const bool is_float = f.type == GLX_TYPE_FLOAT;
                const int32_t vend = is_float ? (int32_t)GLX_FLOAT_VECTOR_END_BITS : GLX_INT_VECTOR_END;
                if (!trimmed || (nc != 'A' && nc != 'R' && nc != 'G')) vals.assign(v, v + cnt);
                else if (nc == 'G') {
                    for (unsigned j = 0; j < cnt; j++) {
                        if (v[j] == vend) { vals.push_back(vend); break; }
                        unsigned jb = (unsigned)((sqrt(8.0 * j + 1) - 1.0) / 2.0), ja = j - jb * (jb + 1) / 2;
                        if ((int)jb < A && renum[ja] >= 0 && renum[jb] >= 0) vals.push_back(v[j]);
                    }
                } else {
                    const int shift = nc == 'A' ? 1 : 0;
                    for (unsigned j = 0; j < cnt; j++) {
                        if (v[j] == vend) { vals.push_back(vend); break; }
                        if ((int)j + shift < A && renum[j + shift] >= 0) vals.push_back(v[j]);
                    }
                }
                bool first = true;
                for (int32_t x : vals) {
                    if (x == vend) break;
                    if (!first) o += ",";
                    first = false;
                    if (is_float) {
                        if ((uint32_t)x == GLX_FLOAT_MISSING_BITS) o += ".";
                        else put_float(o, bits_float(x));
                    } else {
                        if (x == GLX_INT_MISSING) o += ".";
                        else o += std::to_string(x);
                    }
                }
                if (first) o += ".";
            }
            o += ":";
            o += (char)out.rnc[(si * N + n) * 2];
            o += (char)out.rnc[(si * N + n) * 2 + 1];
        }
        o += "\n";
    }
    return Status::OK();
}

// ---------------------------------------------------------------------------------------------
// Range sharding: sites are independent, so a batch is cut into contiguous site ranges, one per
// GPU; each shard carries every record overlapping its sites' query hull (records straddling a cut are replicated,
// like the reference's bucket "danglers", src/BCFKeyValueData.cc:815-827). No collective; outputs concatenate in order.
// ---------------------------------------------------------------------------------------------
void partition_sites(const PackedBatch& b, int n_parts, std::vector<uint32_t>& bounds) {
    const uint32_t S = b.sites.n_sites;
    bounds.assign(n_parts + 1, S);
    bounds[0] = 0;
    std::vector<double> w(S + 1, 0.0);  // prefix of output bytes per site: N * (4 + 4 * values per sample)
    for (uint32_t s = 0; s < S; s++) {
        double vals = 1.0;
        for (uint32_t k = 0; k < b.cfg.n_fields; k++) vals += b.fld_cnt[k][s];
        w[s + 1] = w[s] + vals;
    }
    for (int p = 1; p < n_parts; p++) {
        const double target = w[S] * p / n_parts;
        bounds[p] = (uint32_t)(std::lower_bound(w.begin(), w.end(), target) - w.begin());
        if (bounds[p] < bounds[p - 1]) bounds[p] = bounds[p - 1];
        if (bounds[p] > S) bounds[p] = S;
    }
}

// the same over a cohort held as consecutive segments: cut p = (segment, site within it); cuts[n_parts] = (n_segments, 0)
void partition_cohort(const std::vector<const PackedBatch*>& segments, int n_parts, std::vector<std::pair<uint32_t, uint32_t>>& cuts) {
    std::vector<double> seg_w(segments.size() + 1, 0.0);
    auto site_w = [](const PackedBatch& b, uint32_t s) {
        double vals = 1.0;
        for (uint32_t k = 0; k < b.cfg.n_fields; k++) vals += b.fld_cnt[k][s];
        return vals;
    };
    for (size_t g = 0; g < segments.size(); g++) {
        double w = 0;
        for (uint32_t s = 0; s < segments[g]->sites.n_sites; s++) w += site_w(*segments[g], s);
        seg_w[g + 1] = seg_w[g] + w;
    }
    cuts.assign(n_parts + 1, {(uint32_t)segments.size(), 0u});
    cuts[0] = {0u, 0u};
    for (int p = 1; p < n_parts; p++) {
        const double target = seg_w.back() * p / n_parts;
        size_t g = 0;
        while (g + 1 < segments.size() && seg_w[g + 1] <= target) g++;
        double acc = seg_w[g];
        uint32_t s = 0;
        while (s < segments[g]->sites.n_sites && acc < target) acc += site_w(*segments[g], s++);
        cuts[p] = {(uint32_t)g, s};
        if (cuts[p] < cuts[p - 1]) cuts[p] = cuts[p - 1];
    }
}

Status slice_batch(const PackedBatch& b, uint32_t site_lo, uint32_t site_hi, PackedBatch& o) {
    if (site_lo > site_hi || site_hi > b.sites.n_sites) return Status::Invalid("slice_batch: bad site range");
    Status s;
    std::vector<unified_site> sub(b.usites.begin() + site_lo, b.usites.begin() + site_hi);
    GLX_S(pack_sites(b.gcfg, sub, b.sites.n_samples, o));
    o.sample_names = b.sample_names;
    o.contigs = b.contigs;
    o.filter_names = b.filter_names;
    int qmin = INT32_MAX, qmax = INT32_MIN;
    for (uint32_t i = site_lo; i < site_hi; i++) {
        qmin = std::min(qmin, b.s_qbeg[i]);
        qmax = std::max(qmax, b.s_qend[i]);
    }
    const uint32_t D = b.recs.n_datasets, C = b.cfg.n_cols;
    const uint64_t R = b.recs.n_records;
    o.ds_n_sample = b.ds_n_sample;
    o.ds_sample_off = b.ds_sample_off;
    o.sample_out = b.sample_out;
    o.ds_rec_off.assign(1, 0);
    std::vector<std::pair<uint64_t, uint64_t>> spans(D);
    uint64_t Rn = 0;
    for (uint32_t d = 0; d < D; d++) {
        const uint64_t lo = b.ds_rec_off[d], hi = b.ds_rec_off[d + 1];
        uint64_t first = hi, last = hi;
        if (site_lo < site_hi) {
            first = std::upper_bound(b.r_maxend.begin() + lo, b.r_maxend.begin() + hi, qmin) - b.r_maxend.begin();
            last = first;
            while (last < hi && b.r_beg[last] < qmax) last++;
        }
        spans[d] = {first, last};
        Rn += last - first;
        o.ds_rec_off.push_back(Rn);
    }
    o.col_val.assign((size_t)C * Rn, 0);
    o.col_len.assign((size_t)C * Rn, GLX_LEN_ABSENT);
    uint64_t rn = 0;
    for (uint32_t d = 0; d < D; d++) {
        const uint32_t ns = b.ds_n_sample[d];
        int32_t maxend = INT32_MIN;
        for (uint64_t r = spans[d].first; r < spans[d].second; r++, rn++) {
            maxend = std::max(maxend, b.r_end[r]);
            o.r_beg.push_back(b.r_beg[r]);
            o.r_end.push_back(b.r_end[r]);
            o.r_maxend.push_back(maxend);
            o.r_qual.push_back(b.r_qual[r]);
            o.r_n_allele.push_back(b.r_n_allele[r]);
            o.r_flags.push_back(b.r_flags[r]);
            o.r_filt.push_back(b.r_filt[r]);
            if (ns == 1 || (b.r_flags[r] & GLX_RF_NO_GT)) o.r_gt.push_back(b.r_gt[r]);
            else {
                o.r_gt.push_back((uint32_t)o.gt_pool.size());
                o.gt_pool.insert(o.gt_pool.end(), b.gt_pool.begin() + b.r_gt[r], b.gt_pool.begin() + b.r_gt[r] + 2 * ns);
            }
            o.r_allele_off.push_back((uint32_t)o.al_len.size());
            if (!(b.r_flags[r] & GLX_RF_IS_REF)) {
                for (uint32_t a = b.r_allele_off[r]; a < b.r_allele_off[r] + b.r_n_allele[r]; a++) {
                    o.al_len.push_back(b.al_len[a]);
                    o.al_flags.push_back(b.al_flags[a]);
                    o.al_prefix.push_back(b.al_prefix[a]);
                    o.al_str.push_back((uint32_t)o.al_pool.size());
                    o.al_pool.insert(o.al_pool.end(), b.al_pool.begin() + b.al_str[a], b.al_pool.begin() + b.al_str[a] + b.al_len[a]);
                }
            }
            for (uint32_t c = 0; c < C; c++) {
                const uint16_t len = b.col_len[(size_t)c * R + r];
                const int32_t val = b.col_val[(size_t)c * R + r];
                o.col_len[(size_t)c * Rn + rn] = len;
                if (len >= 0xFFF0) continue;
                if (ns == 1 && len == 1) o.col_val[(size_t)c * Rn + rn] = val;
                else {
                    o.col_val[(size_t)c * Rn + rn] = (int32_t)o.pool.size();
                    o.pool.insert(o.pool.end(), b.pool.begin() + val, b.pool.begin() + val + (size_t)ns * len);
                }
            }
        }
    }
    o.recs.n_datasets = D;
    o.recs.n_samples_out = b.recs.n_samples_out;
    o.recs.n_cols = C;
    o.recs.n_records = Rn;
    o.rebind();
    return Status::OK();
}

}  // namespace glx
