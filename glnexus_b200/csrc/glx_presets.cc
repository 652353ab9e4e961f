// glx_presets.cc — genotyper half of the `glnexus_cli --config NAME` presets.
// Same names and values as the reference's preset table (src/cli_utils.cc:465-1188, load_config :1236-1265),
// expressed with a few field builders instead of YAML. The unifier half is out of scope.
#include "glx_host.hpp"

namespace glx {
namespace {

using F = retained_format_field;
using N = RetainedFieldNumber;
using C = FieldCombinationMethod;
using T = RetainedFieldType;

F field(const char* name, std::vector<std::string> orig, T type, N number, int count, C combi, bool ignore_non_variants, const char* desc,
        RetainedFieldFrom from = RetainedFieldFrom::FORMAT, DefaultValueFiller dflt = DefaultValueFiller::MISSING) {
    F f;
    f.name = name;
    f.orig_names = std::move(orig);
    f.type = type;
    f.number = number;
    f.count = count;
    f.combi_method = combi;
    f.ignore_non_variants = ignore_non_variants;
    f.description = desc;
    f.from = from;
    f.default_type = dflt;
    return f;
}

const char* DP_GATK = "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Approximate read depth (reads with MQ=255 or with bad mates are filtered)\">";
const char* DP_PLAIN = "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Read Depth\">";
const char* AD_DESC = "##FORMAT=<ID=AD,Number=R,Type=Integer,Description=\"Allelic depths for the ref and alt alleles in the order listed\">";
const char* GQ_DESC = "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype Quality\">";
const char* PL_DESC = "##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"Phred-scaled genotype Likelihoods\">";
const char* SB_DESC = "##FORMAT=<ID=SB,Number=4,Type=Integer,Description=\"Per-sample component statistics which comprise the Fishers Exact Test to detect strand bias.\">";
const char* SBPV_DESC = "##FORMAT=<ID=SBPV,Number=A,Type=Float,Description=\"Strand bias P-value; probability that the fraction of forward reads (VCF) amongst reads supporting alt allele (VC) is more extreme than expected assuming a beta-binomial distribution.\">";

F dp(std::vector<std::string> orig, const char* desc, bool inv = true) { return field("DP", std::move(orig), T::INT, N::BASIC, 1, C::MIN, inv, desc); }
F ad(bool inv = false) { return field("AD", {"AD"}, T::INT, N::ALLELES, 0, C::MIN, inv, AD_DESC, RetainedFieldFrom::FORMAT, DefaultValueFiller::ZERO); }
F gq() { return field("GQ", {"GQ"}, T::INT, N::BASIC, 1, C::MIN, true, GQ_DESC); }
F pl() { return field("PL", {"PL"}, T::INT, N::GENOTYPE, 0, C::MISSING, true, PL_DESC); }
F ft(const char* desc) { return field("FT", {"FILTER"}, T::STRING, N::BASIC, 1, C::SEMICOLON, true, desc); }

genotyper_config gatk(bool revise) {
    genotyper_config c;
    c.required_dp = 1;
    c.revise_genotypes = revise;
    c.liftover_fields = {dp({"MIN_DP", "DP"}, DP_GATK), ad(), field("SB", {"SB"}, T::INT, N::BASIC, 4, C::MISSING, false, SB_DESC), gq(), pl()};
    return c;
}

genotyper_config xatlas(bool revise) {
    genotyper_config c;
    c.required_dp = 1;
    c.allow_partial_data = true;
    c.revise_genotypes = revise;
    c.ref_dp_format = "RR";
    c.allele_dp_format = "VR";
    c.liftover_fields = {
        dp({"DP"}, DP_PLAIN, false),
        field("RR", {"RR"}, T::INT, N::BASIC, 1, C::MIN, false, "##FORMAT=<ID=RR,Number=1,Type=Integer,Description=\"Reference Read Depth\">"),
        field("VR", {"VR"}, T::INT, N::BASIC, 1, C::MAX, true, "##FORMAT=<ID=VR,Number=1,Type=Integer,Description=\"Major Variant Read Depth\">"),
        gq(),
        pl(),
        field("P", {"P"}, T::FLOAT, N::BASIC, 1, C::MISSING, true, "##FORMAT=<ID=P,Number=1,Type=Float,Description=\"xAtlas variant p-value\">", RetainedFieldFrom::INFO),
        ft("##FORMAT=<ID=FT,Number=1,Type=String,Description=\"FILTER field from sample gVCF (other than PASS)\">")};
    return c;
}

genotyper_config wecall(bool revise) {
    genotyper_config c;
    c.required_dp = 1;
    c.revise_genotypes = revise;
    c.liftover_fields = {dp({"MIN_DP", "DP"}, DP_PLAIN),
                         ad(),
                         field("SBPV", {"SBPV"}, T::FLOAT, N::ALT, 0, C::MISSING, true, SBPV_DESC, RetainedFieldFrom::INFO),
                         gq(),
                         pl(),
                         ft("##FORMAT=<ID=FT,Number=1,Type=String,Description=\"FILTER field from sample gVCF\">")};
    return c;
}

genotyper_config deepvariant(bool revise, float snv, float indel, const char* ref_dp) {
    genotyper_config c;
    c.required_dp = 0;
    c.revise_genotypes = revise;
    c.snv_prior_calibration = snv;
    c.indel_prior_calibration = indel;
    c.allow_partial_data = true;
    c.more_PL = true;
    c.trim_uncalled_alleles = true;
    c.ref_dp_format = ref_dp;
    c.liftover_fields = {dp({ref_dp, "DP"}, DP_GATK), ad(), gq(), pl()};
    return c;
}

genotyper_config strelka2() {
    genotyper_config c;
    c.required_dp = 0;
    c.revise_genotypes = false;
    c.liftover_fields = {dp({"MIN_DP", "DP", "DPI"}, DP_GATK), ad(true), field("GQ", {"GQ"}, T::FLOAT, N::BASIC, 1, C::MIN, true, GQ_DESC),
                         ft("##FORMAT=<ID=FT,Number=1,Type=String,Description=\"FILTER field from sample gVCF\">")};
    return c;
}

genotyper_config gxs() {
    genotyper_config c;
    c.required_dp = 0;
    c.revise_genotypes = false;
    c.allow_partial_data = true;
    c.more_PL = true;
    c.trim_uncalled_alleles = false;
    auto info_alt = [](const char* name, const char* orig, T type, const char* desc) {
        return field(name, {orig}, type, N::ALT, 0, C::MISSING, true, desc, RetainedFieldFrom::INFO);
    };
    c.liftover_fields = {
        field("DS", {"DS"}, T::FLOAT, N::BASIC, 1, C::MAX, true, "##FORMAT=<ID=DS,Number=1,Type=Float,Description=\"Genotype dosage\">"),
        field("GP", {"GP"}, T::FLOAT, N::BASIC, 3, C::MISSING, true, "##FORMAT=<ID=GP,Number=3,Type=Float,Description=\"Genotype posteriors\">"),
        field("HS", {"HS"}, T::INT, N::BASIC, 1, C::MISSING, true,
              "##FORMAT=<ID=HS,Number=1,Type=Integer,Description=\"Sampled haplotype pairs packed into integers (max: 16 pairs, see NMAIN header line)\">"),
        info_alt("RAF", "RAF", T::FLOAT, "##FORMAT=<ID=RAF,Number=A,Type=Float,Description=\"ALT allele frequency in the reference panel\">"),
        info_alt("TAF", "AF", T::FLOAT, "##FORMAT=<ID=TAF,Number=A,Type=Float,Description=\"ALT allele frequency computed from DS/GP field across target samples\">"),
        info_alt("IIS", "INFO", T::FLOAT, "##FORMAT=<ID=IIS,Number=A,Type=Float,Description=\"Imputation information or quality score\">"),
        info_alt("BUF", "BUF", T::INT, "##FORMAT=<ID=BUF,Number=A,Type=Integer,Description=\"Is it a variant site falling within buffer regions? (0=no/1=yes)\">")};
    return c;
}

}  // namespace

std::vector<std::string> config_preset_names() {
    return {"gatk", "gatk_unfiltered", "xAtlas", "xAtlas_unfiltered", "weCall", "weCall_unfiltered", "DeepVariant", "DeepVariantWGS",
            "DeepVariantWES", "DeepVariantWES_MED_DP", "DeepVariant_unfiltered", "Strelka2", "GxS"};
}

Status load_config_preset(const std::string& name, genotyper_config& ans) {
    if (name == "gatk") ans = gatk(true);
    else if (name == "gatk_unfiltered") ans = gatk(false);
    else if (name == "xAtlas") ans = xatlas(true);
    else if (name == "xAtlas_unfiltered") ans = xatlas(false);
    else if (name == "weCall") ans = wecall(true);
    else if (name == "weCall_unfiltered") ans = wecall(false);
    else if (name == "DeepVariant" || name == "DeepVariantWGS") ans = deepvariant(true, 0.6f, 0.45f, "MIN_DP");
    else if (name == "DeepVariantWES") ans = deepvariant(true, 0.375f, 0.375f, "MIN_DP");
    else if (name == "DeepVariantWES_MED_DP") ans = deepvariant(true, 0.375f, 0.375f, "MED_DP");
    else if (name == "DeepVariant_unfiltered") ans = deepvariant(false, 1.0f, 1.0f, "MIN_DP");
    else if (name == "Strelka2") ans = strelka2();
    else if (name == "GxS") ans = gxs();
    else return Status::NotFound("unknown configuration preset", name);
    return Status::OK();
}

}  // namespace glx
