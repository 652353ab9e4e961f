#!/usr/bin/env python3
"""Generate tests/golden/*.json from the reference's own golden vectors.

Run in the build container (needs /root/reference; the GPU box never reads it):
    python tests/golden/make_golden.py

Sources (relative to /root/reference):
  * test/data/gvcf_test_cases/*.yml  — input gVCF text per dataset, genotyper config, truth_unified_sites and
    truth_output_vcf; the per-case FORMAT/INFO whitelists are those of test/gvcf_test_cases.cc:496-896.
  * test/genotyper.cc:25-172         — the 14 revise_genotypes known-answer cases.
  * test/diploid.cc:9-94             — diploid index / PL->log-likelihood KATs (restated in tests/test_oracle_kats.py).
The YAML is flattened to the line formats of glx::genotyper_config::of_text / glx::sites_of_text so the C++
side needs no YAML parser. Coordinates become 0-based half-open (range_of_yaml, src/types.cc:123-161).
"""
import json
import math
import os
import re
import sys

import yaml

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

# test/gvcf_test_cases.cc:496-896 — (formats, infos) compared per case; default is GT,RNC and no INFO
DEFAULT = (["GT", "RNC"], [])
DV = (["DP", "GT", "GQ", "PL", "AD", "RNC"], ["ANR", "AF", "AQ"])
WHITELIST = {
    "format_fields_combi": (["GT", "RNC", "AD", "DP", "SB"], []),
    "format_fields_AD_fallback": (["GT", "RNC", "AD"], []),
    "format_fields_default_missing": (["GT", "RNC", "AD", "SB", "SBF"], []),
    "format_fields_default_none": (["GT", "RNC", "AD", "SB"], []),
    "format_fields_ignore_non_variants": (["GT", "RNC", "DP", "SB", "AD", "GQ"], []),
    "format_fields_integrated": (["GT", "RNC", "DP", "SB", "AD", "GQ"], []),
    "lost_deletion_monoallelic_site": (["GT", "RNC", "DP", "AD"], []),
    "DP0_noAD": (["GT", "RNC", "DP", "SB", "AD", "GQ", "RNC"], []),
    "revise_overlapping": (["GT", "RNC", "GQ"], []),
    "censor_rnc_format_fields": (["DP", "GT", "RNC", "GQ", "AD"], []),
    "xAtlas": (["DP", "GT", "GQ", "PL", "P", "RR", "VR", "FT"], []),
    "weCall": (["DP", "GT", "GQ", "PL", "AD", "FT", "SBPV", "RNC"], ["ANR", "AF", "AQ"]),
    "weCall_squeeze": (["DP", "GT", "GQ", "PL", "AD", "FT", "SBPV", "RNC"], ["ANR", "AF", "AQ"]),
    "edge_spanning_deletion": (["GT", "RNC", "AD", "DP", "SB"], []),
    "edge_spanning_deletion2": (["GT", "RNC", "AD", "DP", "SB"], []),
    "edge_spanning_deletion3": (["GT", "RNC", "AD", "DP", "SB"], []),
    "deepvariant": DV,
    "deepvariant2": DV,
    "deepvariant_norm": DV,
    "dv_1000G_chr21_5583275": DV,
    "dv_1000G_chr17_2517459": DV,
    "dv_1000G_chr17_8375536": DV,
    "strelka2": (["DP", "GT", "GQ", "AD", "FT"], ["AF", "AQ"]),
    "star_allele": (["GT", "RNC", "DP", "SB", "AD", "GQ", "PL"], []),
    "trim_uncalled_alleles": (["DP", "GT", "GQ", "PL", "AD"], ["ANR", "AF", "AQ"]),
}
# cases whose ingest skips validate_bcf (GVCFTestCase(..., validate_on_ingest=false))
NO_VALIDATE = {"discover_alleles_gvcf_bogus"}


def b(x):
    return "1" if x else "0"


def fnum(x):
    # yaml-cpp reads the scalar text with operator>>(float); PyYAML gives a double. repr() round-trips the
    # double and strtof of that text is the correctly rounded float of the original decimal.
    return repr(float(x))


def config_text(g):
    """genotyper_config::of_yaml (src/types.cc:854-967) -> line format."""
    lines = []
    for k in ("revise_genotypes", "allow_partial_data", "output_residuals", "more_PL", "squeeze",
              "trim_uncalled_alleles", "top_two_half_calls"):
        if k in g:
            lines.append(f"{k} {b(g[k])}")
    for k in ("min_assumed_allele_frequency", "snv_prior_calibration", "indel_prior_calibration"):
        if k in g:
            lines.append(f"{k} {fnum(g[k])}")
    if "required_dp" in g:
        lines.append(f"required_dp {int(g['required_dp'])}")
    for k in ("allele_dp_format", "ref_dp_format", "output_format"):
        if k in g:
            lines.append(f"{k} {g[k]}")
    for f in g.get("liftover_fields", []) or []:
        # retained_format_field::of_yaml (src/types.cc:693-812)
        number = f["number"]
        lines.append("\t".join([
            "field", f["name"], ",".join(str(x) for x in f["orig_names"]), f.get("from", "format"), f["type"], number,
            str(int(f.get("count", 0))) if number == "basic" else "0", f.get("default_type", "missing"),
            f["combi_method"], b(f.get("ignore_non_variants", False)), f["description"]]))
    return "\n".join(lines) + "\n"


def sites_text(sites, default_contig=None):
    """unified_site::of_yaml (src/types.cc:409-511) -> line format, 0-based half-open."""
    lines = []
    for s in sites:
        rng = s["range"]
        contig = str(rng.get("ref", default_contig))
        beg, end = int(rng["beg"]) - 1, int(rng["end"])
        lines.append("\t".join(["site", contig, str(beg), str(end), str(int(s["quality"])), b(s.get("monoallelic", False)),
                                fnum(s.get("lost_allele_frequency", 0.0))]))
        for al in s["alleles"]:
            dna = str(al["dna"])
            nb, ne, nd = beg, end, dna
            if "normalized" in al:
                nb, ne = int(al["normalized"]["range"]["beg"]) - 1, int(al["normalized"]["range"]["end"])
                nd = str(al["normalized"]["dna"])
            freq = "nan" if "frequency" not in al else fnum(al["frequency"])
            lines.append("\t".join(["allele", dna, str(nb), str(ne), nd, str(int(al.get("quality", 0))), freq]))
        for u in s.get("unification", []) or []:
            lines.append("\t".join(["unif", str(int(u["range"]["beg"]) - 1), str(int(u["range"]["end"])), str(u["dna"]), str(int(u["to"]))]))
    return "\n".join(lines) + "\n"


def convert_case(path):
    name = os.path.splitext(os.path.basename(path))[0]
    y = yaml.safe_load(open(path))
    if "truth_output_vcf" not in y:
        return None
    header = y["input"]["header"]
    datasets = []
    for d in y["input"]["body"]:
        for fn, body in d.items():
            # GVCFTestCase::write_gvcfs: header + "\t" + body; dataset name = file stem (test/utils.cc:108)
            datasets.append({"name": fn[: fn.rfind(".")], "vcf": header + "\t" + body})
    truth = None
    for d in y["truth_output_vcf"]:
        for fn, body in d.items():
            truth = body
    fm, inf = WHITELIST.get(name, DEFAULT)
    ans = {
        "name": name,
        "source": f"test/data/gvcf_test_cases/{name}.yml",
        "preset": y.get("config_preset"),
        # genotyper_config::of_yaml resets to defaults first, so a genotyper_config node replaces a preset wholesale
        "config_text": config_text(y["genotyper_config"]) if "genotyper_config" in y else None,
        "sites_text": sites_text(y["truth_unified_sites"]),
        "datasets": datasets,
        "truth_vcf": truth,
        "formats": fm,
        "infos": inf,
        "validate_on_ingest": name not in NO_VALIDATE,
    }
    if ans["preset"] is None and ans["config_text"] is None:
        ans["config_text"] = "\n"  # default genotyper_config()
    return ans


def convert_revise_kats():
    src = open(os.path.join(REF, "test/genotyper.cc")).read()
    raw = re.findall(r'R"\(\n(.*?)\)";', src, re.S)
    cfg_yml, us_yml, us_yml2 = raw[0], raw[1], raw[2]
    header_txt = re.search(r'R"eof\(\n(.*?)\)eof"', src, re.S).group(1)
    assert "revise_genotypes" in cfg_yml and "alleles" in us_yml and "#CHROM" in header_txt and "0.67" in us_yml2
    split_at = src.index("us_yml2")
    cases = []
    for m in re.finditer(r'REVISE_GENOTYPES_CASE\((\d+), (\d+), (\d+), "(.*?)"\);', src):
        a1, a2, gq, rec = int(m.group(1)), int(m.group(2)), int(m.group(3)), m.group(4)
        site = us_yml if m.start() < split_at else us_yml2
        cases.append({"a1": a1, "a2": a2, "gq": gq, "vcf": header_txt + rec.replace("\\t", "\t") + "\n",
                      "sites_text": sites_text([yaml.safe_load(site)])})
    assert len(cases) == 14, len(cases)
    return {"source": "test/genotyper.cc:25-172", "config_text": config_text(yaml.safe_load(cfg_yml)), "cases": cases}


def convert_presets():
    """The `glnexus_cli --config` presets (src/cli_utils.cc:465-1188): the YAML literal parsed with PyYAML; per preset the
    genotyper_config half in the line format (config_text) and as plain values for a field-by-field check."""
    src = open(os.path.join(REF, "src/cli_utils.cc")).read()
    m = re.search(r'config_presets_yml = R"eof\(\n(.*?)\)eof"', src, re.S)
    presets = yaml.safe_load(m.group(1))
    out = {"source": "src/cli_utils.cc:465-1188", "presets": {}}
    for name, node in presets.items():
        g = node.get("genotyper_config", {}) or {}
        out["presets"][name] = {"config_text": config_text(g), "genotyper_config": g}
    assert len(out["presets"]) == 13, sorted(out["presets"])
    return out


def main():
    import glob
    n = 0
    for p in sorted(glob.glob(os.path.join(REF, "test/data/gvcf_test_cases/*.yml"))):
        c = convert_case(p)
        if c is None:
            continue
        json.dump(c, open(os.path.join(OUT, c["name"] + ".json"), "w"), indent=1)
        n += 1
    json.dump(convert_revise_kats(), open(os.path.join(OUT, "_revise_genotypes_kats.json"), "w"), indent=1)
    json.dump(convert_presets(), open(os.path.join(OUT, "_presets.json"), "w"), indent=1)
    print(f"wrote {n} fixture cases + revise_genotypes KATs to {OUT}")


if __name__ == "__main__":
    main()


def convert_real_slice():
    """A slice of the reference's real GATK gVCF (test/data/NA12878.g.vcf.gz, used by its storage tests) as a 3-sample
    cohort: the slice itself, a copy with every variant call made hom-ALT and every 7th band dropped, and a copy without
    variant records. Sites = one per variant record (its DNA alleles). No truth pVCF exists for it: the tests compare the
    kernels with the oracle on real record shapes (long bands, real PL/AD/SB, INFO fields, multi-ALT records)."""
    import gzip
    lines = [l.rstrip("\n") for l in gzip.open(os.path.join(REF, "test/data/NA12878.g.vcf.gz"), "rt")]
    header = [l for l in lines if l.startswith("##")]
    recs = [l for l in lines if not l.startswith("#")]
    window = recs[240000:242200]
    chrom = window[0].split("\t")[0]
    window = [l for l in window if l.split("\t")[0] == chrom]
    hdr_txt = "\n".join(h for h in header if not h.startswith("##contig") or f"ID={chrom}," in h) + "\n"
    col = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t"
    sites, a_rows, b_rows, c_rows = [], [], [], []
    last_end = -1
    for i, l in enumerate(window):
        t = l.split("\t")
        alts = [a for a in t[4].split(",") if not a.startswith("<")]
        a_rows.append(l)
        if alts:
            beg, end = int(t[1]) - 1, int(t[1]) - 1 + len(t[3])
            if beg >= last_end:
                sites.append((beg, end, t[3], alts))
                last_end = end
            fmt = t[8].split(":")
            vals = t[9].split(":")
            vals[fmt.index("GT")] = "1/1"
            b_rows.append("\t".join(t[:9] + [":".join(vals)]))
        else:
            if i % 7:
                b_rows.append(l)
            c_rows.append(l)
    st = []
    for k, (beg, end, ref, alts) in enumerate(sites):
        st.append("\t".join(["site", chrom, str(beg), str(end), str(100 + k % 800), "0", "0.001"]))
        st.append("\t".join(["allele", ref, str(beg), str(end), ref, "0", "nan"]))
        for j, a in enumerate(alts):
            st.append("\t".join(["allele", a, str(beg), str(end), a, str(50 + j), repr(0.3 / (j + 1))]))
    ans = {"source": "test/data/NA12878.g.vcf.gz records 240000-242200", "preset": "gatk", "sites_text": "\n".join(st) + "\n",
           "datasets": [{"name": n, "vcf": hdr_txt + col + n + "\n" + "\n".join(rows) + "\n"} for n, rows in (("NA12878", a_rows), ("NA12878_homalt", b_rows), ("NA12878_bands", c_rows))]}
    return ans


if __name__ == "__main__":
    json.dump(convert_real_slice(), open(os.path.join(OUT, "_real_na12878_slice.json"), "w"))
