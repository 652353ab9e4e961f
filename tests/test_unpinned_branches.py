"""Hand-built cases that reach the branches DESIGN.md lists as "parity unpinned" (no reference fixture exercises them): the
kernels must agree with the oracle there too — on the CPU through the host build of the device headers, on the GPU through
the C ABI.
  (a) whole-site drop after allele trimming (src/genotyper.cc:963-965)
  (b) two alleles of one record mapping to the same unified allele
  (c) std::sort of half-call tuples when QUAL is NaN (src/genotyper.cc:497)
  (d) revise_genotypes with fewer than two finite likelihoods"""
import numpy as np
import pytest

import emu_lib
import oracle_lib
from test_emu_parity import compare_outs

HDR = """##fileformat=VCFv4.2
##FILTER=<ID=PASS,Description="All filters passed">
##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">
##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="">
##FORMAT=<ID=DP,Number=1,Type=Integer,Description="">
##FORMAT=<ID=MIN_DP,Number=1,Type=Integer,Description="">
##FORMAT=<ID=AD,Number=R,Type=Integer,Description="">
##FORMAT=<ID=PL,Number=G,Type=Integer,Description="">
##INFO=<ID=END,Number=1,Type=Integer,Description="">
##contig=<ID=chrU,length=100000>
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t"""


def vcf(name, rows):
    return (name, HDR + name + "\n" + "\n".join(rows) + "\n")


def band(pos1, end, gq=40, dp=30):
    return f"chrU\t{pos1}\t.\tA\t<*>\t.\t.\tEND={end}\tGT:GQ:MIN_DP:PL\t0/0:{gq}:{dp}:0,{3 * dp},{30 * dp}"


def var(pos1, ref, alts, gt, qual="50", pl=None, ad=None, dp=30):
    na = 2 + len(alts)
    ng = na * (na + 1) // 2
    pl = pl or ",".join(["0"] + ["50"] * (ng - 1))
    ad = ad or ",".join(["10"] * na)
    return f"chrU\t{pos1}\t.\t{ref}\t{','.join(alts)},<*>\t{qual}\t.\t.\tGT:GQ:DP:AD:PL\t{gt}:30:{dp}:{ad}:{pl}"


def site(beg, end, ref, alts, afs=None, extra_unif=()):
    lines = ["\t".join(["site", "chrU", str(beg), str(end), "100", "0", "0.001"]), "\t".join(["allele", ref, str(beg), str(end), ref, "0", "nan"])]
    for i, a in enumerate(alts):
        lines.append("\t".join(["allele", a, str(beg), str(end), a, "77", repr((afs or [0.1] * len(alts))[i])]))
    for (ub, ue, dna, to) in extra_unif:
        lines.append("\t".join(["unif", str(ub), str(ue), dna, str(to)]))
    return lines


def cases():
    out = {}
    # (a) an ALT nobody carries, trimming on: the site must vanish; next to a site that stays
    s = site(1000, 1001, "A", ["G"]) + site(1010, 1011, "A", ["C"])
    ds = [vcf("S0", [band(990, 1009), var(1011, "A", ["C"], "0/1"), band(1012, 1040)]), vcf("S1", [band(990, 1040)])]
    out["site_dropped_after_trim"] = ("\n".join(s) + "\n", ds, "DeepVariant", {})
    # (b) record ALTs G and C both unified into allele 1 (C is listed as another representation of it)
    s = site(2000, 2001, "A", ["G"], extra_unif=[(2000, 2001, "C", 1)])
    ds = [vcf("S0", [band(1990, 2000), var(2001, "A", ["G", "C"], "1/2", pl="90,40,50,30,0,60,99,99,99,99", ad="3,9,9,0"), band(2002, 2040)]),
          vcf("S1", [band(1990, 2000), var(2001, "A", ["G", "C"], "0/2", pl="40,90,99,0,70,80,99,99,99,99", ad="9,0,9,0"), band(2002, 2040)]),
          vcf("S2", [band(1990, 2040)])]
    out["two_alleles_one_target"] = ("\n".join(s) + "\n", ds, "gatk_unfiltered", {})
    out["two_alleles_one_target_revise"] = ("\n".join(s) + "\n", ds, "DeepVariant", {})
    # (c) three overlapping half-call records whose QUALs include NaN
    s = site(3000, 3001, "A", ["G", "C", "T"])
    rows = [band(2990, 3000), var(3001, "A", ["G"], "0/1", qual="nan"), var(3001, "A", ["C"], "0/1", qual="70"), var(3001, "A", ["T"], "0/1", qual="nan"), band(3002, 3040)]
    ds = [vcf("S0", rows), vcf("S1", [band(2990, 3040)])]
    out["nan_qual_half_calls"] = ("\n".join(s) + "\n", ds, "gatk_unfiltered", {})
    out["nan_qual_half_calls_top2"] = ("\n".join(s) + "\n", ds, "gatk_unfiltered", {"top_two_half_calls": 1})
    # (d) a hom-ALT call (triggers revise_genotypes) with a single finite likelihood
    s = site(4000, 4001, "A", ["G"])
    ds = [vcf("S0", [band(3990, 4000), var(4001, "A", ["G"], "1/1", pl=".,.,0,.,.,."), band(4002, 4040)]), vcf("S1", [band(3990, 4040)])]
    out["one_finite_likelihood"] = ("\n".join(s) + "\n", ds, "DeepVariant", {})
    return out


CASES = cases()


def build(name):
    from glnexus_b200 import Batch, preset_text
    sites_text, ds, preset, toggles = CASES[name]
    cfg = []
    for line in preset_text(preset).split("\n"):
        key = line.split(" ")[0]
        cfg.append(f"{key} {toggles[key]}" if key in toggles else line)
    return Batch.from_text(sites_text, ds, config_text="\n".join(cfg))


def oracle(batch):
    ref = batch.alloc_out()
    try:
        oracle_lib.genotype_batch(batch, ref)
        return 0, ref
    except oracle_lib.OracleError as e:
        return e.code, ref


@pytest.mark.parametrize("name", sorted(CASES))
def test_unpinned_branch_device_logic(built, name):
    b = build(name)
    orc, ref = oracle(b)
    got = b.alloc_out()
    rc = emu_lib.general_batch(b, got)
    assert rc == orc
    if orc == 0:
        d = compare_outs(got, ref, b.n_fields)
        assert not d, "\n".join(d)
        got2 = b.alloc_out()
        assert emu_lib.tiled_batch(b, got2)[0] == 0
        d = compare_outs(got2, ref, b.n_fields)
        assert not d, "\n".join(d)


def test_the_cases_reach_their_branches(built):
    """The fixtures do what their names say (so the comparisons above are not vacuous)."""
    b = build("site_dropped_after_trim")
    orc, ref = oracle(b)
    assert orc == 0
    rows = [l for l in b.assemble_vcf(ref).split("\n") if l and not l.startswith("#")]
    assert len(rows) == 1 and rows[0].split("\t")[1] == "1011"  # the site at 1001 is gone
    stream, off = oracle_lib.bcf_records(b, ref)
    assert off[1] == off[0] and off[2] > off[1]
    b = build("two_alleles_one_target")
    orc, ref = oracle(b)
    assert orc == 0 and ref.array("gt")[:2].tolist() == [1, 1]  # 1/2 of the record = G/C = unified 1/1
    b = build("nan_qual_half_calls")
    orc, ref = oracle(b)
    assert orc == 0 and bytes(ref.array("rnc")[:2]) in (b"O.", b".O", b"OO")
    b = build("one_finite_likelihood")
    orc, _ = oracle(b)
    assert orc == -1  # Failure: the reference asserts here; the oracle and the kernels fail the site


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["sig", "dynamic", "general"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_unpinned_branch_gpu(built, name, path):
    from glnexus_b200 import Context, GlxError
    b = build(name)
    orc, ref = oracle(b)
    ctx = Context(0)
    ctx.select_path(path)
    got = b.alloc_out(pinned=True)
    try:
        ctx.genotype_batch(b, got)
        rc = 0
    except GlxError as e:
        rc = e.code
    assert rc == orc
    if orc == 0:
        d = compare_outs(got, ref, b.n_fields)
        assert not d, "\n".join(d)
        # and through the BCF encoder: dropped sites have no record
        dev = ctx.upload(b)
        dev.launch()
        dev.check()
        dev.bcf_attach(b)
        dev.encode_bcf(0)
        want, woff = oracle_lib.bcf_records(b, ref)
        assert dev.download_bcf() == want and dev.bcf_record_offsets() == woff
        dev.close()
    ctx.close()
