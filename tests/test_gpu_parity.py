"""Parity tests proper: the CUDA path, called through the C ABI (glx_genotype_batch / glx_upload+launch+download),
against the oracle — bit-exact on every output column — and against the reference's truth pVCFs."""
import json
import os

import numpy as np
import pytest

import golden_util as gu
import oracle_lib
import synth_cases
import vcfcmp
from test_emu_parity import compare_outs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    from glnexus_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", gu.case_names())
def test_gpu_golden(ctx, name):
    from glnexus_b200 import GlxError
    case = gu.load_case(name)
    texts = []
    for batch in gu.batches_of_case(case):
        ref, got = batch.alloc_out(), batch.alloc_out(pinned=True)
        if not batch.device_supported:
            with pytest.raises(GlxError) as ei:
                ctx.genotype_batch(batch, got)
            assert ei.value.code == -101  # rejected, not mis-computed (a dataset only partly in the sample set)
            return
        oracle_lib.genotype_batch(batch, ref)
        ctx.genotype_batch(batch, got)
        d = compare_outs(got, ref, batch.n_fields)
        assert not d, "\n".join(d)
        texts.append(batch.assemble_vcf(got))
    diffs = vcfcmp.compare(gu.merge_vcf_texts(texts), case["truth_vcf"], case["formats"], case["infos"])
    assert not diffs, "\n".join(diffs[:20])


KATS = json.load(open(os.path.join(gu.GOLDEN, "_revise_genotypes_kats.json")))


@pytest.mark.parametrize("path", ["sig", "dynamic", "general"])
@pytest.mark.parametrize("i", range(len(KATS["cases"])))
def test_gpu_revise_genotypes_kats(ctx, i, path):
    """The 14 revise_genotypes known answers of test/genotyper.cc:110-172 through glx_genotype_batch on the GPU: every
    column equals the oracle's genotype_site, and where the revised call is representable at the site GT and GQ equal the
    reference's expected (allele1, allele2, GQ)."""
    from glnexus_b200 import Batch
    case = KATS["cases"][i]
    b = Batch.from_text(case["sites_text"], [("A", case["vcf"])], config_text=KATS["config_text"])
    assert oracle_lib.revise_kat(b) == (case["a1"], case["a2"], case["gq"])
    ref, got = b.alloc_out(), b.alloc_out(pinned=True)
    oracle_lib.genotype_batch(b, ref)
    ctx.select_path(path)
    try:
        ctx.genotype_batch(b, got)
    finally:
        ctx.select_path("sig")
    d = compare_outs(got, ref, b.n_fields)
    assert not d, "\n".join(d)
    if got.array("rnc").tobytes() == b"..":
        assert sorted(int(x) for x in got.array("gt")[:2]) == sorted([case["a1"], case["a2"]])
        assert int(got.array(0)[0]) == case["gq"]  # the KAT config lifts GQ only


def test_gpu_diploid_kats(ctx):
    """test/diploid.cc:9-39 on the device's index math, exhaustive for n_allele <= 32."""
    for n_allele in range(1, 33):
        got = ctx.selftest_diploid(n_allele)
        want, x = [], 0
        for j in range(n_allele):
            for i in range(j + 1):
                want.append((x, x))
                x += 1
        assert got == want, n_allele


@pytest.mark.parametrize("shape,n_samples,n_sites,region_len", [
    ("c2", 1000, 3000, 3000),
    ("c3", 1500, 2000, 128000),
    ("c4", 2000, 1500, 96000),
    ("c5", 600, 2000, 60000),
])
@pytest.mark.parametrize("path", ["sig", "sig-unfused", "dynamic", "general"])
def test_gpu_synth(ctx, shape, n_samples, n_sites, region_len, path):
    """Signature-specialised tiled kernel (default), the dynamic tiled kernel, and the general kernel alone must all
    equal the oracle on every output column."""
    ctx.select_path(path)  # glx_set_option switches; applies to batches uploaded from here on
    force_general = path == "general"
    b = synth_cases.make(shape, n_samples=n_samples, n_sites=n_sites, region_len=region_len)
    ref, got = b.alloc_out(), b.alloc_out(pinned=True)
    oracle_lib.genotype_batch(b, ref, threads=8)
    dev = ctx.upload(b)
    dev.launch()
    dev.check()
    dev.download(got)
    assert dev.launch_count >= (1 if force_general else 5)  # resolve + per chunk: tiled, single-record, general
    dev.close()
    ctx.select_path("sig")
    d = compare_outs(got, ref, b.n_fields)
    assert not d, "\n".join(d)


@pytest.mark.parametrize("preset", ["gatk_unfiltered", "DeepVariant_unfiltered", "DeepVariantWES", "weCall", "Strelka2", "xAtlas"])
def test_gpu_presets(ctx, preset):
    from glnexus_b200 import GlxError
    b = synth_cases.make("c3", n_samples=300, n_sites=800, region_len=24000, preset=preset, frac_two_record_cells=0.05,
                         frac_lost_allele=0.05, frac_homalt=0.2)
    ref, got = b.alloc_out(), b.alloc_out(pinned=True)
    try:
        oracle_lib.genotype_batch(b, ref, threads=8)
        orc = 0
    except oracle_lib.OracleError as e:
        orc = e.code
    if orc != 0:
        with pytest.raises(GlxError) as ei:
            ctx.genotype_batch(b, got)
        assert ei.value.code == orc
        return
    ctx.genotype_batch(b, got)
    d = compare_outs(got, ref, b.n_fields)
    assert not d, "\n".join(d)


def test_gpu_empty_and_ragged(ctx):
    """No sites; a cohort where many samples have no records at all in the range."""
    b = synth_cases.make("c2", n_samples=33, n_sites=0, region_len=500)
    got = b.alloc_out(pinned=True)
    ctx.genotype_batch(b, got)
    b = synth_cases.make("c2", n_samples=257, n_sites=400, region_len=400, frac_uncovered=0.9)
    ref, got = b.alloc_out(), b.alloc_out(pinned=True)
    oracle_lib.genotype_batch(b, ref, threads=4)
    ctx.genotype_batch(b, got)
    d = compare_outs(got, ref, b.n_fields)
    assert not d, "\n".join(d)


def test_gpu_pipelined_async_api(ctx):
    """glx_upload_on / glx_launch / glx_download_on on two streams, two batches in flight (the e2e mode of bench.py):
    every batch's columns must still equal the oracle, and glx_status must report clean error words."""
    batches = [synth_cases.make("c3", n_samples=512, n_sites=1500, region_len=96000, seed=20200211 + i).pin() for i in range(3)]
    refs = []
    for b in batches:
        r = b.alloc_out()
        oracle_lib.genotype_batch(b, r, threads=8)
        refs.append(r)
    outs = [b.alloc_out(pinned=True) for b in batches]
    streams = [ctx.stream_create(), ctx.stream_create()]
    devs = []
    for rep in range(2):  # second round re-uses the cached arenas
        devs = []
        for i, b in enumerate(batches):
            st = streams[i & 1]
            d = ctx.upload_async(b, st)
            d.launch(st)
            d.download_async(outs[i], st)
            devs.append(d)
        for st in streams:
            ctx.stream_sync(st)
        for i, d in enumerate(devs):
            d.status()
            d.close()
            diffs = compare_outs(outs[i], refs[i], batches[i].n_fields)
            assert not diffs, f"batch {i} round {rep}:\n" + "\n".join(diffs)
    for st in streams:
        ctx.stream_destroy(st)


def test_gpu_error_status_propagates(ctx):
    """A malformed record (PL vector of the wrong length on a revised call) must fail the batch with the oracle's status."""
    from glnexus_b200 import Batch, GlxError
    hdr = ("##fileformat=VCFv4.2\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"\">\n##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"\">\n"
           "##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"\">\n##FORMAT=<ID=AD,Number=R,Type=Integer,Description=\"\">\n"
           "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"\">\n##contig=<ID=21,length=48129895>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\n")
    rec = "21\t1000\t.\tT\tA\t50\t.\t.\tGT:GQ:DP:AD:PL\t1/1:30:20:0,20:100,30\n"  # 2 PL values, 3 expected
    sites = "site\t21\t999\t1000\t100\t0\t0.0\nallele\tT\t999\t1000\tT\t0\tnan\nallele\tA\t999\t1000\tA\t88\t0.01\n"
    b = Batch.from_text(sites, [("A", hdr + rec)], preset="DeepVariant")
    ref, got = b.alloc_out(), b.alloc_out(pinned=True)
    with pytest.raises(oracle_lib.OracleError) as oe:
        oracle_lib.genotype_batch(b, ref)
    with pytest.raises(GlxError) as ge:
        ctx.genotype_batch(b, got)
    assert ge.value.code == oe.value.code == -1


def test_gpu_full_size_configs1(ctx):
    """BASELINE configs[1] at FULL size (1,000 samples x 50,000 sites = 5e7 cells, the bench workload): every output
    column bit-exact against the oracle (all host threads), plus size-independent properties of the columns."""
    import os
    b = synth_cases.make("c2", n_samples=1000, n_sites=50000, region_len=50000)
    ref, got = b.alloc_out(), b.alloc_out(pinned=True)
    oracle_lib.genotype_batch(b, ref, threads=max(4, len(os.sched_getaffinity(0))))
    ctx.genotype_batch(b, got)
    ga, ra = got.arrays(), ref.arrays()
    for k in ra:
        assert np.array_equal(ga[k], ra[k]), k
    gt = ga["gt"].reshape(-1, 2)
    rnc = ga["rnc"].reshape(-1, 2)
    assert (gt[:, 0] <= gt[:, 1]).all()                       # slots ordered, missing first (genotyper.cc:862-867)
    assert ((gt >= 0) == (rnc == ord("."))).all()               # an allele is called exactly where RNC is n/a
    used = np.zeros(b.n_sites, dtype=np.uint32)                 # allele_used = OR over the site's called alleles
    g = ga["gt"].reshape(b.n_sites, -1)
    for a in range(3):
        used |= ((g == a).any(axis=1).astype(np.uint32) << a)
    assert np.array_equal(used, ga["allele_used"])


@pytest.mark.parametrize("shape,n_samples,n_sites,region_len", [
    ("c3", 10000, 4096, 262144),   # one streamed batch of configs[2]
    ("c4", 100000, 512, 32768),    # one streamed batch of configs[3]
    ("c5", 5000, 10000, 300000),   # a tenth of configs[4] (gatk preset, up to 7 alleles)
])
def test_gpu_full_size_batches(ctx, shape, n_samples, n_sites, region_len):
    """The other BASELINE shapes at the batch sizes bench.py runs (4e7-5e7 cells each): every output column bit-exact
    against the oracle (all host threads), plus the size-independent column properties."""
    import os
    b = synth_cases.make(shape, n_samples=n_samples, n_sites=n_sites, region_len=region_len)
    ref, got = b.alloc_out(), b.alloc_out(pinned=True)
    oracle_lib.genotype_batch(b, ref, threads=max(4, len(os.sched_getaffinity(0))))
    ctx.genotype_batch(b, got)
    ga, ra = got.arrays(), ref.arrays()
    for k in ra:
        assert np.array_equal(ga[k], ra[k]), k
    gt = ga["gt"].reshape(-1, 2)
    rnc = ga["rnc"].reshape(-1, 2)
    assert (gt[:, 0] <= gt[:, 1]).all()
    assert ((gt >= 0) == (rnc == ord("."))).all()
    used = np.zeros(b.n_sites, dtype=np.uint32)
    g = ga["gt"].reshape(b.n_sites, -1)
    for a in range(int(g.max()) + 1):
        used |= ((g == a).any(axis=1).astype(np.uint32) << a)
    assert np.array_equal(used, ga["allele_used"])


def test_gpu_narrow_download(ctx):
    """glx_download_narrow_on: int16 columns (BCF 16-bit sentinels) where a field fits, int32 where it does not; widened
    on the host they must equal the oracle's int32 columns exactly."""
    from glnexus_b200 import Batch
    b = synth_cases.make("c3", n_samples=768, n_sites=1200, region_len=76800).pin()
    ref = b.alloc_out()
    oracle_lib.genotype_batch(b, ref, threads=8)
    st = ctx.stream_create()
    o16, wide = b.alloc_out16(), b.alloc_out()
    dev = ctx.upload_async(b, st)
    dev.launch(st)
    dev.download_narrow_async(o16, st)
    ctx.stream_sync(st)
    dev.status()
    dev.narrow_finish(o16)
    dev.close()
    o16.widen(wide)
    d = compare_outs(wide, ref, b.n_fields)
    assert not d, "\n".join(d)
    wide_bytes = sum(a.nbytes for k, a in ref.arrays().items() if k not in ("allele_used", "site_flags"))
    n16 = o16.narrow_nbytes()
    assert n16 < 0.6 * wide_bytes
    # the next batch of the same context is sent in int8 wherever the previous one fitted BCF's 8-bit range (GQ at least)
    o8, wide8 = b.alloc_out16(), b.alloc_out()
    dev = ctx.upload_async(b, st)
    dev.launch(st)
    dev.download_narrow_async(o8, st)
    ctx.stream_sync(st)
    dev.status()
    dev.narrow_finish(o8)
    dev.close()
    o8.widen(wide8)
    d = compare_outs(wide8, ref, b.n_fields)
    assert not d, "\n".join(d)
    assert o8.narrow_nbytes() < n16
    # a field with a value beyond the width it is sent in (here: beyond int16, and beyond the int8 the previous batch
    # suggested) falls back to its int32 column
    hdr = ("##fileformat=VCFv4.2\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"\">\n##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"\">\n"
           "##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"\">\n##FORMAT=<ID=AD,Number=R,Type=Integer,Description=\"\">\n"
           "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"\">\n##contig=<ID=21,length=48129895>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\n")
    rec = "21\t1000\t.\tT\tA\t50\t.\t.\tGT:GQ:DP:AD:PL\t0/1:30:100000:50000,50000:100,0,70000\n"
    sites = "site\t21\t999\t1000\t100\t0\t0.0\nallele\tT\t999\t1000\tT\t0\tnan\nallele\tA\t999\t1000\tA\t88\t0.01\n"
    b2 = Batch.from_text(sites, [("A", hdr + rec)], preset="DeepVariant")
    ref2 = b2.alloc_out()
    oracle_lib.genotype_batch(b2, ref2)
    o16, wide = b2.alloc_out16(), b2.alloc_out()
    dev = ctx.upload_async(b2, st)
    dev.launch(st)
    dev.download_narrow_async(o16, st)
    ctx.stream_sync(st)
    dev.status()
    dev.narrow_finish(o16)
    dev.close()
    o16.widen(wide)
    d = compare_outs(wide, ref2, b2.n_fields)
    assert not d, "\n".join(d)
    assert 100000 in wide.array(0) and 70000 in wide.array(3)
    ctx.stream_destroy(st)


@pytest.mark.parametrize("fmt", ["BCF", "VCF"])
@pytest.mark.parametrize("name", gu.case_names())
def test_gpu_service_genotype_sites(name, fmt, tmp_path):
    """The upper seam (glx::genotype_sites = Service::genotype_sites on the device path): whole cases, all contigs, tiny range
    batches (2 sites) so that batching, pipelining and ordered writing are exercised. Output follows cfg.output_format like the
    reference's BCFFileSink: BCF (default) = a BGZF-compressed BCF2 file encoded on the device, read back here with the
    independent BCF reader; VCF = pVCF text. Both against the reference's truth pVCF."""
    import bcfdec
    import glnexus_b200
    case = gu.load_case(name)
    ds = [(d["name"], d["vcf"]) for d in case["datasets"]]
    out = str(tmp_path / ("out." + fmt.lower()))
    cfg_text = glnexus_b200.load_config(case["preset"], case["config_text"])
    assert "output_format BCF" in cfg_text  # the reference's default
    cfg_text = cfg_text.replace("output_format BCF", "output_format " + fmt)
    # (vcf_services holds multi-sample gVCFs: its batches take glx_cells_multi_kernel)
    st = glnexus_b200.genotype_sites(case["sites_text"], ds, out, config_text=cfg_text, batch_sites=2, test_ingest_rules=True)
    raw = open(out, "rb").read()
    if fmt == "BCF":
        assert raw[:4] == b"\x1f\x8b\x08\x04" and raw.endswith(glnexus_b200.bgzf_eof())
        got = bcfdec.bcf_file_to_vcf_text(raw)
    else:
        got = raw.decode()
    diffs = vcfcmp.compare(got, case["truth_vcf"], case["formats"], case["infos"])
    assert not diffs, "\n".join(diffs[:20])
    n_sites = sum(1 for l in case["sites_text"].split("\n") if l.startswith("site\t"))
    assert st["sites"] == n_sites and st["batches"] >= (n_sites + 1) // 2
    assert st["rows"] == sum(1 for l in got.split("\n") if l and not l.startswith("#"))


def test_gpu_service_abort_and_errors(tmp_path):
    import ctypes
    import glnexus_b200
    case = gu.load_case("dv_1000G_chr21_5583275")
    ds = [(d["name"], d["vcf"]) for d in case["datasets"]]
    flag = ctypes.c_int(1)
    with pytest.raises(glnexus_b200.GlxError) as ei:
        glnexus_b200.genotype_sites(case["sites_text"], ds, str(tmp_path / "a.vcf"), preset=case["preset"], config_text=case["config_text"], abort=flag)
    assert ei.value.code == -7  # Status::Aborted
    with pytest.raises(glnexus_b200.GlxError) as ei:
        glnexus_b200.genotype_sites(case["sites_text"], ds, "/nonexistent-dir/x.vcf", preset=case["preset"], config_text=case["config_text"])
    assert ei.value.code == -5  # IOError


def test_gpu_real_gvcf_slice(ctx):
    """Real GATK gVCF records (slice of the reference's NA12878.g.vcf.gz fixture) through the CUDA path vs the oracle."""
    from test_emu_synth import _real_slice_batch
    b = _real_slice_batch()
    ref, got = b.alloc_out(), b.alloc_out(pinned=True)
    oracle_lib.genotype_batch(b, ref)
    ctx.genotype_batch(b, got)
    d = compare_outs(got, ref, b.n_fields)
    assert not d, "\n".join(d)
    vcf = b.assemble_vcf(got)
    assert vcf.count("\n") > b.n_sites  # header + one row per site (gatk preset does not trim)
