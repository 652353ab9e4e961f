"""N1 on the GPU: the device BCF2 record encoder and its BGZF stage, through the C ABI, against the oracle.
Raw mode must equal the oracle's record stream byte for byte (and decode to the reference's truth pVCFs); BGZF mode must
inflate (zlib: CRC-32, ISIZE, BSIZE checked per block) to the same stream."""
import pytest

import bcfdec
import fuzz_cases
import golden_util as gu
import oracle_lib
import synth_cases
import vcfcmp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    from glnexus_b200 import Context
    c = Context(0)
    yield c
    c.close()


def encode(ctx, batch, mode):
    dev = ctx.upload(batch)
    dev.launch()
    dev.check()
    dev.bcf_attach(batch)
    dev.encode_bcf(mode)
    n, raw_n, blocks = dev.bcf_length()
    data = dev.download_bcf()
    off = dev.bcf_record_offsets()
    dev.close()
    assert len(data) == n
    return data, raw_n, blocks, off


def check(ctx, batch):
    ref = batch.alloc_out()
    oracle_lib.genotype_batch(batch, ref, threads=8)
    want, woff = oracle_lib.bcf_records(batch, ref)
    got, raw_n, _, off = encode(ctx, batch, 0)
    assert off == woff and raw_n == len(want)
    if got != want:
        i = next(i for i in range(min(len(got), len(want))) if got[i] != want[i])
        raise AssertionError(f"raw stream differs at byte {i} of {len(want)}: got {got[i:i+16].hex()} want {want[i:i+16].hex()}")
    blob, raw_n2, blocks, _ = encode(ctx, batch, 1)
    inflated, n_blocks = bcfdec.bgzf_decompress(blob)
    assert inflated == want and raw_n2 == len(want) and n_blocks == blocks == (len(want) + 65279) // 65280
    return want, blob


@pytest.mark.parametrize("name", gu.case_names())
def test_gpu_bcf_golden(ctx, name):
    case = gu.load_case(name)
    rows, header = [], None
    for batch in gu.batches_of_case(case):
        if not batch.device_supported:
            return
        want, _ = check(ctx, batch)
        text, ids, contigs, samples, _ = bcfdec.parse_header(batch.bcf_prologue())
        header = header or bcfdec.vcf_header_from_bcf(text)
        rows += bcfdec.records_to_rows(want, ids, contigs, batch.n_samples)
    if header:
        diffs = vcfcmp.compare(header + "".join(r + "\n" for r in rows), case["truth_vcf"], case["formats"], case["infos"])
        assert not diffs, "\n".join(diffs[:20])


@pytest.mark.parametrize("shape,ns,nsites,region", [("c2", 1000, 3000, 3000), ("c3", 1500, 2000, 128000), ("c4", 2000, 1500, 96000), ("c5", 600, 2000, 60000),
                                                    ("c2", 37, 5000, 5000), ("c3", 1, 300, 9000)])
def test_gpu_bcf_synth(ctx, shape, ns, nsites, region):
    b = synth_cases.make(shape, n_samples=ns, n_sites=nsites, region_len=region)
    want, blob = check(ctx, b)
    if ns >= 600:
        assert len(blob) < len(want) / 2  # a real cohort's stream must actually compress


@pytest.mark.parametrize("preset", ["xAtlas", "weCall", "Strelka2", "gatk_unfiltered", "DeepVariantWES"])
def test_gpu_bcf_presets(ctx, preset):
    b = synth_cases.make("c3", n_samples=300, n_sites=800, region_len=24000, preset=preset, frac_two_record_cells=0.05, frac_lost_allele=0.05, frac_homalt=0.2)
    try:
        oracle_lib.genotype_batch(b, b.alloc_out(), threads=8)
    except oracle_lib.OracleError:
        pytest.skip("oracle rejects this cohort under the preset")
    check(ctx, b)


def test_gpu_bcf_fuzz(ctx):
    n = 0
    for seed in range(1000, 1100):
        b = fuzz_cases.build_batch(seed)
        try:
            oracle_lib.genotype_batch(b, b.alloc_out())
        except oracle_lib.OracleError:
            continue
        check(ctx, b)
        n += 1
    assert n >= 40


def test_gpu_bcf_empty(ctx):
    b = synth_cases.make("c2", n_samples=33, n_sites=0, region_len=500)
    got, raw_n, blocks, off = encode(ctx, b, 1)
    assert got == b"" and raw_n == 0 and blocks == 0 and off == [0]


def test_gpu_bcf_file(ctx, tmp_path):
    """A whole BCF file: host prologue block + device BGZF blocks + EOF marker decodes to the text assembler's pVCF."""
    import glnexus_b200
    b = synth_cases.make("c2", n_samples=200, n_sites=1500, region_len=1500)
    ref = b.alloc_out()
    oracle_lib.genotype_batch(b, ref, threads=8)
    blob, _, _, _ = encode(ctx, b, 1)
    whole = glnexus_b200.bgzf_stored(b.bcf_prologue()) + blob + glnexus_b200.bgzf_eof()
    got = bcfdec.bcf_file_to_vcf_text(whole)
    want = b.assemble_vcf(ref)
    assert [l for l in got.split("\n") if not l.startswith("##")] == [l for l in want.split("\n") if not l.startswith("##")]


def test_gpu_bcf_three_streams(ctx):
    """The pipelined call pattern of glx::RangeStream: attach on a copy-only stream, launch + encode on the batch's stream,
    download on a download-only stream, sizes polled with glx_bcf_ready, completion through a mark — same bytes as the oracle."""
    import glnexus_b200
    b = synth_cases.make("c2", n_samples=257, n_sites=900, region_len=900)
    b.pin()
    ref = b.alloc_out()
    oracle_lib.genotype_batch(b, ref, threads=8)
    want, _ = oracle_lib.bcf_records(b, ref)
    s_main, s_copy, s_dl = ctx.stream_create(), ctx.stream_create(), ctx.stream_create()
    try:
        for mode in (0, 1):
            dev = ctx.upload_async(b, s_main)
            dev.bcf_attach(b, stream=s_copy)
            t0 = ctx.mark(s_main)
            dev.launch(stream=s_main)
            dev.encode_bcf(mode, stream=s_main)
            while not dev.bcf_ready():
                pass
            n, raw_n, blocks = dev.bcf_length()
            buf = glnexus_b200.PinnedBuffer(n + 64)
            dev.download_bcf(buf, stream=s_dl, sync=False)
            t1 = ctx.mark(s_dl)
            ctx.mark_wait(t1)
            assert ctx.mark_ms(t0, t1) > 0
            ctx.mark_free(t0)
            ctx.mark_free(t1)
            dev.check()
            got = buf.bytes(n)
            assert raw_n == len(want)
            assert (got if mode == 0 else bcfdec.bgzf_decompress(got)[0]) == want
            buf.close()
            dev.close()
    finally:
        for s_ in (s_main, s_copy, s_dl):
            ctx.stream_destroy(s_)
