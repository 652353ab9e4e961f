"""Pin the BCF2 record oracle (oracle/glx_bcf_oracle.cc): on every golden case the records it writes, decoded by an
independent BCF2 reader (tests/bcfdec.py), must equal — string for string — the rows of the text assembler that the
reference's truth pVCFs pin, and compare equal to those truth files themselves (test/gvcf_test_cases.cc:415-461);
plus the header / BGZF framing the host writes."""
import gzip
import struct

import pytest

import bcfdec
import fuzz_cases
import golden_util as gu
import oracle_lib
import synth_cases
import vcfcmp


def rows_of_text(vcf_text):
    return [l for l in vcf_text.split("\n") if l and not l.startswith("#")]


def check_batch(batch, out):
    stream, off = oracle_lib.bcf_records(batch, out)
    assert off[0] == 0 and off[-1] == len(stream) and all(a <= b for a, b in zip(off, off[1:]))
    _, ids, contigs, samples, _ = bcfdec.parse_header(batch.bcf_prologue())
    rows = bcfdec.records_to_rows(stream, ids, contigs, batch.n_samples)
    want = rows_of_text(batch.assemble_vcf(out))
    assert len(rows) == len(want) == sum(1 for a, b in zip(off, off[1:]) if b > a)
    for i, (g, w) in enumerate(zip(rows, want)):
        assert g == w, f"row {i}:\n bcf  {g[:400]}\n text {w[:400]}"
    return stream, rows


@pytest.mark.parametrize("name", gu.case_names())
def test_bcf_oracle_golden(built, name):
    case = gu.load_case(name)
    all_rows, header = [], None
    for batch in gu.batches_of_case(case):
        out = batch.alloc_out()
        oracle_lib.genotype_batch(batch, out)
        _, rows = check_batch(batch, out)
        all_rows += rows
        text, *_ = bcfdec.parse_header(batch.bcf_prologue())
        header = header or bcfdec.vcf_header_from_bcf(text)
    if header is None:
        return
    diffs = vcfcmp.compare(header + "".join(r + "\n" for r in all_rows), case["truth_vcf"], case["formats"], case["infos"])
    assert not diffs, "\n".join(diffs[:20])


@pytest.mark.parametrize("shape,ns,nsites,region", [("c2", 40, 300, 300), ("c3", 50, 200, 12000), ("c4", 60, 150, 9000), ("c5", 40, 200, 6000)])
def test_bcf_oracle_synth(built, shape, ns, nsites, region):
    b = synth_cases.make(shape, n_samples=ns, n_sites=nsites, region_len=region)
    out = b.alloc_out()
    oracle_lib.genotype_batch(b, out, threads=4)
    stream, rows = check_batch(b, out)
    assert len(rows) > nsites // 4


@pytest.mark.parametrize("preset", ["xAtlas", "weCall", "Strelka2", "gatk_unfiltered", "DeepVariantWES"])
def test_bcf_oracle_presets(built, preset):
    """Float fields, FT strings, squeeze-free and trimmed configurations."""
    b = synth_cases.make("c3", n_samples=30, n_sites=200, region_len=6000, preset=preset, frac_two_record_cells=0.05, frac_lost_allele=0.05, frac_homalt=0.2)
    out = b.alloc_out()
    try:
        oracle_lib.genotype_batch(b, out, threads=4)
    except oracle_lib.OracleError:
        pytest.skip("oracle rejects this cohort under the preset")
    check_batch(b, out)


def test_bcf_oracle_fuzz(built):
    n = 0
    for seed in range(1000, 1080):
        b = fuzz_cases.build_batch(seed)
        out = b.alloc_out()
        try:
            oracle_lib.genotype_batch(b, out)
        except oracle_lib.OracleError:
            continue
        check_batch(b, out)
        n += 1
    assert n >= 30


def test_typed_encoding_edges(built):
    """Descriptor bytes: counts >= 15 spill into a typed integer; integer widths follow the value range after trimming."""
    b = synth_cases.make("c5", n_samples=9, n_sites=60, region_len=1800)
    out = b.alloc_out()
    oracle_lib.genotype_batch(b, out)
    stream, off = oracle_lib.bcf_records(b, out)
    # gatk PL at a site with >= 5 alleles has >= 15 values per sample: the descriptor byte is 0xF? followed by a typed count
    assert any(bytes([0xf1]) in stream[a:b2] or bytes([0xf2]) in stream[a:b2] for a, b2 in zip(off, off[1:]) if b2 > a)
    for a, e in zip(off, off[1:]):
        if e > a:
            l_shared, l_indiv = struct.unpack_from("<II", stream, a)
            assert 8 + l_shared + l_indiv == e - a


def test_prologue_and_bgzf_framing(built):
    import glnexus_b200
    b = synth_cases.make("c2", n_samples=5, n_sites=20, region_len=20)
    pro = b.bcf_prologue()
    text, ids, contigs, samples, end = bcfdec.parse_header(pro)
    assert end == len(pro) and ids[:8] == ["PASS", "AF", "AQ", "AC", "AN", "MONOALLELIC", "GT", "RNC"] and ids[8:] == ["DP", "AD", "GQ", "PL"]
    assert contigs == ["chr20"] and len(samples) == 5
    out = b.alloc_out()
    oracle_lib.genotype_batch(b, out)
    stream, _ = oracle_lib.bcf_records(b, out)
    blob = glnexus_b200.bgzf_stored(pro + stream * 40)  # several blocks
    raw, n_blocks = bcfdec.bgzf_decompress(blob)
    assert raw == pro + stream * 40 and n_blocks == (len(raw) + 65279) // 65280
    eof = glnexus_b200.bgzf_eof()
    assert eof == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")  # the BGZF EOF marker of the SAM spec
    # a whole file decodes with Python's gzip (multi-member) and with the BCF reader
    whole = glnexus_b200.bgzf_stored(pro) + glnexus_b200.bgzf_stored(stream) + eof
    assert gzip.decompress(whole) == pro + stream
    assert bcfdec.bcf_file_to_vcf_text(whole).count("\n") == text.count("\n") + len(bcfdec.records_to_rows(stream, ids, contigs))


# ---- the device encoder's stages (glx_bcf.cuh) compiled for the host vs the oracle: byte for byte ----
import emu_lib  # noqa: E402


def check_emu(batch, out, spans=(0, 997, 64)):
    want, woff = oracle_lib.bcf_records(batch, out)
    for span in spans:
        got, goff = emu_lib.bcf_records(batch, out, span=span, chunk=7 if span else 0)
        assert goff == woff, f"record offsets differ (span {span})"
        if got != want:
            i = next(i for i in range(min(len(got), len(want))) if got[i] != want[i])
            raise AssertionError(f"span {span}: first difference at byte {i} of {len(want)}: got {got[i:i+16].hex()} want {want[i:i+16].hex()}")


@pytest.mark.parametrize("name", gu.case_names())
def test_bcf_device_logic_golden(built, name):
    for batch in gu.batches_of_case(gu.load_case(name)):
        out = batch.alloc_out()
        oracle_lib.genotype_batch(batch, out)
        check_emu(batch, out)


@pytest.mark.parametrize("shape,ns,nsites,region", [("c2", 40, 300, 300), ("c3", 50, 200, 12000), ("c4", 60, 150, 9000), ("c5", 40, 200, 6000)])
def test_bcf_device_logic_synth(built, shape, ns, nsites, region):
    b = synth_cases.make(shape, n_samples=ns, n_sites=nsites, region_len=region)
    out = b.alloc_out()
    oracle_lib.genotype_batch(b, out, threads=4)
    check_emu(b, out, spans=(0, 4093))


@pytest.mark.parametrize("preset", ["xAtlas", "weCall", "Strelka2", "gatk_unfiltered", "DeepVariantWES"])
def test_bcf_device_logic_presets(built, preset):
    b = synth_cases.make("c3", n_samples=30, n_sites=200, region_len=6000, preset=preset, frac_two_record_cells=0.05, frac_lost_allele=0.05, frac_homalt=0.2)
    out = b.alloc_out()
    try:
        oracle_lib.genotype_batch(b, out, threads=4)
    except oracle_lib.OracleError:
        pytest.skip("oracle rejects this cohort under the preset")
    check_emu(b, out, spans=(0, 1021))


def test_bcf_device_logic_fuzz(built):
    n = 0
    for seed in range(1000, 1080):
        b = fuzz_cases.build_batch(seed)
        out = b.alloc_out()
        try:
            oracle_lib.genotype_batch(b, out)
        except oracle_lib.OracleError:
            continue
        check_emu(b, out, spans=(0, 333))
        n += 1
    assert n >= 30


# ---- the device BGZF stage (glx_bgzf.cuh) compiled for the host: blocks must inflate (zlib) to the oracle's stream ----
def check_bgzf(batch, out, spans=(0,)):
    want, _ = oracle_lib.bcf_records(batch, out)
    ratios = []
    for span in spans:
        blob, n_stored = emu_lib.bgzf_blocks(batch, out, span=span)
        raw, n_blocks = bcfdec.bgzf_decompress(blob)  # checks BSIZE, CRC32, ISIZE of every block
        assert raw == want, f"span {span}: inflated stream differs"
        assert n_blocks == (len(want) + (span or 65280) - 1) // (span or 65280)
        ratios.append((len(want) / max(1, len(blob)), n_stored, n_blocks))
    return ratios


@pytest.mark.parametrize("name", gu.case_names())
def test_bgzf_device_logic_golden(built, name):
    for batch in gu.batches_of_case(gu.load_case(name)):
        out = batch.alloc_out()
        oracle_lib.genotype_batch(batch, out)
        check_bgzf(batch, out, spans=(0, 300))


@pytest.mark.parametrize("shape,ns,nsites,region", [("c2", 400, 300, 300), ("c3", 500, 200, 12000), ("c4", 600, 150, 9000), ("c5", 200, 200, 6000)])
def test_bgzf_device_logic_synth(built, shape, ns, nsites, region):
    b = synth_cases.make(shape, n_samples=ns, n_sites=nsites, region_len=region)
    out = b.alloc_out()
    oracle_lib.genotype_batch(b, out, threads=4)
    r = check_bgzf(b, out, spans=(0, 20000))
    assert r[0][0] > 1.5, r  # the stream of a real cohort must actually compress
    print(shape, r)


def test_bgzf_device_logic_incompressible_and_tiny(built):
    """Random-looking payloads fall back to stored blocks; a batch with no records gives no blocks."""
    b = synth_cases.make("c5", n_samples=3, n_sites=40, region_len=1200)
    out = b.alloc_out()
    oracle_lib.genotype_batch(b, out)
    check_bgzf(b, out, spans=(0, 64, 17))
    b = synth_cases.make("c2", n_samples=4, n_sites=0, region_len=100)
    out = b.alloc_out()
    blob, _ = emu_lib.bgzf_blocks(b, out)
    assert blob == b""
