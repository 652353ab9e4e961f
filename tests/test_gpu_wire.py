"""N2 (start): the compact wire form of the record store. Uploading through it (glx_upload_wire_on: one narrow H2D copy +
expansion kernels) must leave the device with exactly the record-store arrays glx_upload_on copies, and produce the same
outputs as the oracle."""
import pytest

import fuzz_cases
import golden_util as gu
import oracle_lib
import synth_cases
from test_emu_parity import compare_outs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    from glnexus_b200 import Context
    c = Context(0)
    yield c
    c.close()


def check(ctx, batch, want_ratio=None):
    ref, got = batch.alloc_out(), batch.alloc_out(pinned=True)
    try:
        oracle_lib.genotype_batch(batch, ref, threads=8)
        orc = 0
    except oracle_lib.OracleError as e:
        orc = e.code
    dev = ctx.upload_wire(batch)
    bad = ctx.compare_records(dev, batch)
    assert not bad, f"record-store arrays rebuilt from the wire form differ: {bad}"
    dev.launch()
    from glnexus_b200 import GlxError
    if orc != 0:
        with pytest.raises(GlxError) as ei:
            dev.check()
        assert ei.value.code == orc
        dev.close()
        return
    dev.check()
    dev.download(got)
    wire_bytes = batch.wire()[1]
    full_bytes = ctx.upload(batch)
    ratio = full_bytes.in_bytes / max(1, dev.in_bytes)
    full_bytes.close()
    dev.close()
    d = compare_outs(got, ref, batch.n_fields)
    assert not d, "\n".join(d)
    if want_ratio:
        assert ratio >= want_ratio, f"H2D bytes only {ratio:.2f}x smaller ({wire_bytes} wire bytes)"
    return ratio


@pytest.mark.parametrize("name", gu.case_names())
def test_wire_golden(ctx, name):
    for batch in gu.batches_of_case(gu.load_case(name)):
        if batch.has_fast_path:  # (multi-sample datasets have no wire form: they upload full-width)
            check(ctx, batch)


@pytest.mark.parametrize("shape,ns,nsites,region,ratio", [("c2", 1000, 3000, 3000, 1.8), ("c3", 1500, 2000, 128000, 1.8), ("c4", 2000, 1500, 96000, 1.8),
                                                          ("c5", 600, 2000, 60000, 1.5), ("c2", 3, 50, 50, None)])
def test_wire_synth(ctx, shape, ns, nsites, region, ratio):
    b = synth_cases.make(shape, n_samples=ns, n_sites=nsites, region_len=region)
    check(ctx, b, ratio)


@pytest.mark.parametrize("preset", ["xAtlas", "weCall", "Strelka2", "gatk_unfiltered"])
def test_wire_presets(ctx, preset):
    b = synth_cases.make("c3", n_samples=300, n_sites=800, region_len=24000, preset=preset, frac_two_record_cells=0.05, frac_lost_allele=0.05, frac_homalt=0.2)
    check(ctx, b)


def test_wire_fuzz(ctx):
    n = 0
    for seed in range(1000, 1120):
        b = fuzz_cases.build_batch(seed)
        check(ctx, b)
        n += 1
    assert n == 120


def test_wire_sliced_batches_and_stream(ctx):
    """Slices (glxh_batch_slice) have a wire form too; the streamed driver uploads through it."""
    import glnexus_b200
    whole = synth_cases.make("c3", n_samples=400, n_sites=2400, region_len=153600)
    ref = whole.alloc_out()
    oracle_lib.genotype_batch(whole, ref, threads=8)
    want, _ = oracle_lib.bcf_records(whole, ref)
    rs = glnexus_b200.RangeStream(0, "bcf", 3, capture=True, wire=True)
    parts = [whole.slice(a, min(whole.n_sites, a + 500)) for a in range(0, whole.n_sites, 500)]
    for p in parts:
        check(ctx, p)
        rs.submit(p)
    rs.drain()
    assert rs.output() == want
    assert rs.stats()["h2d_bytes"] < sum(p.input_bytes for p in parts) * 0.7
    rs.close()


def test_wire_empty(ctx):
    b = synth_cases.make("c2", n_samples=33, n_sites=0, region_len=500)
    dev = ctx.upload_wire(b)
    dev.launch()
    dev.check()
    dev.close()
    b = synth_cases.make("c2", n_samples=7, n_sites=40, region_len=40, frac_uncovered=1.0)  # sites but (almost) no records
    check(ctx, b)


@pytest.mark.parametrize("cap", [0, 2])
def test_wire_shape_exceptions(ctx, cap):
    """GLX_WF_SHAPES with a shape table too small for the cohort: most records take the exception path of glx_wire_shapes_kernel
    (code 0xFF, own row found by binary search). The expanded store must still equal the full-width one, outputs the oracle's."""
    import glnexus_b200
    glnexus_b200.wire_shape_cap(cap)
    try:
        check(ctx, synth_cases.make("c5", n_samples=300, n_sites=1500, region_len=45000))
        for seed in (7, 8, 9, 10):
            b = fuzz_cases.build_batch(seed)
            if b.has_fast_path:
                check(ctx, b)
    finally:
        glnexus_b200.wire_shape_cap(255)
