"""The wire form's record-shape coding (GLX_WF_SHAPES, include/glx.h) on the host: every record's one-byte code must decode —
table row or its own exception row — to the flags, allele count, GT pair and column lengths of the record store it was built from.
(The device-side expansion of the whole wire form is compared array by array in test_gpu_wire.py.)"""
import pytest

import emu_lib
import fuzz_cases
import golden_util as gu
import synth_cases


@pytest.mark.parametrize("shape,ns,nsites,region", [("c2", 300, 3000, 3000), ("c3", 200, 1500, 96000), ("c5", 200, 1500, 45000)])
def test_wire_shapes_synth(built, shape, ns, nsites, region):
    b = synth_cases.make(shape, n_samples=ns, n_sites=nsites, region_len=region)
    bad, n_shapes, n_exc = emu_lib.wire_shapes_check(b)
    assert bad == 0 and 0 < n_shapes <= 255
    assert n_exc * 100 <= b.n_records  # a cohort's records are a few dozen shapes: next to none may need its own row
    _, n_bytes = b.wire()
    print(f"{shape}: {n_shapes} shapes, {n_exc} exceptions, {n_bytes / b.n_records:.1f} wire bytes per record")


def test_wire_shapes_fuzz_and_golden(built):
    n = 0
    for seed in range(300, 360):
        b = fuzz_cases.build_batch(seed)
        bad, _, _ = emu_lib.wire_shapes_check(b)
        assert bad in (0, -1), f"seed {seed}: {bad} records decode to another shape"
        n += bad == 0
    for name in gu.case_names():
        for b in gu.batches_of_case(gu.load_case(name)):
            if b.has_fast_path and b.n_records:
                bad, _, _ = emu_lib.wire_shapes_check(b)
                assert bad in (0, -1), f"{name}: {bad}"
                n += bad == 0
    assert n >= 60


@pytest.mark.parametrize("cap", [0, 1, 3])
def test_wire_shapes_exception_list(built, cap):
    """A cohort with more shapes than the table holds: the records of the rarer shapes carry code 0xFF and find their own row by
    binary search in the sorted exception list (forced here by capping the table at `cap` rows)."""
    import glnexus_b200
    glnexus_b200.wire_shape_cap(cap)
    try:
        b = synth_cases.make("c5", n_samples=60, n_sites=800, region_len=24000)
        bad, n_shapes, n_exc = emu_lib.wire_shapes_check(b)
        assert bad == 0 and n_shapes == cap and n_exc > 0
        for seed in range(400, 420):
            fb = fuzz_cases.build_batch(seed)
            bad, _, _ = emu_lib.wire_shapes_check(fb)
            assert bad in (0, -1)
    finally:
        glnexus_b200.wire_shape_cap(255)


def test_wire_bytes_are_reproducible(built):
    """The wire buffer is allocated uninitialised and filled by several threads: every byte — section padding included — must be
    written. Two builds of the same cohort (fresh allocations, other buffers freed in between) must be identical, and rebuilding with
    a poisoned allocator state must not change them."""
    import ctypes as C

    def image(seed_fill):
        junk = bytearray([seed_fill]) * (64 << 20)  # dirty the heap the next allocation is likely to reuse
        del junk
        b = synth_cases.make("c2", n_samples=400, n_sites=4000, region_len=4000)  # > 65,536 records: the multi-threaded path
        ptr, n = b.wire()
        return C.string_at(ptr, n)

    a, b2 = image(0xAA), image(0x55)
    assert a == b2
