"""Randomised irregular cohorts (tests/fuzz_cases.py): device cell logic (host build) vs the oracle on CPU, the CUDA
path vs the oracle on the GPU. Status codes must agree; on success every output column must be bit-exact."""
import numpy as np
import pytest

import emu_lib
import fuzz_cases
import oracle_lib
from test_emu_parity import compare_outs


def _oracle(b):
    ref = b.alloc_out()
    try:
        oracle_lib.genotype_batch(b, ref)
        return 0, ref
    except oracle_lib.OracleError as e:
        return e.code, ref


@pytest.mark.parametrize("chunk", range(8))
def test_fuzz_emu(built, chunk):
    n_ok = 0
    for seed in range(chunk * 60, chunk * 60 + 60):
        b = fuzz_cases.build_batch(seed)
        orc, ref = _oracle(b)
        for name, run in (("general", lambda o: emu_lib.general_batch(b, o)), ("tiled", lambda o: emu_lib.tiled_batch(b, o)[0]),
                          ("tiled-dynamic", lambda o: emu_lib.tiled_batch(b, o, True)[0])):
            got = b.alloc_out()
            rc = run(got)
            assert rc == orc, f"seed {seed} {name}: status {rc} vs oracle {orc}"
            if orc == 0:
                d = compare_outs(got, ref, b.n_fields)
                assert not d, f"seed {seed} {name}:\n" + "\n".join(d)
        n_ok += orc == 0
    assert n_ok >= 24  # enough cases must be well-formed enough to genotype (the rest compare error statuses)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["sig", "sig-unfused", "dynamic", "general"])
def test_fuzz_gpu(built, path):
    from glnexus_b200 import Context, GlxError
    ctx = Context(0)
    ctx.select_path(path)
    n_ok = 0
    for seed in range(1000, 1240):
        b = fuzz_cases.build_batch(seed)
        orc, ref = _oracle(b)
        got = b.alloc_out(pinned=False)
        try:
            ctx.genotype_batch(b, got)
            rc = 0
        except GlxError as e:
            rc = e.code
        assert rc == orc, f"seed {seed}: status {rc} vs oracle {orc}"
        if orc == 0:
            d = compare_outs(got, ref, b.n_fields)
            assert not d, f"seed {seed}:\n" + "\n".join(d)
            n_ok += 1
    ctx.close()
    assert n_ok >= 100


# ---- multi-sample datasets (one to three samples per gVCF): glx_cells_multi_kernel's logic vs the oracle ----
def _multi_cases(seeds):
    n_multi = 0
    for seed in seeds:
        b = fuzz_cases.build_batch(seed, multi=True)
        assert b.device_supported
        n_multi += not b.has_fast_path
        yield seed, b
    assert n_multi >= len(seeds) // 2  # most cases must really hold a multi-sample dataset


@pytest.mark.parametrize("chunk", range(6))
def test_fuzz_multi_sample_emu(built, chunk):
    n_ok = 0
    for seed, b in _multi_cases(range(5000 + chunk * 60, 5000 + chunk * 60 + 60)):
        orc, ref = _oracle(b)
        got = b.alloc_out()
        rc = emu_lib.general_batch(b, got)
        assert rc == orc, f"seed {seed}: status {rc} vs oracle {orc}"
        if orc == 0:
            d = compare_outs(got, ref, b.n_fields)
            assert not d, f"seed {seed}:\n" + "\n".join(d)
            n_ok += 1
    assert n_ok >= 20


@pytest.mark.gpu
def test_fuzz_multi_sample_gpu(built):
    from glnexus_b200 import Context, GlxError
    ctx = Context(0)
    n_ok = 0
    for seed, b in _multi_cases(range(6000, 6200)):
        orc, ref = _oracle(b)
        got = b.alloc_out(pinned=False)
        try:
            ctx.genotype_batch(b, got)
            rc = 0
        except GlxError as e:
            rc = e.code
        assert rc == orc, f"seed {seed}: status {rc} vs oracle {orc}"
        if orc == 0:
            d = compare_outs(got, ref, b.n_fields)
            assert not d, f"seed {seed}:\n" + "\n".join(d)
            n_ok += 1
    ctx.close()
    assert n_ok >= 60


def test_multi_sample_slices_concatenate(built):
    """Range sharding of a cohort with multi-sample gVCFs (glxh_batch_slice carries the datasets' GT and value pools): the two halves'
    columns, computed separately, concatenate to the whole batch's (site-major arrays)."""
    n = 0
    for seed in range(7000, 7060):
        b = fuzz_cases.build_batch(seed, multi=True)
        if b.has_fast_path or b.n_sites < 3:
            continue
        orc, ref = _oracle(b)
        if orc != 0:
            continue
        want = {k: v.copy() for k, v in ref.arrays().items()}
        cut = b.n_sites // 2
        got = {k: [] for k in want}
        for part in (b.slice(0, cut), b.slice(cut, b.n_sites)):
            o = part.alloc_out()
            assert emu_lib.general_batch(part, o) == 0
            for k, v in o.arrays().items():
                got[k].append(v.copy())
        for k in want:
            assert np.array_equal(np.concatenate(got[k]), want[k]), f"seed {seed}: column {k} of the halves differs from the whole"
        n += 1
    assert n >= 30
