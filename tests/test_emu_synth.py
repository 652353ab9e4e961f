"""Device cell logic (host build, tests/emu) vs the oracle on seeded synthetic cohorts of the BASELINE shapes."""
import numpy as np
import pytest

import emu_lib
import oracle_lib
import synth_cases
from test_emu_parity import compare_outs


@pytest.mark.parametrize("shape", ["c2", "c3", "c4", "c5"])
def test_emu_general_synth(built, shape):
    b = synth_cases.make(shape, n_samples=96, n_sites=700, region_len=6000 if shape == "c2" else 40000)
    ref, got = b.alloc_out(), b.alloc_out()
    oracle_lib.genotype_batch(b, ref, threads=4)
    assert emu_lib.general_batch(b, got) == 0
    d = compare_outs(got, ref, b.n_fields)
    assert not d, "\n".join(d)
    for dyn in (False, True):
        got2 = b.alloc_out()
        rc, n_slow = emu_lib.tiled_batch(b, got2, dynamic_only=dyn)
        assert rc == 0
        d = compare_outs(got2, ref, b.n_fields)
        assert not d, f"tiled (dynamic_only={dyn}):\n" + "\n".join(d)
    assert n_slow < 0.5 * b.n_sites * b.n_samples  # the fast path takes the bulk of the cells
    rnc = set(ref.array("rnc").tobytes().decode())
    assert {".", "M"} <= rnc


@pytest.mark.parametrize("preset", ["gatk_unfiltered", "DeepVariant_unfiltered", "DeepVariantWES", "weCall", "Strelka2"])
def test_emu_general_presets(built, preset):
    b = synth_cases.make("c3", n_samples=40, n_sites=300, region_len=9000, preset=preset, frac_two_record_cells=0.05, frac_lost_allele=0.05,
                         frac_homalt=0.2)
    ref, got = b.alloc_out(), b.alloc_out()
    try:
        oracle_lib.genotype_batch(b, ref)
        orc = 0
    except oracle_lib.OracleError as e:
        orc = e.code
    assert emu_lib.general_batch(b, got) == orc
    got2, got3 = b.alloc_out(), b.alloc_out()
    assert emu_lib.tiled_batch(b, got2)[0] == orc
    assert emu_lib.tiled_batch(b, got3, dynamic_only=True)[0] == orc
    if orc == 0:
        d = compare_outs(got, ref, b.n_fields)
        assert not d, "\n".join(d)
        d = compare_outs(got2, ref, b.n_fields)
        assert not d, "tiled:\n" + "\n".join(d)
        d = compare_outs(got3, ref, b.n_fields)
        assert not d, "tiled dynamic:\n" + "\n".join(d)


def _real_slice_batch():
    import json
    import os
    import golden_util as gu
    from glnexus_b200 import Batch
    fx = json.load(open(os.path.join(gu.GOLDEN, "_real_na12878_slice.json")))
    return Batch.from_text(fx["sites_text"], [(d["name"], d["vcf"]) for d in fx["datasets"]], preset=fx["preset"])


def test_emu_real_gvcf_slice(built):
    """Real GATK gVCF records (a slice of the reference's test/data/NA12878.g.vcf.gz) as a 3-sample cohort."""
    b = _real_slice_batch()
    assert b.n_samples == 3 and b.n_sites > 20 and b.n_records > 4000
    ref = b.alloc_out()
    oracle_lib.genotype_batch(b, ref)
    for run in (lambda o: emu_lib.general_batch(b, o), lambda o: emu_lib.tiled_batch(b, o)[0], lambda o: emu_lib.tiled_batch(b, o, True)[0]):
        got = b.alloc_out()
        assert run(got) == 0
        d = compare_outs(got, ref, b.n_fields)
        assert not d, "\n".join(d)
    rnc = ref.array("rnc").tobytes().decode()
    assert ".." in rnc


@pytest.mark.parametrize("tile_sites,block", [(64, 256), (32, 256), (32, 64), (16, 32)])
def test_resolve_hint_arithmetic(built, tile_sites, block):
    """The resolve kernels' confined searches (per-block hints, neighbour's t_hi as t_lo) must claim exactly the tiles
    fill_tile_cursors defines, on the synthetic shapes and on irregular cohorts (several samples per block, empty samples)."""
    import fuzz_cases
    for shape, (ns, nsite, rl) in {"c2": (96, 700, 700), "c3": (300, 512, 32768), "c4": (2000, 128, 8192), "c5": (64, 300, 9000)}.items():
        b = synth_cases.make(shape, n_samples=ns, n_sites=nsite, region_len=rl)
        assert emu_lib.check_hints(b, tile_sites, block) == 0, shape
    for seed in range(120):
        assert emu_lib.check_hints(fuzz_cases.build_batch(seed), tile_sites, block) == 0, seed

