"""Pin the oracle on the reference's unit-level known-answer tests (CPU only):
  * diploid index math            test/diploid.cc:9-39   (exhaustive for n_allele < 16)
  * PL -> log-likelihood          test/diploid.cc:41-94  (incl. '.' -> log(0))
  * revise_genotypes              test/genotyper.cc:110-172 (14 cases: expected allele pair + GQ)
and run the same revise_genotypes cases through the device cell logic (host build) via a one-site batch."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import emu_lib
import golden_util as gu
import oracle_lib

KATS = json.load(open(os.path.join(gu.GOLDEN, "_revise_genotypes_kats.json")))


def test_diploid_genotypes(built):
    L = oracle_lib.lib()
    assert [L.glx_oracle_genotypes(n) for n in (1, 2, 3)] == [1, 3, 6]


def test_diploid_gt_alleles_and_alleles_gt(built):
    L = oracle_lib.lib()
    for n_allele in range(16):
        x = 0
        for j in range(n_allele):
            for i in range(j + 1):
                a, b = C.c_uint(), C.c_uint()
                L.glx_oracle_gt_alleles(x, C.byref(a), C.byref(b))
                assert (a.value, b.value) == (i, j)
                assert L.glx_oracle_alleles_gt(i, j) == x
                assert L.glx_oracle_alleles_gt(j, i) == x
                x += 1


def test_pl_to_log_likelihood(built):
    L = oracle_lib.lib()
    phred2log = 1.0 / (-10.0 * math.log10(math.exp(1.0)))
    for pl in (16, 240, 46, 246, 292):
        assert abs(L.glx_oracle_pl_to_gll(pl) - pl * phred2log) < 1e-9
    assert L.glx_oracle_pl_to_gll(0) == 0.0
    assert L.glx_oracle_pl_to_gll(-2 ** 31) == -math.inf      # '.' (bcf_int32_missing) -> log(0)
    assert L.glx_oracle_pl_to_gll(-2 ** 31 + 1) == -math.inf  # vector_end


def _kat_batch(case):
    from glnexus_b200 import Batch
    return Batch.from_text(case["sites_text"], [("A", case["vcf"])], config_text=KATS["config_text"])


@pytest.mark.parametrize("i", range(len(KATS["cases"])))
def test_revise_genotypes_kat_oracle(built, i):
    case = KATS["cases"][i]
    b = _kat_batch(case)
    assert oracle_lib.revise_kat(b) == (case["a1"], case["a2"], case["gq"])


@pytest.mark.parametrize("i", range(len(KATS["cases"])))
def test_revise_genotypes_kat_device_logic(built, i):
    """The same records through genotype_site (oracle) and the device cell logic: GT and lifted GQ must agree, and
    where the revised call is representable at the site they must equal the KAT."""
    case = KATS["cases"][i]
    b = _kat_batch(case)
    ref, got = b.alloc_out(), b.alloc_out()
    oracle_lib.genotype_batch(b, ref)
    assert emu_lib.general_batch(b, got) == 0
    for k, a in ref.arrays().items():
        assert np.array_equal(a, got.arrays()[k]), k
    got2 = b.alloc_out()
    assert emu_lib.tiled_batch(b, got2)[0] == 0
    for k, a in ref.arrays().items():
        assert np.array_equal(a, got2.arrays()[k]), k
    gq = int(got.array(0)[0])  # the KAT config lifts GQ only
    if (ref.array("rnc").tobytes() == b".."):
        gt = sorted(int(x) for x in got.array("gt")[:2])
        assert gt == sorted([case["a1"], case["a2"]])
        assert gq == case["gq"]
