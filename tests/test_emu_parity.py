"""Device cell logic (compiled for the host, tests/emu) vs the oracle on the reference's golden fixtures: every
output column bit-exact. CPU only; the same comparison runs against the real kernels in test_gpu_parity.py."""
import numpy as np
import pytest

import emu_lib
import golden_util as gu
import oracle_lib


def compare_outs(a, b, n_fields):
    diffs = []
    aa, bb = a.arrays(), b.arrays()
    for k in aa:
        if not np.array_equal(aa[k], bb[k]):
            idx = np.nonzero(aa[k] != bb[k])[0]
            diffs.append(f"{k}: {len(idx)} differ, first at {idx[:5]}: {aa[k][idx[:5]]} vs {bb[k][idx[:5]]}")
    return diffs


@pytest.mark.parametrize("name", gu.case_names())
def test_emu_general_matches_oracle(built, name):
    case = gu.load_case(name)
    for batch in gu.batches_of_case(case):
        ref, got = batch.alloc_out(), batch.alloc_out()
        try:
            oracle_lib.genotype_batch(batch, ref)
            orc = 0
        except oracle_lib.OracleError as e:
            orc = e.code
        rc = emu_lib.general_batch(batch, got)
        got2 = batch.alloc_out()
        rc2, _ = emu_lib.tiled_batch(batch, got2)
        if not batch.device_supported:
            assert rc == -101 and rc2 == -101
            continue
        assert rc == orc and rc2 == orc
        if rc == 0:
            d = compare_outs(got, ref, batch.n_fields)
            assert not d, "general:\n" + "\n".join(d)
            d = compare_outs(got2, ref, batch.n_fields)
            assert not d, "tiled:\n" + "\n".join(d)
