"""pack_records packs datasets on several threads (independent per dataset; the value / GT / allele pools laid end to end afterwards).
The record store must not depend on the thread count: the same text cohort packed with 1 and with several threads gives the same wire
form byte for byte (every column, pool and offset of the store goes into it) and the same oracle outputs."""
import ctypes as C

import numpy as np
import pytest

import fuzz_cases
import oracle_lib


def _cohort(multi):
    """Several fuzz cohorts' gVCFs side by side over one site list: ~50 datasets."""
    sites_text, datasets, preset, toggles = fuzz_cases.make_case(4242, "DeepVariant", multi)
    all_ds = []
    for k in range(12):
        _, ds, _, _ = fuzz_cases.make_case(4242 + 1000 * k, "DeepVariant", multi)
        # (other seeds draw other sites; their records still join by position against this case's sites)
        for name, body in ds:
            if multi:
                hdr_line = [l for l in body.split("\n") if l.startswith("#CHROM")][0]
                cols = hdr_line.split("\t")
                new = cols[:9] + [f"k{k}_{c}" for c in cols[9:]]
                body = body.replace(hdr_line, "\t".join(new))
                all_ds.append((f"k{k}_{name}", body))
            else:
                all_ds.append((f"k{k}_{name}", body.replace("\tFORMAT\t" + name, "\tFORMAT\tk%d_%s" % (k, name))))
    return sites_text, all_ds


def _image(b):
    ptr, n = b.wire()
    return C.string_at(ptr, n)


@pytest.mark.parametrize("multi", [False, True])
def test_pack_threads_do_not_change_the_store(built, multi):
    import glnexus_b200
    from glnexus_b200 import Batch
    sites_text, ds = _cohort(multi)
    outs, images = [], []
    try:
        for threads in (1, 4, 7):
            glnexus_b200.pack_threads(threads)
            b = Batch.from_text(sites_text, ds, preset="DeepVariant")
            assert b.n_samples >= 20 and b.n_records > 200
            ref = b.alloc_out()
            try:
                oracle_lib.genotype_batch(b, ref)
                rc = 0
            except oracle_lib.OracleError as e:
                rc = e.code
            outs.append((rc, {k: v.copy() for k, v in ref.arrays().items()}))
            if b.has_fast_path:
                images.append(_image(b))
    finally:
        glnexus_b200.pack_threads(0)
    for rc, arrs in outs[1:]:
        assert rc == outs[0][0]
        if rc == 0:
            for k in arrs:
                assert np.array_equal(arrs[k], outs[0][1][k]), k
    for im in images[1:]:
        assert im == images[0]
    assert multi or len(images) == 3
