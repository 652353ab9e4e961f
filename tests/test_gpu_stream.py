"""The streamed range-batch driver (glx::RangeStream / glxh_stream_*) and range sharding on the GPU: a cohort cut into
contiguous site ranges with the product's glxh_partition_sites / glxh_batch_slice, each range streamed as several range
batches, outputs concatenated in order == the unsharded oracle (ordered drain of src/service.cc:388-418)."""
import pytest

import bcfdec
import oracle_lib
import synth_cases

pytestmark = pytest.mark.gpu


def n_devices():
    import torch
    return torch.cuda.device_count()


def cut(batch, lo, hi, batch_sites):
    """[lo, hi) as consecutive range batches of batch_sites sites (each with every record its sites can touch)."""
    out = []
    for a in range(lo, hi, batch_sites):
        out.append(batch.slice(a, min(hi, a + batch_sites)))
    return out


def run_ranks(whole, n_ranks, devices, batch_sites, mode, depth=3):
    import glnexus_b200
    bounds = whole.partition(n_ranks)
    assert bounds[0] == 0 and bounds[-1] == whole.n_sites and all(a <= b for a, b in zip(bounds, bounds[1:]))
    streams, parts = [], []
    for r in range(n_ranks):  # every rank queues its whole range before anyone drains: the devices run concurrently
        rs = glnexus_b200.RangeStream(devices[r % len(devices)], mode, depth, capture=True)
        batches = cut(whole, bounds[r], bounds[r + 1], batch_sites)
        for b in batches:
            rs.submit(b)
        streams.append((rs, batches))
    for rs, batches in streams:
        rs.drain()
        st = rs.stats()
        assert st["batches"] == st["batches_retired"] == len(batches)
        parts.append(rs.output())
        rs.close()
    return b"".join(parts), bounds


@pytest.mark.parametrize("shape,ns,nsites,region", [("c3", 700, 3000, 192000), ("c2", 500, 4000, 4000), ("c5", 300, 1500, 45000)])
@pytest.mark.parametrize("n_ranks", [1, 2, 4])
def test_stream_shards_concat_equals_unsharded_oracle(built, shape, ns, nsites, region, n_ranks):
    whole = synth_cases.make(shape, n_samples=ns, n_sites=nsites, region_len=region)
    ref = whole.alloc_out()
    oracle_lib.genotype_batch(whole, ref, threads=8)
    want, _ = oracle_lib.bcf_records(whole, ref)
    devices = list(range(max(1, min(n_devices(), n_ranks))))
    got, _ = run_ranks(whole, n_ranks, devices, batch_sites=max(64, nsites // (5 * n_ranks)), mode="bcf")
    assert got == want, "concatenated shard records differ from the unsharded oracle"
    blob, _ = run_ranks(whole, n_ranks, devices, batch_sites=max(64, nsites // (5 * n_ranks)), mode="bgzf", depth=2)
    inflated, _ = bcfdec.bgzf_decompress(blob)
    assert inflated == want


def test_stream_two_devices(built):
    """Two ranks on two different GPUs (skipped on a one-GPU box; the same code path as above with devices [0, 1])."""
    if n_devices() < 2:
        pytest.skip("needs two devices")
    whole = synth_cases.make("c3", n_samples=2000, n_sites=6000, region_len=384000)
    ref = whole.alloc_out()
    oracle_lib.genotype_batch(whole, ref, threads=8)
    want, _ = oracle_lib.bcf_records(whole, ref)
    got, bounds = run_ranks(whole, 2, [0, 1], batch_sites=512, mode="bcf")
    assert (bounds[2] - bounds[1]) >= 4 * 512 and got == want  # >= 4 batches per rank


def test_stream_error_stops_output(built):
    """A failing batch in the middle: the error surfaces, batches after it produce no output (no rows after a gap)."""
    import glnexus_b200
    import fuzz_cases
    good = synth_cases.make("c2", n_samples=50, n_sites=200, region_len=200)
    bad = None
    for seed in range(1000, 1240):
        b = fuzz_cases.build_batch(seed)
        try:
            oracle_lib.genotype_batch(b, b.alloc_out())
        except oracle_lib.OracleError:
            bad = b
            break
    assert bad is not None
    ref = good.alloc_out()
    oracle_lib.genotype_batch(good, ref)
    one, _ = oracle_lib.bcf_records(good, ref)
    rs = glnexus_b200.RangeStream(0, "bcf", 3, capture=True)
    with pytest.raises(glnexus_b200.GlxError):
        for b in (good, good, bad, good, good, good, good):
            rs.submit(b)
        rs.drain()
    assert rs.output() == one * 2
    rs.close()


def test_stream_columns_mode(built):
    import glnexus_b200
    b = synth_cases.make("c2", n_samples=300, n_sites=900, region_len=900)
    rs = glnexus_b200.RangeStream(0, "columns", 2)
    for _ in range(5):
        rs.submit(b)
    rs.drain()
    st = rs.stats()
    assert st["batches_retired"] == 5 and st["cells"] == 5 * 300 * 900 and st["d2h_bytes"] > 0 and st["h2d_bytes"] > 0
    rs.close()
