"""Multi-GPU host logic on CPU: contiguous site-range sharding with no data-path collective. Two gloo ranks each
take one shard of the same seeded cohort, compute it (device cell logic, host build — there is no GPU here), and
rank 0 concatenates the shards' columns in order; the result must equal the unsharded oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import synth_cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import emu_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = synth_cases.make("c3", n_samples=70, n_sites=501, region_len=30000, frac_two_record_cells=0.02)
    bounds = b.partition(world)
    shard = b.slice(bounds[rank], bounds[rank + 1])
    out = shard.alloc_out()
    rc, _ = emu_lib.tiled_batch(shard, out)
    cols = {k: v.copy() for k, v in out.arrays().items()}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((rc, bounds, shard.n_records, cols), gathered, dst=0)
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_range_sharding(built):
    import oracle_lib
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port() if r == 0 else None, q)) for r in range(world)]
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = synth_cases.make("c3", n_samples=70, n_sites=501, region_len=30000, frac_two_record_cells=0.02)
    ref = b.alloc_out()
    oracle_lib.genotype_batch(b, ref)
    bounds = gathered[0][1]
    assert bounds[0] == 0 and bounds[-1] == b.n_sites and 0 < bounds[1] < b.n_sites
    assert all(g[0] == 0 for g in gathered)
    assert sum(g[2] for g in gathered) >= b.n_records  # records straddling the cut are replicated, none lost
    for k, full in ref.arrays().items():
        cat = np.concatenate([g[3][k] for g in gathered])
        assert np.array_equal(cat, full), k


def test_slice_edges(built):
    import emu_lib
    import oracle_lib
    b = synth_cases.make("c5", n_samples=40, n_sites=300, region_len=12000)
    ref = b.alloc_out()
    oracle_lib.genotype_batch(b, ref)
    full = ref.arrays()
    for lo, hi in [(0, 0), (0, 1), (299, 300), (17, 230), (0, 300)]:
        sh = b.slice(lo, hi)
        assert sh.n_sites == hi - lo
        out = sh.alloc_out()
        assert emu_lib.general_batch(sh, out) == 0
        N = b.n_samples
        assert np.array_equal(out.array("gt"), full["gt"][lo * N * 2:hi * N * 2])
        assert np.array_equal(out.array("rnc"), full["rnc"][lo * N * 2:hi * N * 2])
        assert np.array_equal(out.array("allele_used"), full["allele_used"][lo:hi])
    assert b.partition(8)[0] == 0 and b.partition(8)[-1] == 300 and b.partition(8) == sorted(b.partition(8))
