#!/usr/bin/env python3
"""bench.py — site x sample genotype cells/s of the joint-genotyping hot path on N B200s (one process per GPU).

A "step" is one pass of the hot path (glx_launch: every kernel of the path) over one contig-range batch of a
seeded synthetic DeepVariant-style cohort (BASELINE.json configs[1]: 1,000 samples x 50,000 dense exome-like
sites, 5e7 cells). The cohort is 6 x N such segments; it is cut into N contiguous site ranges of equal output weight by
the product's own partitioner (glxh_partition_cohort / glxh_batch_slice) and rank r takes range r (weak scaling; sites
shard with no collective). `value` is the cells all ranks processed over the slowest rank's device time.
`--segments G` fixes the cohort at G segments whatever N is (strong scaling).

  value     device-resident throughput: batch uploaded once, K launches timed with CUDA events on the launching
            stream (L2 flushed between steps, flush not timed)
  e2e       same metric through the reference-facing C ABI with HOST buffers: the C++ range-batch driver (glxh_stream_*)
            streams the rank's range batches — H2D of the record store + sites, all kernels, the on-device BCF2/BGZF
            encoder, D2H of the finished BGZF blocks — all inside the timed region
  roofline  algorithmic bytes (SURVEY.md §8d definition) / tiled-kernel duration vs the measured HBM copy peak
  cpu_baseline  the C++ oracle (port of the reference algorithm) on the box's host cores over a bounded sample

`--impl reference` times the reference algorithm's CPU port (oracle/, all host threads) on the same config.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "site x sample genotype cells/s"
WORKLOADS = {
    # name: (synthetic shape, samples, sites, region_len)
    "c2": ("c2", 1000, 50000, 50000),
    "c3": ("c3", 10000, 4096, 262144),   # one streamed batch of configs[2]
    "c4": ("c4", 100000, 512, 32768),    # one streamed batch of configs[3]
    "c5": ("c5", 5000, 10000, 300000),   # a tenth of configs[4]
}


def algorithmic_bytes(batch, n_var_cells):
    """int32-equivalent bytes the reference's inner loops read/write per cell.
    out/cell = 2 (GT) + 2 (RNC) + 4 per lifted value; in/cell = 34 for a reference-band cell,
    22 + 5a + 2a(a+1) for a variant cell with an a-allele record (a=3 for REF,ALT,<*>)."""
    import glnexus_b200
    out_b = glnexus_b200.out_bytes_of(batch)
    cells = batch.n_sites * batch.n_samples
    in_b = 34 * (cells - n_var_cells) + 61 * n_var_cells
    return out_b + in_b, out_b, in_b


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs local to its GPU's PCIe root complex (sysfs local_cpulist of the device), so that the
    pinned host buffers it allocates are first-touched on that NUMA node. Returns the node, or None if unknown."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        base = "/sys/bus/pci/devices/" + bus
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        node = int(open(base + "/numa_node").read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                clk = float(f[0])
                mx = float(f[1])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to every sample taken
            sm = [float(l.split(",")[0]) for _, l in self.lines if l and l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(shape, n_samples, region_per_site, max_sites, target_s=12.0):
    """Oracle (kind 'port') on all host threads over the first sites of the same workload (at most the workload)."""
    import oracle_lib
    import synth_cases
    cores = host_cores()
    sites = min(64, max_sites)
    while True:
        b = synth_cases.make(shape, n_samples=n_samples, n_sites=sites, region_len=max(sites, int(sites * region_per_site)))
        out = b.alloc_out()
        t = time.time()
        oracle_lib.genotype_batch(b, out, threads=cores)
        dt = time.time() - t
        cells = b.n_sites * b.n_samples
        if dt >= target_s * 0.6 or sites >= max_sites:
            break
        sites = int(min(max(sites * 2, sites * target_s / max(dt, 1e-3) * 0.9), sites * 16, max_sites))
    reps = 1
    while dt < target_s * 0.6 and reps < 64:  # whole workload is shorter than the target: repeat it
        t = time.time()
        oracle_lib.genotype_batch(b, out, threads=cores)
        dt += time.time() - t
        reps += 1
    rate = cells * reps / dt
    return {"value": rate, "unit": "cells/s", "cores": cores, "kind": "port",
            "sample": f"first {b.n_sites} sites x {n_samples} samples of the same synthetic cohort ({cells} cells) x {reps} passes in {dt:.1f} s, "
                      f"one task per site on {cores} host threads, reference-faithful data structures"}, b


def run_reference(args, rank, world):
    if rank != 0:
        return
    shape, n_samples, n_sites, region_len = WORKLOADS[args.workload]
    if args.samples:
        n_samples = args.samples
    import oracle_lib
    import synth_cases
    cores = host_cores()
    # bounded sample per step: sized from a probe so that (steps+warmup) steps end within a few minutes
    probe, b = cpu_baseline(shape, n_samples, region_len / n_sites, n_sites, target_s=2.0)
    budget = 150.0 / max(1, args.steps + args.warmup)
    sites = min(n_sites, max(16, int(probe["value"] * min(budget, 15.0) / n_samples)))
    b = synth_cases.make(shape, n_samples=n_samples, n_sites=sites, region_len=max(sites, int(sites * region_len / n_sites)))
    out = b.alloc_out()
    cells = b.n_sites * b.n_samples
    for _ in range(args.warmup):
        oracle_lib.genotype_batch(b, out, threads=cores)
    t = time.time()
    for _ in range(args.steps):
        oracle_lib.genotype_batch(b, out, threads=cores)
    dt = time.time() - t
    v = cells * args.steps / dt
    sample = f"{b.n_sites} sites x {n_samples} samples per step ({cells} cells), oracle port of genotype_site, one task per site on {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64",
        "data": "synthetic", "config": {"workload": workload_name(args.workload, n_samples, n_sites), "preset": synth_cases.SHAPES[shape][0]},
        "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_name(w, n_samples, n_sites):
    desc = {"c2": "BASELINE configs[1]: synthetic DeepVariant-style cohort, exome-like dense sites, biallelic-dominant",
            "c3": "BASELINE configs[2] (one streamed 4096-site batch): synthetic WGS cohort with reference bands and MIN_DP",
            "c4": "BASELINE configs[3] (one streamed batch): large sparse cohort, revise_genotypes AF priors on",
            "c5": "BASELINE configs[4] (subset): multi-allelic stress, --config gatk, up to 6 ALT"}[w]
    return f"{desc}; {n_samples} samples x {n_sites} sites per GPU"


def build_rank_batches(shape, n_samples, n_sites, region_len, n_segments, world, rank, pin=True):
    """The cohort is `n_segments` consecutive genomic segments of the workload's shape (segment g: its own region and seed).
    It is cut into `world` contiguous site ranges of equal output weight with the product's glxh_partition_cohort; this rank
    generates the segments its range touches and slices the two boundary segments with glxh_batch_slice. Returns the rank's
    range batches in genomic order."""
    import glnexus_b200
    import synth_cases
    seed0 = synth_cases.SHAPES[shape][1]["seed"]

    def seg(g, sites_only):
        return synth_cases.make(shape, n_samples=n_samples, n_sites=n_sites, region_len=region_len, region_beg=10_000_000 + g * region_len,
                                seed=seed0 + 7919 * g, sites_only=1 if sites_only else 0)

    skeleton = [seg(g, True) for g in range(n_segments)]
    cuts = glnexus_b200.partition_cohort(skeleton, world)
    (g_lo, s_lo), (g_hi, s_hi) = cuts[rank], cuts[rank + 1]
    batches = []
    for g in range(g_lo, min(n_segments, g_hi + 1)):
        a = s_lo if g == g_lo else 0
        b = s_hi if g == g_hi else skeleton[g].n_sites
        if b <= a:
            continue
        full = seg(g, False)
        part = full if (a == 0 and b == full.n_sites) else full.slice(a, b)
        if part is not full:
            full.close()
        if pin:
            part.pin()
        batches.append(part)
    for sk in skeleton:
        sk.close()
    return batches, cuts


def stream_pass(glnexus_b200, device, mode, batches, passes, depth=3, wire=False):
    """Wall seconds to stream `batches` x `passes` through the C++ range-batch driver (glxh_stream_*): per batch the H2D of the
    record store + sites, all kernels, the on-device encoder and the D2H of its output. Returns (seconds, stats)."""
    rs = glnexus_b200.RangeStream(device, mode, depth, wire=wire)
    for _ in range(depth):  # untimed passes: the arena cache and EVERY slot's pinned result buffer have then seen every batch size
        for b in batches:   # (slots rotate, so one pass is not enough when the batch count is not a multiple of the depth)
            rs.submit(b)
    rs.drain()
    st0 = rs.stats()
    debug = bool(os.environ.get("GLX_BENCH_DEBUG"))
    if debug:
        rs.trace(True)
    t = time.time()
    marks = []
    for _ in range(passes):
        for b in batches:
            rs.submit(b)
            marks.append(time.time() - t)
    rs.drain()
    dt = time.time() - t
    if debug:
        print(f"[rank {os.environ.get('RANK', '0')}] {mode} wire={wire}: sites per batch {[b.n_sites for b in batches]}, submit returns at (ms) "
              f"{[round(m * 1e3, 1) for m in marks]}, drained at {dt * 1e3:.1f} ms", file=sys.stderr)
        for i, r in enumerate(rs.trace_rows()):
            print(f"[rank {os.environ.get('RANK', '0')}]   batch {i}: device upload {r[0]:.2f}-{r[1]:.2f}, kernels done {r[2]:.2f}, encoder done {r[3]:.2f}, download done {r[4]:.2f}"
                  f" | host submit {r[5]:.2f}-{r[6]:.2f}, download queued {r[7]:.2f}", file=sys.stderr)
    st = {k: v - st0[k] for k, v in rs.stats().items()}
    rs.close()
    return dt, st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--sites", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--segments", type=int, default=0, help="segments (range batches) of the streamed cohort; default 6 per GPU (weak scaling)")
    ap.add_argument("--passes", type=int, default=4, help="times the rank's batches are streamed in the timed end-to-end region (4 x 6 range batches per GPU: the "
                    "pipeline's fill and drain, ~1 batch time, are then ~4 % of the run instead of ~15 %; a chromosome is hundreds of batches)")
    ap.add_argument("--e2e-modes", default="bgzf+wire,bgzf,bcf,columns")
    ap.add_argument("--depth", type=int, default=3, help="range batches in flight in the streamed end-to-end driver")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # Libraries (NCCL prints its version banner) write to fd 1; keep stdout clean for the one JSON line.
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import glnexus_b200
    import synth_cases

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)  # pinned host buffers land next to the GPU's PCIe root (first touch)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    shape, n_samples, n_sites, region_len = WORKLOADS[args.workload]
    if args.samples:
        n_samples = args.samples
    if args.sites:
        region_len = int(region_len * args.sites / n_sites)
        n_sites = args.sites
    strong = args.segments > 0
    n_segments = args.segments if strong else 6 * world
    t_gen = time.time()
    batches, cuts = build_rank_batches(shape, n_samples, n_sites, region_len, n_segments, world, rank)
    gen_s = time.time() - t_gen
    batch = max(batches, key=lambda b: b.n_sites)  # the device-resident measurement runs on the rank's largest range batch
    cells = batch.n_sites * batch.n_samples
    rank_cells = sum(b.n_sites * b.n_samples for b in batches)

    ctx = glnexus_b200.Context(local_rank)
    ctx.set_option("profile_upload", 1)
    dev = ctx.upload(batch)
    up_ms = dev.upload_kernel_ms()
    dev.set_profile(True)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            dev.launch(stream.cuda_stream)
        dev.check()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.3)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        kms = []
        t0 = time.time()
        barrier()
        for a, b in ev:
            flush.zero_()
            a.record(stream)
            dev.launch(stream.cuda_stream)
            b.record(stream)
            kms.append(dev.kernel_ms())  # waits for this step's kernels
        barrier()
        t1 = time.time()
        clocks = sampler.stop(t0, t1)
        # the on-device BCF2 encoder (N1), timed the same way: raw records, and records + BGZF (DEFLATE + CRC-32)
        dev.bcf_attach(batch, stream.cuda_stream)
        enc_ms = {}
        for mode, name in ((0, "bcf_records_ms"), (1, "bcf_bgzf_ms")):
            dev.encode_bcf(mode, stream.cuda_stream)
            torch.cuda.synchronize()
            evs = []
            for _ in range(max(3, args.steps // 2)):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                dev.encode_bcf(mode, stream.cuda_stream)
                b.record(stream)
                evs.append((a, b))
            torch.cuda.synchronize()
            enc_ms[name] = statistics.mean(a.elapsed_time(b) for a, b in evs)
        enc_bytes, enc_raw, enc_blocks = dev.bcf_length()
    dev.check()
    n_worklist = dev.work_count()
    n_fused = dev.fused_count()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    launches = dev.launch_count * args.steps
    h2d_one = dev.in_bytes
    d2h_cols = dev.out_bytes
    dev.close()
    ctx.close()

    # ---- end to end through the C ABI with host buffers: the streamed range-batch driver over the rank's share of the cohort ----
    modes = [m for m in args.e2e_modes.split(",") if m]
    if world > 1:
        modes = modes[:1]
    t_wire = time.time()
    wire_bytes = sum(b.wire()[1] for b in batches)  # built (and page-locked) once per batch, outside the timed region, like the packing itself
    wire_s = time.time() - t_wire

    def run_mode(m):
        base, _, opt = m.partition("+")
        return stream_pass(glnexus_b200, local_rank, base, batches, args.passes, depth=args.depth, wire=opt == "wire")

    e2e = {}
    solo = None
    if world > 1:  # rank 0 alone first: the one-GPU rate of the same code on the same box, for the scaling efficiency
        if rank == 0:
            solo = run_mode(modes[0])
        barrier()
    for m in modes:
        barrier()
        dt, st = run_mode(m)
        barrier()
        e2e[m] = (dt, st)

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    total_ms = allmax(total_ms)
    cells_all = int(allsum(cells))
    e2e_red = {}
    for m in modes:
        dt, st = e2e[m]
        e2e_red[m] = {"s": allmax(dt), "cells": int(allsum(st["cells"])), "h2d": int(allsum(st["h2d_bytes"])), "d2h": int(allsum(st["d2h_bytes"])),
                      "raw": int(allsum(st["raw_bcf_bytes"])), "batches": int(allsum(st["batches"]))}

    if rank == 0:
        value = cells_all * args.steps / (total_ms * 1e-3)
        head = e2e_red[modes[0]]
        e2e_v = head["cells"] / head["s"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        names = ["glx_resolve_records_kernel", "glx_cells_tiled_sig_kernel", "closed-form worklist kernels (glx_cells_bands/variant/single_kernel)",
                 "glx_cells_general_kernel"]
        kmean = [statistics.mean(k[i] for k in kms) for i in range(4)]
        n_var_cells = int(batch.variant_records)
        alg, out_b, in_b = algorithmic_bytes(batch, n_var_cells)
        # dominant kernel and ITS algorithmic bytes (SURVEY §8d per-cell figures x the cells that kernel finishes):
        # the tiled kernel finishes the reference-band / uncovered cells (34 B in + out), the worklist kernels the rest
        dom_i = max(range(4), key=lambda i: kmean[i])
        out_per_cell = out_b / cells
        fast_cells = cells - n_worklist
        alg_by_kernel = [None, fast_cells * (34 + out_per_cell) + n_fused * (61 + out_per_cell), None, None]
        alg_by_kernel[2] = (n_worklist - n_fused) * (61 + out_per_cell)
        dom_ms = kmean[dom_i]
        dom = names[dom_i]
        dom_alg = alg_by_kernel[dom_i] if alg_by_kernel[dom_i] is not None else alg
        achieved = dom_alg / (dom_ms * 1e-3) / 1e9
        traffic = None
        for tf in ("r2_traffic.json", "r1_traffic.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", tf)))
                if tr.get("workload") == args.workload and tr.get("cells") == cells:
                    traffic = tr["dram_bytes"].get(dom)
                    break
            except Exception:
                pass
        if traffic is None:  # other BASELINE shapes: the per-kernel DRAM bytes of profiles/r2_traffic_c345.json (same shape and sizes only)
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic_c345.json")))["dram_bytes"].get(args.workload, {})
                if (n_samples, n_sites) == WORKLOADS[args.workload][1:3]:
                    key = "tiled_sig" if dom_i == 1 else ("resolve" if dom_i == 0 else None)
                    traffic = next((v for k, v in tr.items() if key and key in k), None)
            except Exception:
                pass
        n_b = max(1, head["batches"])
        line = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, n_samples, n_sites), "preset": synth_cases.SHAPES[shape][0]},
            "details": {"cells_per_step_per_gpu": cells, "numa_node": numa, "records_per_step_per_gpu": int(batch.n_records),
                        "variant_records_per_step_per_gpu": int(batch.variant_records),
                        "cohort": f"{n_segments} consecutive segments of {n_sites} sites x {n_samples} samples, cut into {world} contiguous site ranges of equal "
                                  f"output weight (glxh_partition_cohort); this line's rank 0 range = segments/sites {cuts[0]}..{cuts[1]}",
                        "synthetic_generation_s": gen_s,
                        "l2": "flushed between timed steps (256 MiB write, untimed); outputs (%.2f GB/step) exceed L2" % (out_b / 1e9),
                        "value_excludes": "the two upload-side kernels (glx_prepare_inputs_kernel %.3f ms, glx_block_hints_kernel %.3f ms per upload) and three O(records)/O(sites) "
                                          "host loops of glx_upload_on; e2e includes them" % (up_ms[0], up_ms[1])},
            "e2e": {"value": e2e_v, "unit": "cells/s", "h2d_bytes_per_step": int(head["h2d"] / n_b), "d2h_bytes_per_step": int(head["d2h"] / n_b),
                    "steps": int(n_b), "cells": head["cells"], "seconds": head["s"], "passes_over_the_cohort": args.passes,
                    "mode": ("glxh_stream_* (C++ range-batch driver, %d batches in flight): per range batch H2D of the record store in its compact wire form (glx_wire) + sites "
                             "from pinned host memory, expansion + all genotyping kernels, the on-device BCF2 record encoder + BGZF (DEFLATE, CRC-32), D2H of the finished BGZF "
                             "blocks into pinned host memory; every rank streams its own contiguous site range of one cohort") % args.depth,
                    "wire_build_s_untimed": wire_s, "wire_bytes_per_record": wire_bytes / max(1, sum(b.n_records for b in batches)),
                    "host_packing": {"wire_build_ms_per_batch": wire_s / max(1, len(batches)) * 1e3, "threads": min(16, os.cpu_count() or 1),
                                     "what": "full-width record store -> wire form (glx::build_wire) + cudaHostRegister of the buffer, once per range batch, "
                                             "outside the timed region (a production caller builds batch i+1 while batch i is on the device; at this rate the "
                                             "host side, not the GPU, bounds a single stream)"},
                    "d2h_bytes_per_cell": head["d2h"] / max(1, head["cells"]), "h2d_bytes_per_cell": head["h2d"] / max(1, head["cells"]),
                    "uncompressed_bcf_bytes_per_cell": head["raw"] / max(1, head["cells"]),
                    "other_modes": {m: {"cells_per_s": e2e_red[m]["cells"] / e2e_red[m]["s"], "d2h_bytes_per_cell": e2e_red[m]["d2h"] / max(1, e2e_red[m]["cells"])}
                                    for m in modes[1:]}},
            "gpu_launches": launches,
            "kernels": {"resolve_ms": kmean[0], "tiled_ms": kmean[1], "single_ms": kmean[2], "general_ms": kmean[3], "worklist_cells": int(n_worklist),
                        "fused_cells": int(n_fused), "upload_prepare_ms": up_ms[0], "upload_hints_ms": up_ms[1],
                        "step_ms_min": min(step_ms), "step_ms_max": max(step_ms), **enc_ms,
                        "bcf_bytes_per_cell": enc_raw / cells, "bgzf_bytes_per_cell": enc_bytes / cells, "bgzf_blocks": enc_blocks},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_alg,
                         "definition": "SURVEY 8(d): 34 B in per reference-band cell (61 B per variant cell) + output bytes, x the cells this kernel finishes "
                                       "(the tiled kernel: its band cells and the lone-variant-record cells it finishes at the end of each tile)",
                         "whole_step": {"algorithmic_bytes": alg, "bytes_out": out_b, "bytes_in": in_b, "achieved": alg / (total_ms / args.steps * 1e-3) / 1e9,
                                        "frac": alg / (total_ms / args.steps * 1e-3) / 1e9 / peak}},
            "clocks": clocks,
        }
        if solo is not None:
            one = solo[1]["cells"] / solo[0]
            line["e2e"]["one_gpu_same_run_cells_per_s"] = one
            line["e2e"]["e2e_efficiency"] = e2e_v / (world * one)
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core back
            cb, _ = cpu_baseline(shape, n_samples, region_len / n_sites, n_sites)
            line["cpu_baseline"] = cb
        else:
            line["cpu_baseline"] = None
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    for b in batches:
        b.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
