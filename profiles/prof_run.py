#!/usr/bin/env python3
"""Minimal launcher for ncu captures: build one synthetic batch of a BASELINE shape, upload it once and run the hot path
`launches` times on the context stream. Usage: prof_run.py <c2|c3|c4|c5> [launches] [samples] [sites] [encoder mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_cases  # noqa: E402
from glnexus_b200 import Context  # noqa: E402

SHAPES = {"c2": (1000, 50000, 50000), "c3": (10000, 4096, 262144), "c4": (100000, 512, 32768), "c5": (5000, 10000, 300000)}


def main():
    w = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    ns, nsites, region = SHAPES[w]
    if len(sys.argv) > 3:
        ns = int(sys.argv[3])
    if len(sys.argv) > 4:
        region = int(region * int(sys.argv[4]) / nsites)
        nsites = int(sys.argv[4])
    b = synth_cases.make(w, n_samples=ns, n_sites=nsites, region_len=region)
    ctx = Context(0)
    dev = ctx.upload(b)
    enc = int(sys.argv[5]) if len(sys.argv) > 5 else -1  # -1: hot path only; 0 / 1: also the BCF encoder (raw records / BGZF)
    for _ in range(n):
        dev.launch()
    dev.check()
    if enc >= 0:
        dev.bcf_attach(b)
        for _ in range(n):
            dev.encode_bcf(enc)
        print("encoded bytes, raw bytes, blocks:", dev.bcf_length())
        if enc == 1:
            cyc = ctx.bgzf_phase_cycles()
            names = ["gather", "segments", "setup", "crc parts", "tokenize", "crc tree", "codes+header", "count+scan", "emit", "finish+store"]
            tot = sum(cyc) or 1
            tot = sum(cyc[:10]) or 1
            print("BGZF kernel, thread-0 cycles per phase:", ", ".join(f"{nm} {100 * c / tot:.1f}%" for nm, c in zip(names, cyc)),
                  f"; {tot / max(1, n * dev.bcf_length()[2]):.0f} cycles per block")
            if any(cyc[10:]):  # -DGLX_FILL_PROFILE builds: parts of the gather (these also count the sampling kernel's blocks)
                sub = ["stage loads", "serial part", "piece table", "(build total)", "fill", "closing barrier"]
                print("  gather parts, % of the same total:", ", ".join(f"{nm} {100 * c / tot:.1f}%" for nm, c in zip(sub, cyc[10:])))
    print(f"{w}: {b.n_sites} sites x {b.n_samples} samples, {b.n_records} records, {dev.launch_count} kernels per launch")
    dev.close()
    ctx.close()


if __name__ == "__main__":
    main()
