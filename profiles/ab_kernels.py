#!/usr/bin/env python3
"""A/B aid: kernel times of one BASELINE-shaped batch under the library GLX_LIB_PATH names (default: the product build).
Usage: [GLX_LIB_PATH=build_alt/libglx_X.so] ab_kernels.py [c2|c3|c4|c5] [launches]. Prints min and median of
(resolve, tiled, closed-form worklist, general worklist) ms, a flush of L2 between launches."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import synth_cases  # noqa: E402
from glnexus_b200 import Context  # noqa: E402

SHAPES = {"c2": (1000, 50000, 50000), "c3": (10000, 4096, 262144), "c4": (100000, 512, 32768), "c5": (5000, 10000, 300000)}


def main():
    w = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    ns, nsites, region = SHAPES[w]
    b = synth_cases.make(w, n_samples=ns, n_sites=nsites, region_len=region)
    ctx = Context(0)
    dev = ctx.upload(b)
    dev.set_profile(True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    rows = []
    for i in range(n + 3):
        flush.fill_(i & 0xff)
        torch.cuda.synchronize()
        dev.launch()
        dev.check()
        if i >= 3:
            rows.append(dev.kernel_ms())
    names = ("resolve", "tiled", "worklist", "general")
    out = ", ".join(f"{nm} {min(r[k] for r in rows):.4f}/{statistics.median(r[k] for r in rows):.4f}" for k, nm in enumerate(names))
    tot = [sum(r) for r in rows]
    print(f"{os.environ.get('GLX_LIB_PATH', 'libglx.so'):40s} {w}: min/median ms: {out}; sum {min(tot):.4f}/{statistics.median(tot):.4f}")
    dev.close()
    ctx.close()


if __name__ == "__main__":
    main()
