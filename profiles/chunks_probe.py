#!/usr/bin/env python3
"""Step time of one BASELINE-shaped batch against the context option `chunks` (the tile grid launched in k pieces, the worklist
kernels of piece j on the context's second stream while the tiled kernel of piece j+1 runs). Usage: chunks_probe.py [c2|c3|c4|c5]...
Prints min / median ms over 10 launches per setting, L2 flushed between launches, and checks that the outputs do not change."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import synth_cases  # noqa: E402
from glnexus_b200 import Context  # noqa: E402

SHAPES = {"c2": (1000, 50000, 50000), "c3": (10000, 4096, 262144), "c4": (100000, 512, 32768), "c5": (5000, 10000, 300000)}


def main():
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    for w in sys.argv[1:] or ["c5"]:
        ns, nsites, region = SHAPES[w]
        b = synth_cases.make(w, n_samples=ns, n_sites=nsites, region_len=region)
        ref = None
        for chunks in (1, 2, 4, 8):
            ctx = Context(0)
            ctx.set_option("chunks", chunks)
            dev = ctx.upload(b)
            ms = []
            for i in range(13):
                flush.fill_(i & 0xff)
                torch.cuda.synchronize()
                t0 = time.perf_counter()  # (the launch runs on the context's own stream: wall clock around launch + check, which waits for it)
                dev.launch()
                dev.check()
                t1 = time.perf_counter()
                if i >= 3:
                    ms.append((t1 - t0) * 1e3)
            out = b.alloc_out()
            dev.download(out)
            arrs = out.arrays()
            same = True
            if ref is None:
                ref = {k: v.copy() for k, v in arrs.items()}
            else:
                same = all(np.array_equal(ref[k], arrs[k]) for k in ref)
            print(f"{w} chunks {chunks}: step min/median {min(ms):.4f}/{statistics.median(ms):.4f} ms, outputs {'identical' if same else 'DIFFER'}", flush=True)
            dev.close()
            ctx.close()


if __name__ == "__main__":
    main()
