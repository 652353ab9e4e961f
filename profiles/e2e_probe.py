#!/usr/bin/env python3
"""One process, one GPU (CUDA_VISIBLE_DEVICES picks it): stream a few c2 batches through the range-batch driver and print the
rate. Run several at once to see what the host side does to concurrent ranks. Usage: e2e_probe.py [mode] [n_batches] [wire]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth_cases
import glnexus_b200

mode = sys.argv[1] if len(sys.argv) > 1 else "bgzf"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 6
wire = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
start_at = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
batches = [synth_cases.make("c2", n_samples=1000, n_sites=50000, region_len=50000, region_beg=10_000_000 + g * 50000, seed=20200210 + g) for g in range(3)]
for b in batches:
    b.pin()
    if wire:
        b.wire()
rs = glnexus_b200.RangeStream(0, mode, 3, wire=wire)
for b in batches:
    rs.submit(b)
rs.drain()
while time.time() < start_at:
    time.sleep(0.001)
t = time.time()
for i in range(nb):
    rs.submit(batches[i % 3])
rs.drain()
dt = time.time() - t
st = rs.stats()
print(f"pid {os.getpid()} dev {os.environ.get('CUDA_VISIBLE_DEVICES')} {mode} wire={wire}: {nb * 5e7 / dt:.3e} cells/s, {dt / nb * 1e3:.2f} ms per batch")
