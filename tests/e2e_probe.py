"""Phase timing of the pipelined end-to-end loop (diagnostic, not a test): host time spent in each C ABI call."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa
import synth_cases
from glnexus_b200 import Context

ctx = Context(0)
batch = synth_cases.make("c2", n_samples=1000, n_sites=50000, region_len=50000)
batch.pin()
outs16 = [batch.alloc_out16(), batch.alloc_out16()]
streams = [ctx.stream_create(), ctx.stream_create()]
T = {"upload": 0.0, "launch": 0.0, "download": 0.0, "sync": 0.0, "finish": 0.0, "close": 0.0, "status": 0.0}


def run(n):
    slots = [None, None]

    def retire(k):
        t = time.time(); ctx.stream_sync(streams[k]); T["sync"] += time.time() - t
        t = time.time(); slots[k].status(); T["status"] += time.time() - t
        t = time.time(); slots[k].narrow_finish(outs16[k]); T["finish"] += time.time() - t
        t = time.time(); slots[k].close(); T["close"] += time.time() - t
        slots[k] = None

    for i in range(n):
        k = i & 1
        if slots[k] is not None:
            retire(k)
        t = time.time(); dv = ctx.upload_async(batch, streams[k]); T["upload"] += time.time() - t
        t = time.time(); dv.launch(streams[k]); T["launch"] += time.time() - t
        t = time.time(); dv.download_narrow_async(outs16[k], streams[k]); T["download"] += time.time() - t
        slots[k] = dv
    for k in (0, 1):
        if slots[k] is not None:
            retire(k)


run(2)
for k in T: T[k] = 0.0
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 8
t0 = time.time(); run(NS); torch.cuda.synchronize(); tot = time.time() - t0
print("total ms/step %.2f" % (tot / NS * 1e3), {k: round(v / NS * 1e3, 2) for k, v in T.items()})
