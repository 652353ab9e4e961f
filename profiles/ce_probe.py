"""Do a pinned H2D and a pinned D2H copy on two streams overlap on this box? (round-2 question about the streamed driver: profiles/r2_summary.md)"""
import torch, time
n1, n2 = 171_000_000, 144_000_000
h1 = torch.empty(n1, dtype=torch.uint8).pin_memory(); d1 = torch.empty(n1, dtype=torch.uint8, device='cuda')
h2 = torch.empty(n2, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n2, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(a, b, reps=5):
    torch.cuda.synchronize(); t = time.time()
    for _ in range(reps):
        if a:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if b:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); return (time.time() - t) / reps * 1e3
for _ in range(2): run(True, True)
print("H2D alone ms", run(True, False), "GB/s", n1 / run(True, False) / 1e6)
print("D2H alone ms", run(False, True), "GB/s", n2 / run(False, True) / 1e6)
print("both ms", run(True, True))
# event-level: does an H2D queued right after a D2H (other stream) start immediately?
e0, e1, e2, e3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True); base.record(); 
with torch.cuda.stream(s2):
    e0.record(); h2.copy_(d2, non_blocking=True); e1.record()
with torch.cuda.stream(s1):
    e2.record(); d1.copy_(h1, non_blocking=True); e3.record()
torch.cuda.synchronize()
print("D2H", base.elapsed_time(e0), base.elapsed_time(e1), " H2D (queued after, other stream)", base.elapsed_time(e2), base.elapsed_time(e3))
print("asyncEngineCount", torch.cuda.get_device_properties(0).multi_processor_count)
