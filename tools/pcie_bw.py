"""Host <-> device copy bandwidth with pinned memory (the ceiling of bench.py's e2e number)."""
import time
import torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (a, b) in (("d2h", (h, d)), ("h2d", (d, h))):
    for _ in range(2):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (5 * n / (time.perf_counter() - t) / 1e9))
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print("bidirectional: %.1f GB/s each way" % (5 * n / dt / 1e9))
