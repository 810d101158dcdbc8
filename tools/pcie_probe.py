"""Host<->device copy bandwidth of the box with page-locked buffers (development probe for the e2e overlap design)."""
import torch
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: d.copy_(h, non_blocking=True)); print("H2D %.1f GB/s" % (n / ms / 1e6))
ms = t(lambda: h.copy_(d, non_blocking=True)); print("D2H %.1f GB/s" % (n / ms / 1e6))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
ms = t(both); print("H2D + D2H concurrently: %.1f GB/s each direction" % (n / ms / 1e6))
