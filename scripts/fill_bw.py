"""Write-only HBM bandwidth of this GPU (torch fill), the ceiling for kernels that only write (labels)."""
import torch
for gib in (1, 4, 8):
    o = torch.empty(gib << 30, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        o.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(10):
        e0.record(); o.zero_(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("fill %d GiB: %.3f ms, %.1f GB/s" % (gib, best, (gib << 30) / best / 1e6))
    del o
