"""PCIe probe (development tool): pinned host -> device copy bandwidth and small-copy latency on this box."""
import time
import torch
dev = torch.device("cuda:0")
for mb in (0.25, 2, 8, 64):
    n = int(mb * (1 << 20))
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    print("H2D %6.2f MiB: %.3f ms device (%.1f GB/s), %.3f ms wall" % (mb, ms, n / ms / 1e6, wall))
    e0.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("D2H %6.2f MiB: %.3f ms device (%.1f GB/s)" % (mb, ms, n / ms / 1e6))
