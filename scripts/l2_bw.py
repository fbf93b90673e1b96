"""Development aid: copy bandwidth as a function of the working set (L2-resident vs HBM), torch copy_ only."""
import torch
def run(mib, iters=200):
    n = mib * (1 << 20) // 8 // 2
    a = torch.empty(n, dtype=torch.float64, device="cuda").normal_()
    b = torch.empty_like(a)
    for _ in range(5): b.copy_(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("working set %5d MiB: %.1f GB/s (read+write)" % (mib, 2 * n * 8 / ms / 1e6), flush=True)
for mib in (8, 16, 32, 48, 64, 96, 128, 256, 1024, 4096):
    run(mib, iters=200 if mib <= 256 else 20)
