"""HBM write / copy / read bandwidth with plain torch ops (what a store-bound kernel can hope for)."""
import torch
x = torch.empty(4 << 30, dtype=torch.uint8, device='cuda')
y = torch.empty(4 << 30, dtype=torch.uint8, device='cuda')
for name, fn, nbytes in (("fill", lambda: x.fill_(1), 4 << 30), ("memset", lambda: x.zero_(), 4 << 30), ("copy", lambda: y.copy_(x), 8 << 30),
                        ("read(sum)", lambda: x.view(torch.int64).sum(), 4 << 30)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, nbytes * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e12, "TB/s")
