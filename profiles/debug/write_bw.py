"""Debug helper: write-only HBM bandwidth (torch fill) vs copy bandwidth, the denominator question for K1."""
import torch
x = torch.empty(512 * 1024 * 1024, dtype=torch.float32, device="cuda")   # 2 GiB
y = torch.empty_like(x)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.0)); print(f"fill   : {x.numel()*4/ms/1e6:8.1f} GB/s written")
ms = t(lambda: x.zero_());    print(f"memset : {x.numel()*4/ms/1e6:8.1f} GB/s written")
ms = t(lambda: y.copy_(x));   print(f"copy   : {2*x.numel()*4/ms/1e6:8.1f} GB/s read+written")
ms = t(lambda: x.sum());      print(f"read   : {x.numel()*4/ms/1e6:8.1f} GB/s read")
