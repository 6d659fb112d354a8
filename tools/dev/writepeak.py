import torch
x = torch.empty(6 * 1024**3, dtype=torch.uint8, device="cuda")
y = torch.empty(6 * 1024**3, dtype=torch.uint8, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.fill_(7))
print("fill  6 GiB: %.3f ms  %.0f GB/s written" % (ms, x.numel() / ms / 1e6))
ms = t(lambda: y.copy_(x))
print("copy  6 GiB: %.3f ms  %.0f GB/s read+written" % (ms, 2 * x.numel() / ms / 1e6))
ms = t(lambda: x.sum(dtype=torch.int64))
print("sum   6 GiB: %.3f ms  %.0f GB/s read" % (ms, x.numel() / ms / 1e6))
