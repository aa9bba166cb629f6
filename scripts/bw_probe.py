import torch, time
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")  # 1 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); b = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return b.elapsed_time(e) / n
gb = x.numel() * 4 / 1e9
print("fill (write only)  %.0f GB/s" % (gb / t(lambda: x.fill_(1.0)) * 1e3))
print("sum  (read only)   %.0f GB/s" % (gb / t(lambda: x.sum()) * 1e3))
print("copy (read+write)  %.0f GB/s total" % (2 * gb / t(lambda: y.copy_(x)) * 1e3))
