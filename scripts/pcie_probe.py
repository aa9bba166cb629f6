"""PCIe floor of the end-to-end predictive step: 364 MB host->device and 400 MB device->host, alone and concurrently."""
import time, torch
h_in = torch.empty(364_544_000 // 4, dtype=torch.float32, pin_memory=True); d_in = torch.empty_like(h_in, device="cuda")
d_out = torch.empty(400_000_000 // 4, dtype=torch.float32, device="cuda"); h_out = torch.empty(d_out.shape, dtype=torch.float32, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for _ in range(2): run(True, True, 2)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"h2d alone {a:.2f} ms ({0.364544 / a * 1e3:.1f} GB/s)  d2h alone {b:.2f} ms ({0.4 / b * 1e3:.1f} GB/s)  both {c:.2f} ms")
