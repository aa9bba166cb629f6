"""Per-kernel times of one class batch (B = C = 32768): python scripts/kfac_time.py [D=512] [siglip]"""
import math, sys, torch
sys.path.insert(0, ".")
from bayesvlm_b200 import _lib
from bayesvlm_b200.hessians import _ggn
n, D = 32768, int(sys.argv[1]) if len(sys.argv) > 1 else 512
siglip = len(sys.argv) > 2 and sys.argv[2] == "siglip"
ls, lb = (4.765, -12.93) if siglip else (math.log(100.0), 0.0)
gen = torch.Generator(device="cuda").manual_seed(1)
z = torch.randn(n, D, generator=gen, device="cuda")
X = z + 1.5 * torch.randn(n, D, generator=gen, device="cuda")
Y = z + 1.5 * torch.randn(n, D, generator=gen, device="cuda")
for _ in range(3):
    H = _ggn(X, Y, ls, lb, siglip, precision="fp16")
for rep in range(2):
    torch.cuda.synchronize(); _lib.timing_enable(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        H = _ggn(X, Y, ls, lb, siglip, precision="fp16")
    b.record(); torch.cuda.synchronize(); _lib.timing_enable(False); k = _lib.timing_collect()
    print("class batch %.3f ms" % (a.elapsed_time(b) / 5), {t: round(v[1] / v[0], 3) for t, v in k.items()}, float(H.abs().max()))
