"""One InfoNCE class batch (B = C = 32768, D = 512) through the GGN pipeline -- profiling target for ncu."""
import math, sys, torch
sys.path.insert(0, ".")
from bayesvlm_b200.hessians import _ggn
n, D = 32768, 512
gen = torch.Generator(device="cuda").manual_seed(1)
z = torch.randn(n, D, generator=gen, device="cuda")
X = z + 1.5 * torch.randn(n, D, generator=gen, device="cuda")
Y = z + 1.5 * torch.randn(n, D, generator=gen, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    H = _ggn(X, Y, math.log(100.0), 0.0, False, precision="fp16")
torch.cuda.synchronize()
print("ok", float(H.abs().max()))
