"""Probit softmax at the headline shape (50k x 1000) -- profiling target for ncu."""
import sys, torch
sys.path.insert(0, ".")
from bayesvlm_b200.vlm import probit_softmax
g = torch.Generator(device="cuda").manual_seed(1)
mean = torch.randn(50000, 1000, generator=g, device="cuda") * 4
var = torch.rand(50000, 1000, generator=g, device="cuda") * 3 + 0.1
for _ in range(2):
    p = probit_softmax(mean, var)
torch.cuda.synchronize(); print("ok", float(p[0].sum()))
