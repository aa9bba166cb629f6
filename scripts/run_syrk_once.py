"""One SYRK over 262144 x 768 activations -- profiling target for ncu."""
import sys, torch
sys.path.insert(0, ".")
from bayesvlm_b200.hessians import syrk_accumulate
X = torch.randn(262144, 768, device="cuda")
out = torch.zeros(768, 768, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    syrk_accumulate(X, out=out)
torch.cuda.synchronize(); print("ok", float(out[0, 0]))
