"""Accuracy of the A-factor SYRK against fp64 as the row count (K extent of the tensor-core accumulation) grows:
python scripts/syrk_bias.py"""
import sys, torch
sys.path.insert(0, ".")
from bayesvlm_b200.hessians import syrk_accumulate
gen = torch.Generator(device="cuda").manual_seed(1)
d = 768
for n in (32768, 131072, 524288, 1048576):
    X = torch.randn(n, d, generator=gen, device="cuda")
    A = syrk_accumulate(X)
    ref = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    for lo in range(0, n, 65536):
        xb = X[lo:lo + 65536].double()
        ref += xb.T @ xb
    diag_rel = ((A.diagonal().double() - ref.diagonal()) / ref.diagonal())
    print(n, "fro rel err %.3e" % float((A.double() - ref).norm() / ref.norm()), "diag signed mean rel %.3e" % float(diag_rel.mean()),
          "diag max abs rel %.3e" % float(diag_rel.abs().max()))
