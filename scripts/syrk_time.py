"""K1 alone: python scripts/syrk_time.py [d=768] -- tensor-core launch time / fraction of the sustained peak over row counts."""
import sys, torch
sys.path.insert(0, ".")
from bayesvlm_b200 import _lib
from bayesvlm_b200.hessians import syrk_accumulate
d = int(sys.argv[1]) if len(sys.argv) > 1 else 768
for rows in (32768, 65536, 131072, 262144, 524288, 1048576):
    X = torch.randn(rows, d, device="cuda")
    out = torch.zeros(d, d, device="cuda")
    for _ in range(3):
        syrk_accumulate(X, out=out)
    torch.cuda.synchronize(); _lib.timing_enable(True)
    for _ in range(5):
        syrk_accumulate(X, out=out)
    torch.cuda.synchronize(); _lib.timing_enable(False)
    k = _lib.timing_collect()["syrk"]
    ms = k[1] / k[0]
    ref = X[:2048].double().T @ X[:2048].double()
    err = float((syrk_accumulate(X[:2048]).double() - ref).norm() / ref.norm())
    print(f"rows {rows:8d}  gemm {ms:.4f} ms  {rows * d * (d + 1.0) / (ms * 1e-3) / 1e12:7.1f} TFLOP/s (symmetric)  frac {rows * d * (d + 1.0) / (ms * 1e-3) / 1e12 / 1361.2:.3f}  err {err:.1e}")
    del X
