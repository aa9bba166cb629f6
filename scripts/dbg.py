import math, sys, numpy as np, torch
sys.path.insert(0, ".")
from bayesvlm_b200.hessians import _ggn
from bayesvlm_b200 import _lib
g = dict(np.load("tests/golden/reference_small.npz"))
X = torch.from_numpy(g["ggn_X"]).cuda(); Y = torch.from_numpy(g["ggn_Y"]).cuda()
for prec in ("fp16", "fp16x3"):
    H = _ggn(X, Y, math.log(100.0), 0.0, False, precision=prec)
    torch.cuda.synchronize()
    print(prec, torch.isfinite(H).all().item(), H.abs().max().item(), np.abs(g["ggn_infonce_H64"]).max())
