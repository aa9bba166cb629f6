"""Event-timed quadratic-form and mean GEMM launches of the headline predictive step (same-process comparison aid)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200 import _lib
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
from bayesvlm_b200.vlm import CLIP, EncoderResult
cfg = bench.PRED
t = bench.predictive_inputs(cfg, 0)
Ai, Bi, At, Bt = bench.covariances(t, cfg, "cuda")
m = CLIP(logit_scale=bench.LS, device="cuda")
m.set_covariances(KFC(Ai, Bi), KFC(At, Bt))
img = EncoderResult(t["img_e"].cuda(), t["img_a"].cuda()); txt = EncoderResult(t["txt_e"].cuda(), t["txt_a"].cuda())
mean = torch.empty((cfg["N"], cfg["C"]), device="cuda"); var = torch.empty_like(mean)
with torch.no_grad():
    for _ in range(20):
        m._smith_into(img.embeds, img.activations, txt, mean, var)
    torch.cuda.synchronize(); _lib.timing_enable(True)
    for _ in range(100):
        m._smith_into(img.embeds, img.activations, txt, mean, var)
    torch.cuda.synchronize(); _lib.timing_enable(False); k = _lib.timing_collect()
print({n: round(v[1] / v[0], 4) for n, v in k.items()}, float(mean[0, 0]), float(var[-1, -1]))
