"""Sustained time of the headline predictive step (50k x 1000, device-resident inputs): python scripts/pred_step_time.py [steps]"""
import sys, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
from bayesvlm_b200.vlm import CLIP, EncoderResult
cfg = bench.PRED
t = bench.predictive_inputs(cfg, 0)
Ai, Bi, At, Bt = bench.covariances(t, cfg, "cuda")
m = CLIP(logit_scale=bench.LS, device="cuda")
m.set_covariances(KFC(Ai, Bi), KFC(At, Bt))
img = EncoderResult(t["img_e"].cuda(), t["img_a"].cuda()); txt = EncoderResult(t["txt_e"].cuda(), t["txt_a"].cuda())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
mean = torch.empty((cfg["N"], cfg["C"]), device="cuda"); var = torch.empty_like(mean)
with torch.no_grad():
    for _ in range(10):
        m._smith_into(img.embeds, img.activations, txt, mean, var)
    res = []
    for rep in range(3):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            m._smith_into(img.embeds, img.activations, txt, mean, var)
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / n)
print("ms/step", " ".join(f"{r:.4f}" for r in res), "checksum", float(mean[0, 0]), float(var[-1, -1]))
