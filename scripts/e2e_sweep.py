"""End-to-end predict_host time vs image batch size at the headline shape (PCIe pipeline fill/drain vs launch count)."""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
from bayesvlm_b200.vlm import CLIP, EncoderResult
cfg = bench.PRED
t = bench.predictive_inputs(cfg, 0)
Ai, Bi, At, Bt = bench.covariances(t, cfg, "cuda")
m = CLIP(logit_scale=bench.LS, device="cuda")
m.set_covariances(KFC(Ai, Bi), KFC(At, Bt))
img = EncoderResult(t["img_e"].pin_memory(), t["img_a"].pin_memory()); txt = EncoderResult(t["txt_e"].pin_memory(), t["txt_a"].pin_memory())
out = (torch.empty((cfg["N"], cfg["C"]), pin_memory=True), torch.empty((cfg["N"], cfg["C"]), pin_memory=True))
pageable = EncoderResult(t["img_e"], t["img_a"])  # what reference-style callers pass (bounced through pinned buffers)
for _ in range(2):
    m.predict_host(pageable, txt, batch_size=2048, out=out)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(8):
    m.predict_host(pageable, txt, batch_size=2048, out=out)
torch.cuda.synchronize()
print("pageable inputs, batch 2048:", round((time.perf_counter() - t0) / 8 * 1e3, 3), "ms", flush=True)
for bs in [int(a) for a in sys.argv[1:]] or [1563, 2048, 3125, 4167, 6250, 12500]:
    for _ in range(2):
        m.predict_host(img, txt, batch_size=bs, out=out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(24):
        m.predict_host(img, txt, batch_size=bs, out=out)
    torch.cuda.synchronize()
    print(bs, round((time.perf_counter() - t0) / 24 * 1e3, 3), "ms", flush=True)
