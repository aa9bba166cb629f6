"""GPU probe: which fp16 rounding points does torch's OWN CUDA path of the reference EPIG sequence use, and how often do the
fused kernels reproduce it bit for bit?  Writes gpurun_out/epig_probe.json (+ .npz tables).

    python scripts/probe_epig_parity.py [--quick]

Not a test and not part of the product: a measurement helper (the parity tests proper are tests/test_gpu_epig.py).
"""
import json
import math
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)

from oracle import torch_port as T  # noqa: E402  (checker only)

from bayesvlm_b200 import epig as E  # noqa: E402
from bayesvlm_b200.vlm import ProbabilisticLogits, sample_probas_from_noise  # noqa: E402

dev = torch.device("cuda")
res = {}


def f16(x):
    return np.asarray(x, dtype=np.float32).astype(np.float16)


# ---------------------------------------------------------------------------------------------------- 1. elementwise tables
bits = np.arange(1, 0x3C01, dtype=np.uint16)  # every fp16 in (0, 1]
x16 = torch.from_numpy(bits.view(np.float16)).to(dev)
xl_gpu = torch.xlogy(x16, x16).cpu().numpy()
xl_cpu = torch.xlogy(x16.cpu(), x16.cpu()).numpy()
x32 = bits.view(np.float16).astype(np.float32)
log32 = np.log(x32.astype(np.float64)).astype(np.float32)
h1 = f16(x32 * f16(log32).astype(np.float32))          # log rounded to fp16 first
h2 = f16((x32.astype(np.float64) * np.log(x32.astype(np.float64))))  # one rounding
h3 = f16(x32 * log32)                                  # fp32 log, fp32 product, one fp16 rounding
res["xlogy"] = {"n": int(bits.size),
                "gpu_eq_cpu": int((xl_gpu.view(np.uint16) == xl_cpu.view(np.uint16)).sum()),
                "gpu_eq_logf16_then_mul": int((xl_gpu.view(np.uint16) == h1.view(np.uint16)).sum()),
                "gpu_eq_single_rounding_f64": int((xl_gpu.view(np.uint16) == h2.view(np.uint16)).sum()),
                "gpu_eq_single_rounding_f32": int((xl_gpu.view(np.uint16) == h3.view(np.uint16)).sum()),
                "cpu_eq_logf16_then_mul": int((xl_cpu.view(np.uint16) == h1.view(np.uint16)).sum()),
                "cpu_eq_single_rounding_f32": int((xl_cpu.view(np.uint16) == h3.view(np.uint16)).sum())}
lg_gpu = torch.log(x16).cpu().numpy()
res["log"] = {"gpu_eq_f16_of_log": int((lg_gpu.view(np.uint16) == f16(log32).view(np.uint16)).sum())}

allpos = np.arange(1, 0x7C00, dtype=np.uint16)  # every positive finite fp16
a16 = torch.from_numpy(allpos.view(np.float16)).to(dev)
a32 = allpos.view(np.float16).astype(np.float32)
res["div"] = {}
for k in (100, 16, 33, 64, 2000, 10000, 129, 700):
    g = (a16 / k).cpu().numpy().view(np.uint16)
    c = (a16.cpu() / k).numpy().view(np.uint16)
    true_div = f16(a32 / np.float32(k)).view(np.uint16)
    mul_inv = f16(a32 * np.float32(1.0 / k)).view(np.uint16)
    res["div"][str(k)] = {"n": int(allpos.size), "gpu_eq_true_div": int((g == true_div).sum()),
                          "gpu_eq_mul_by_f32_reciprocal": int((g == mul_inv).sum()),
                          "cpu_eq_true_div": int((c == true_div).sum()), "gpu_eq_cpu": int((g == c).sum())}

# mean over K of fp16 values: fp32 sum then * (1/K) or / K ?
gen = torch.Generator(device=dev).manual_seed(1)
p = torch.softmax(torch.randn(20000, 100, 10, generator=gen, device=dev) * 2, dim=-1).half()
m_gpu = p.mean(dim=1).cpu().numpy().view(np.uint16)
s64 = p.double().sum(dim=1).cpu().numpy()
s32 = s64.astype(np.float32)
res["mean_K100"] = {"n": int(m_gpu.size), "eq_sum_div": int((m_gpu == f16(s32 / np.float32(100)).view(np.uint16)).sum()),
                    "eq_sum_mul_recip": int((m_gpu == f16(s32 * np.float32(1.0 / 100)).view(np.uint16)).sum())}

# -(sum) / N_t on fp16: which rounding?
tot16 = torch.from_numpy(np.arange(0x6000, 0x6C00, dtype=np.uint16).view(np.float16)).to(dev)  # 512 .. 4096
for nt in (2000, 10000):
    g = (-tot16 / nt).cpu().numpy().view(np.uint16)
    t32 = tot16.cpu().numpy().astype(np.float32)
    res[f"neg_div_{nt}"] = {"n": int(tot16.numel()), "eq_true_div": int((g == f16(-t32 / np.float32(nt)).view(np.uint16)).sum()),
                            "eq_mul_recip": int((g == f16(-t32 * np.float32(1.0 / nt)).view(np.uint16)).sum())}

np.savez_compressed(OUT / "epig_probe_tables.npz", xl_gpu=xl_gpu.view(np.uint16), xl_cpu=xl_cpu.view(np.uint16), lg_gpu=lg_gpu.view(np.uint16))
print(json.dumps(res, indent=1), flush=True)


# ---------------------------------------------------------------------------------------------------- 2. kernels vs torch CUDA
def probs(gen_c, n, k, cl, spread=2.0):
    mean = torch.randn(n, cl, generator=gen_c) * spread
    var = torch.rand(n, cl, generator=gen_c) * 3 + 0.1
    eps = torch.randn(k, n, cl, generator=gen_c)
    return mean.to(dev), var.to(dev), eps.to(dev)


def stats(a, b):
    a = a.float().cpu().numpy()
    b = b.float().cpu().numpy()
    d = np.abs(a - b)
    return {"exact": float((d == 0).mean()), "max_abs": float(d.max()), "n": int(d.size)}


quick = "--quick" in sys.argv
cfgs = [dict(Np=300, Nt=200, K=100, Cl=10, chunk=512), dict(Np=1000, Nt=700, K=100, Cl=10, chunk=4096),
        dict(Np=257, Nt=129, K=64, Cl=65, chunk=4096), dict(Np=4096, Nt=2000, K=100, Cl=10, chunk=4096)]
if not quick:
    cfgs += [dict(Np=4096, Nt=10000, K=100, Cl=10, chunk=4096), dict(Np=1024, Nt=10000, K=100, Cl=65, chunk=4096)]
res["kernels"] = []
for cfg in cfgs:
    g = torch.Generator().manual_seed(cfg["Np"] + 13 * cfg["Cl"])
    mp, vp, ep = probs(g, cfg["Np"], cfg["K"], cfg["Cl"])
    mt, vt, et = probs(g, cfg["Nt"], cfg["K"], cfg["Cl"])
    # E0: kernel vs torch's own expression (vlm.py:116-123 then .to(float16))
    p16 = sample_probas_from_noise(mp, vp, ep)
    t16 = sample_probas_from_noise(mt, vt, et)
    p_ref = torch.softmax((ep * torch.sqrt(vp) + mp).permute(1, 0, 2), dim=2).half()
    r = {"cfg": cfg, "E0_sample": stats(p16, p_ref)}
    # E1: marginal entropy kernel vs torch on the same fp16 probabilities
    me = E.marginal_entropy_from_probs(p16)
    me_ref = T._entropy(p16.mean(dim=1))
    r["E1_marginal"] = stats(me, me_ref)
    # E2: joint entropy, kernel vs the reference sequence on this GPU
    hj = E.joint_entropy_from_probs(p16, t16, cfg["chunk"])
    n_t, k, cl = t16.shape
    pool = p16.permute(0, 2, 1)
    targ = t16.permute(1, 0, 2).reshape(k, n_t * cl)
    hj_ref = torch.zeros(pool.shape[0], device=dev)
    n_chunks = 0
    margin = []
    for lo in range(0, n_t * cl, cfg["chunk"]):
        joint = pool @ targ[:, lo:lo + cfg["chunk"]] / k
        xl = torch.xlogy(joint, joint)
        hj_ref += -torch.sum(xl, dim=(-2, -1)) / n_t
        if n_chunks < 2:  # how close do the fp32 chunk sums sit to an fp16 rounding boundary?
            s32 = xl.float().sum(dim=(-2, -1))
            s16 = s32.half().float()
            ulp = torch.from_numpy(np.spacing(np.abs(s16.cpu().numpy()).astype(np.float16)).astype(np.float32)).to(dev)
            margin.append(float(((0.5 - (s32 - s16).abs() / ulp).abs()).min()))
        n_chunks += 1
    r["E2_joint"] = stats(hj, hj_ref)
    r["E2_joint"]["n_chunks"] = n_chunks
    r["E2_joint"]["min_margin_to_boundary_ulps"] = margin
    dj = (hj - hj_ref).abs().cpu().numpy()
    nz = dj[dj > 0]
    r["E2_joint"]["diff_hist"] = {str(v): int(c) for v, c in zip(*np.unique(np.round(nz, 7), return_counts=True))} if nz.size else {}
    # final scores and top-k
    s = E.epig_from_probs_using_matmul(p16, t16, chunk_size=cfg["chunk"]).float()
    s_ref = T.epig_from_probs(p16, t16, chunk_size=cfg["chunk"]).float()
    r["scores"] = stats(s, s_ref)
    kk = min(50, cfg["Np"] // 2)
    a = set(torch.argsort(s, descending=True)[:kk].tolist())
    b = set(torch.argsort(s_ref, descending=True)[:kk].tolist())
    r["scores"]["topk_overlap"] = len(a & b) / kk
    r["scores"]["range"] = [float(s_ref.min()), float(s_ref.max())]
    r["scores"]["distinct_ref_values"] = int(torch.unique(s_ref).numel())
    res["kernels"].append(r)
    print(json.dumps(r), flush=True)
    del p16, t16, p_ref, pool, targ
    torch.cuda.empty_cache()

# E3 with the device generator: the product path vs the reference sequence (torch_port) under the same seeds
g = torch.Generator().manual_seed(9)
mp, vp, _ = probs(g, 5000, 1, 10)
mt, vt, _ = probs(g, 1500, 1, 10)
lp, lt = ProbabilisticLogits(mp, vp), ProbabilisticLogits(mt, vt)
s = E.epig_from_logits_using_matmul(lp, lt, seed=3, num_samples=100, chunk_size=4096)
s_ref = T.epig_from_logits(mp, vp, mt, vt, seed=3, num_samples=100, chunk_size=4096)
res["E3_from_logits"] = stats(s, s_ref)
a = set(torch.argsort(s, descending=True)[:50].tolist())
b = set(torch.argsort(s_ref, descending=True)[:50].tolist())
res["E3_from_logits"]["top50_overlap"] = len(a & b) / 50

(OUT / "epig_probe.json").write_text(json.dumps(res, indent=1))
print(json.dumps(res["E3_from_logits"]))
