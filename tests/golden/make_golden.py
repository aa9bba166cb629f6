"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run once in the build container (the reference tree is mounted read-only at /root/reference; it does not exist on the
GPU box, which is why the outputs are committed):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors for the Laplace hot path, so parity of `oracle/` (and through it of the
CUDA kernels) is pinned on these outputs: every array below is produced by calling the unmodified reference functions
(bayesvlm/hessians.py, bayesvlm/vlm.py, bayesvlm/epig.py, scripts/hessian_estimation.py::kfac_ggn) on seeded inputs with
CPU fp32 torch; inputs are stored next to the outputs.
"""
from __future__ import annotations

import importlib.util
import json
import math
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def load_reference():
    sys.path.insert(0, str(REF))
    import bayesvlm.epig as r_epig
    import bayesvlm.hessians as r_hess
    import bayesvlm.vlm as r_vlm

    # scripts/hessian_estimation.py imports bayesvlm.data.factory (needs pytorch_lightning, absent): stub it.
    stub = types.ModuleType("bayesvlm.data.factory")
    stub.DataModuleFactory = type("DataModuleFactory", (), {})
    sys.modules["bayesvlm.data.factory"] = stub
    spec = importlib.util.spec_from_file_location("ref_hessian_estimation", REF / "scripts" / "hessian_estimation.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return r_hess, r_vlm, r_epig, mod


def randn(gen, *shape):
    return torch.randn(*shape, generator=gen, dtype=torch.float32)


def spd(gen, d, scale):
    w = randn(gen, 4 * d, d)
    return (w.T @ w) / math.sqrt(4 * d) * scale


def paired(gen, n, d):
    z = randn(gen, n, d)
    return z + 1.5 * randn(gen, n, d), z + 1.5 * randn(gen, n, d)


def main():
    torch.set_num_threads(8)
    r_hess, r_vlm, r_epig, r_est = load_reference()
    ls = math.log(100.0)
    out = {}

    # ---- K2 / K3: analytic GGNs -------------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(101)
    X, Y = paired(g, 40, 24)
    X = X[:7]
    out["ggn_X"], out["ggn_Y"] = X.numpy(), Y.numpy()
    out["ggn_infonce_H"] = r_hess.compute_hessian_analytic_InfoNCE(X, Y, torch.tensor(ls)).numpy()
    sl_scale, sl_bias = 2.3, -3.0
    idx = torch.arange(7)
    out["ggn_siglip_H"] = r_hess.compute_hessian_analytic_SigLIP(X, idx, Y, torch.tensor(sl_scale), torch.tensor(sl_bias),
                                                                 chunk_size_j=16).numpy()
    out["ggn_siglip_params"] = np.array([sl_scale, sl_bias], np.float64)
    torch.set_default_dtype(torch.float64)  # hessians.py:39 uses the default dtype for eye()
    out["ggn_infonce_H64"] = r_hess.compute_hessian_analytic_InfoNCE(X.double(), Y.double(), torch.tensor(ls).double()).numpy()
    out["ggn_siglip_H64"] = r_hess.compute_hessian_analytic_SigLIP(X.double(), idx, Y.double(), torch.tensor(sl_scale).double(),
                                                                   torch.tensor(sl_bias).double()).numpy()
    torch.set_default_dtype(torch.float32)

    # ---- K0: the accumulation loop with its dropped remainders ------------------------------------------------------
    g = torch.Generator().manual_seed(202)
    n, ncls, bs, D, d_in = 138, 64, 5, 16, 24
    emb_s, emb_t = paired(g, n, D)
    act_s = randn(g, n, d_in)
    out["kfac_emb_s"], out["kfac_emb_t"], out["kfac_act_s"] = emb_s.numpy(), emb_t.numpy(), act_s.numpy()
    out["kfac_cfg"] = np.array([ncls, bs], np.int64)
    clip = r_vlm.CLIP(logit_scale=ls)
    A, B = r_est.kfac_ggn(clip, ncls, bs, emb_s, act_s, emb_t, "cpu", "info_nce")
    out["kfac_infonce_A"], out["kfac_infonce_B"] = A.numpy(), B.numpy()
    sig = r_vlm.SIGLIP(logit_scale=sl_scale, logit_bias=sl_bias)
    A, B = r_est.kfac_ggn(sig, ncls, bs, emb_s, act_s, emb_t, "cpu", "siglip", siglip_chunk_size_j=20)
    out["kfac_siglip_A"], out["kfac_siglip_B"] = A.numpy(), B.numpy()

    # ---- C1 + P2: covariances and the predictive (CLIP and SIGLIP flavours, surrogate factors) ----------------------
    g = torch.Generator().manual_seed(303)
    N, C, D, d_img, d_txt = 37, 11, 32, 40, 24
    info = {"n_img": 1.0, "n_txt": 1.0, "lambda_img": 600.0, "lambda_txt": 220.0}
    for tag, cls, bias in (("clip", r_vlm.CLIP, 0), ("siglip", r_vlm.SIGLIP, 1)):
        A_img, B_img = spd(g, d_img + bias, 3e3), spd(g, D, 20.0)
        A_txt, B_txt = spd(g, d_txt + bias, 3e3), spd(g, D, 20.0)
        cov_img, cov_txt = r_hess.compute_covariances(A_img, B_img, A_txt, B_txt, info)
        img = r_vlm.EncoderResult(embeds=randn(g, N, D), activations=randn(g, N, d_img))
        txt = r_vlm.EncoderResult(embeds=randn(g, C, D), activations=randn(g, C, d_txt))
        model = cls(logit_scale=ls) if bias == 0 else cls(logit_scale=sl_scale, logit_bias=sl_bias)
        model.set_covariances(cov_img, cov_txt)
        with torch.no_grad():
            pl = model(img, txt)
            pm = model(img, txt, map_estimate=True)
        pre = f"pred_{tag}_"
        for k_, v_ in dict(A_img=A_img, B_img=B_img, A_txt=A_txt, B_txt=B_txt, A_img_inv=cov_img.A_inv,
                           B_img_inv=cov_img.B_inv, A_txt_inv=cov_txt.A_inv, B_txt_inv=cov_txt.B_inv, img_emb=img.embeds,
                           img_act=img.activations, txt_emb=txt.embeds, txt_act=txt.activations, mean=pl.mean, var=pl.var,
                           map=pm.mean).items():
            out[pre + k_] = v_.numpy()
        kappa = 1 / torch.sqrt(1.0 + torch.pi / 8 * pl.var)
        out[pre + "probit"] = torch.softmax(kappa * pl.mean, dim=-1).numpy()             # zeroshot.py:119-120
        out[pre + "softmax0"] = pl.softmax(num_samples=0).numpy()                        # method quirk, vlm.py:74-78
    out["pred_info"] = np.array([info["n_img"], info["n_txt"], info["lambda_img"], info["lambda_txt"]], np.float64)

    # ---- config 1: shipped CLIP ViT-B-32 factors, 10k x 10 (rows subsampled for the fixture) ------------------------
    la_dir = REF / "hessians" / "hessian_CLIP-ViT-B-32-laion2B-s34B-b79K"
    cov_img, cov_txt, info32 = r_hess.load_covariances(str(la_dir), return_info=True)
    g = torch.Generator().manual_seed(1000)
    img = r_vlm.EncoderResult(embeds=randn(g, 10000, 512), activations=randn(g, 10000, 768))
    txt = r_vlm.EncoderResult(embeds=randn(g, 10, 512), activations=randn(g, 10, 512))
    model = r_vlm.CLIP(logit_scale=ls)
    model.set_covariances(cov_img, cov_txt)
    with torch.no_grad():
        pl = model(img, txt)
    b32 = {}
    for tag in ("A_img", "A_txt", "B_img", "B_txt"):
        b32[tag] = torch.load(la_dir / f"{tag}_analytic.pt", map_location="cpu").numpy()
    b32["info"] = np.array([info32["n_img"], info32["n_txt"], info32["lambda_img"], info32["lambda_txt"]], np.float64)
    b32["mean"], b32["var"] = pl.mean.numpy(), pl.var.numpy()
    b32["seed"] = np.array([1000], np.int64)
    np.savez_compressed(OUT / "b32_config1.npz", **b32)

    # ---- E0 / E1 / E2: EPIG on fp16 probabilities --------------------------------------------------------------------
    g = torch.Generator().manual_seed(404)
    Np, Nt, K, Cl, chunk = 45, 30, 16, 5, 64
    mean_p, var_p = randn(g, Np, Cl) * 2, torch.rand(Np, Cl, generator=g) * 3 + 0.1
    mean_t, var_t = randn(g, Nt, Cl) * 2, torch.rand(Nt, Cl, generator=g) * 3 + 0.1
    eps_p, eps_t = randn(g, K, Np, Cl), randn(g, K, Nt, Cl)
    probs_p = torch.softmax((eps_p * var_p.sqrt() + mean_p).permute(1, 0, 2), dim=2)       # vlm.py:121-123
    probs_t = torch.softmax((eps_t * var_t.sqrt() + mean_t).permute(1, 0, 2), dim=2)
    p16, t16 = probs_p.half(), probs_t.half()
    out.update(epig_mean_p=mean_p.numpy(), epig_var_p=var_p.numpy(), epig_eps_p=eps_p.numpy(), epig_mean_t=mean_t.numpy(),
               epig_var_t=var_t.numpy(), epig_eps_t=eps_t.numpy(), epig_probs_p=probs_p.numpy(), epig_probs_t=probs_t.numpy(),
               epig_cfg=np.array([chunk], np.int64))
    out["epig_marginal_p16"] = r_epig.marginal_entropy_from_probs(p16).numpy()
    out["epig_marginal_p32"] = r_epig.marginal_entropy_from_probs(probs_p).numpy()
    out["epig_scores_f16"] = r_epig.epig_from_probs_using_matmul(p16, t16, chunk_size=chunk).numpy()
    out["epig_scores_f32"] = r_epig.epig_from_probs_using_matmul(probs_p, probs_t, chunk_size=chunk).numpy()

    # ---- prior precision optimisation ---------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(505)
    proj = torch.nn.Linear(24, 16, bias=False)
    with torch.no_grad():
        proj.weight.copy_(randn(g, 16, 24) * 0.05)
    A, B = spd(g, 24, 30.0), spd(g, 16, 2.0)
    lam = r_hess.optimize_prior_precision(proj, A, B, lmbda_init=50.0, n=10.0, lr=1e-2, num_steps=60, device="cpu")
    out.update(prior_W=proj.weight.detach().numpy(), prior_A=A.numpy(), prior_B=B.numpy(),
               prior_cfg=np.array([50.0, 10.0, 1e-2, 60], np.float64), prior_lambda=np.array([lam.item()], np.float64))

    np.savez_compressed(OUT / "reference_small.npz", **out)
    meta = {"torch": torch.__version__, "reference": "MridulPandey17/BayesVLM @ /root/reference",
            "files": ["reference_small.npz", "b32_config1.npz", "knn_small.npz", "selection_small.npz"], "keys": sorted(out)}
    (OUT / "golden_meta.json").write_text(json.dumps(meta, indent=1))
    make_knn()
    make_selection()
    make_epig_online()
    make_shipped()
    print("wrote", OUT)


def shipped_problem(which: str):
    """Seeded inputs of the shipped-factor golden problems (also imported by the GPU parity test, which re-creates them).
    L-14 (BASELINE config 3 shape): 4096 images x 1000 classes, D = 768, d_img = 1024, d_txt = 768 -- the reference ships
    A_txt, B_img, B_txt for this model, A_img (1024^2) is a seeded surrogate.  SigLIP (config 5): 1024 images x 257 classes,
    D = 768, d_img = 3072 (+1 bias), d_txt = 768 (+1): shipped A_txt (769^2), B_img, B_txt; A_img (3073^2) surrogate."""
    if which == "l14":
        cfg = dict(name="hessian_CLIP-ViT-L-14-laion2B-s32B-b82K", N=4096, C=1000, D=768, d_img=1024, d_txt=768, bias=0, seed=3100,
                   logit_scale=math.log(100.0), logit_bias=0.0, rows=48)
    else:
        cfg = dict(name="hessian_siglip-base-patch16-256", N=1024, C=257, D=768, d_img=3072, d_txt=768, bias=1, seed=5100,
                   logit_scale=4.765, logit_bias=-12.93, rows=48)
    g = torch.Generator().manual_seed(cfg["seed"])
    t = dict(img_e=randn(g, cfg["N"], cfg["D"]), img_a=randn(g, cfg["N"], cfg["d_img"]), txt_e=randn(g, cfg["C"], cfg["D"]),
             txt_a=randn(g, cfg["C"], cfg["d_txt"]), A_img=spd(g, cfg["d_img"] + cfg["bias"], 3e3))
    t["rows"] = torch.randperm(cfg["N"], generator=g)[:cfg["rows"]].sort().values
    return cfg, t


def sym_from_lower(tri: np.ndarray, d: int) -> np.ndarray:
    """Symmetric [d, d] matrix from its packed lower triangle (np.tril_indices order)."""
    m = np.zeros((d, d), tri.dtype)
    il = np.tril_indices(d)
    m[il] = tri
    return m + np.tril(m, -1).T


def make_shipped():
    """Predictive on the factors the reference SHIPS for ViT-L-14 and SigLIP (hessians/...): the real three-decade spectra.
    The fixture stores the packed lower triangles (the shipped B factors are symmetric to 1.5e-7 relative; BOTH the reference
    run below and the GPU test use the matrices rebuilt from these triangles) + the reference's outputs on 48 seeded rows.
    `python make_golden.py shipped` regenerates only these files."""
    r_hess, r_vlm, _, _ = load_reference()
    for which in ("l14", "siglip"):
        cfg, t = shipped_problem(which)
        la = REF / "hessians" / cfg["name"]
        info = json.loads((la / "prior_precision_analytic.json").read_text())
        fx = {}
        facs = {}
        for tag in ("A_txt", "B_img", "B_txt"):
            full = torch.load(la / f"{tag}_analytic.pt", map_location="cpu").numpy()
            d = full.shape[0]
            fx[tag + "_tril"] = full[np.tril_indices(d)]
            facs[tag] = torch.from_numpy(sym_from_lower(fx[tag + "_tril"], d))
            fx[tag + "_asym"] = np.array([np.abs(full - full.T).max() / np.abs(full).max()], np.float64)
        cov_img, cov_txt = r_hess.compute_covariances(t["A_img"], facs["B_img"], facs["A_txt"], facs["B_txt"], info)
        if which == "l14":
            model = r_vlm.CLIP(logit_scale=cfg["logit_scale"])
        else:
            model = r_vlm.SIGLIP(logit_scale=cfg["logit_scale"], logit_bias=cfg["logit_bias"])
        model.set_covariances(cov_img, cov_txt)
        rows = t["rows"]
        with torch.no_grad():
            pl = model(r_vlm.EncoderResult(embeds=t["img_e"][rows], activations=t["img_a"][rows]),
                       r_vlm.EncoderResult(embeds=t["txt_e"], activations=t["txt_a"]))
        fx.update(mean=pl.mean.numpy(), var=pl.var.numpy(), rows=rows.numpy(),
                  info=np.array([info["n_img"], info["n_txt"], info["lambda_img"], info["lambda_txt"]], np.float64))
        np.savez_compressed(OUT / f"shipped_{which}.npz", **fx)
        print(which, "mean range", float(pl.mean.min()), float(pl.mean.max()), "var range", float(pl.var.min()), float(pl.var.max()))


def epig_online_problem(seed=21, D=32, d_in=40, n_cls=6, n_pool=600, n_targ=300):
    """Seeded inputs of the online-EPIG golden problem (also imported by the GPU parity test, which re-creates them)."""
    g = torch.Generator().manual_seed(seed)
    W = randn(g, D, d_in) / math.sqrt(d_in)
    pool_a, targ_a = randn(g, n_pool, d_in), randn(g, n_targ, d_in)
    label_e, label_a = randn(g, n_cls, D), randn(g, n_cls, D)
    ids = torch.randint(0, n_cls, (n_pool,), generator=g)
    A_img, A_txt, B_img, B_txt = spd(g, d_in, 3e3), spd(g, D, 3e3), spd(g, D, 20.0), spd(g, D, 20.0)
    info = {"n_img": 1.0, "n_txt": 1.0, "lambda_img": 600.0, "lambda_txt": 220.0}
    cfg = dict(budget=3, lr=1e-4, hessian_update_scale=10.0, num_samples=16, seed=0, pool_max_size=512, target_max_size=256,
               chunk_size=256, logit_scale=math.log(20.0))
    return dict(W=W, pool_a=pool_a, targ_a=targ_a, label_e=label_e, label_a=label_a, ids=ids, A_img=A_img, A_txt=A_txt,
                B_img=B_img, B_txt=B_txt, info=info, cfg=cfg)


def make_epig_online():
    """The reference's own select_epig_online (bayesvlm/epig.py:44-273) and fused-chunk EPIG scores on the CPU -> epig_online_small.npz.
    `python make_golden.py epig_online` regenerates only this file."""
    import contextlib
    import io

    r_hess, r_vlm, r_epig, _ = load_reference()
    pr = epig_online_problem()
    c = pr["cfg"]
    proj = torch.nn.Linear(pr["W"].shape[1], pr["W"].shape[0], bias=False)
    with torch.no_grad():
        proj.weight.copy_(pr["W"])
        pool = r_vlm.EncoderResult(embeds=proj(pr["pool_a"]), activations=pr["pool_a"])
        targ = r_vlm.EncoderResult(embeds=proj(pr["targ_a"]), activations=pr["targ_a"])
    labels = r_vlm.EncoderResult(embeds=pr["label_e"], activations=pr["label_a"])
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        idx, scores = r_epig.select_epig_online(
            label_features=labels, pool_features=pool, target_features=targ, pool_class_ids=pr["ids"], image_projection=proj,
            clip=r_vlm.CLIP(logit_scale=c["logit_scale"]), A_img=pr["A_img"], A_txt=pr["A_txt"], B_img=pr["B_img"], B_txt=pr["B_txt"],
            cov_info=dict(pr["info"]), budget=c["budget"], lr=c["lr"], hessian_update_scale=c["hessian_update_scale"],
            device=torch.device("cpu"), num_samples=c["num_samples"], seed=c["seed"], pool_max_size=c["pool_max_size"],
            target_max_size=c["target_max_size"], chunk_size=c["chunk_size"])
    out = dict(selected=np.array(idx, np.int64), scores=np.array(scores, np.float64))
    # fused-path shaped scores (chunk 256) of the reference's epig_from_probs_using_matmul on fp16 probabilities
    g = torch.Generator().manual_seed(707)
    Np, Nt, K, Cl, chunk = 96, 120, 32, 10, 256
    mp, vp = randn(g, Np, Cl) * 2, torch.rand(Np, Cl, generator=g) * 3 + 0.1
    mt, vt = randn(g, Nt, Cl) * 2, torch.rand(Nt, Cl, generator=g) * 3 + 0.1
    ep, et = randn(g, K, Np, Cl), randn(g, K, Nt, Cl)
    p16 = torch.softmax((ep * vp.sqrt() + mp).permute(1, 0, 2), dim=2).half()
    t16 = torch.softmax((et * vt.sqrt() + mt).permute(1, 0, 2), dim=2).half()
    out.update(fused_p16=p16.numpy(), fused_t16=t16.numpy(), fused_chunk=np.array([chunk], np.int64),
               fused_scores=r_epig.epig_from_probs_using_matmul(p16, t16, chunk_size=chunk).numpy(),
               fused_marginal=r_epig.marginal_entropy_from_probs(p16).numpy())
    np.savez_compressed(OUT / "epig_online_small.npz", **out)


def make_knn():
    """Support-set search (bayesvlm/knn.py) on a small clustered problem -> knn_small.npz.  `python make_golden.py knn`
    regenerates only this file."""
    import contextlib
    import io

    sys.path.insert(0, str(REF))
    import bayesvlm.knn as r_knn
    import bayesvlm.vlm as r_vlm
    from bayesvlm.hessians import KroneckerFactorizedCovariance

    g = torch.Generator().manual_seed(606)
    D, d_act, n_train, n_test, n_centres = 24, 32, 300, 40, 12
    centres, centres_a = randn(g, n_centres, D) * 2.0, randn(g, n_centres, d_act)
    lab_tr, lab_te = torch.randint(0, n_centres, (n_train,), generator=g), torch.randint(0, n_centres, (n_test,), generator=g)
    train = r_vlm.EncoderResult(embeds=centres[lab_tr] + 0.7 * randn(g, n_train, D),
                                activations=centres_a[lab_tr] + 0.5 * randn(g, n_train, d_act))
    test = r_vlm.EncoderResult(embeds=centres[lab_te] + 0.7 * randn(g, n_test, D),
                               activations=centres_a[lab_te] + 0.5 * randn(g, n_test, d_act))
    cov = KroneckerFactorizedCovariance(A_inv=spd(g, d_act, 2e-2), B_inv=spd(g, D, 0.4))
    indices_test = torch.randperm(n_test, generator=g)[:9]
    values_test = torch.rand(9, generator=g)
    out = dict(train_e=train.embeds.numpy(), train_a=train.activations.numpy(), test_e=test.embeds.numpy(),
               test_a=test.activations.numpy(), A_inv=cov.A_inv.numpy(), B_inv=cov.B_inv.numpy(),
               indices_test=indices_test.numpy(), values_test=values_test.numpy())
    for tag, fn, k_nearest, buf in (("cos", r_knn.find_similar_samples_cosine, 3, 10),
                                    ("wass", r_knn.find_similar_samples_wasserstein, 3, 10),
                                    ("cos_k5", r_knn.find_similar_samples_cosine, 5, 25)):
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints its progress
            res = fn(train, test, indices_test, values_test, k_nearest=k_nearest, source_covariance=cov, device="cpu",
                     buffersize=buf)
        width = max(len(v["indices"]) for v in res.values())
        idx = np.full((len(res), width), -1, np.int64)
        sim = np.full((len(res), width), np.nan, np.float32)
        for r, v in enumerate(res.values()):
            idx[r, :len(v["indices"])] = v["indices"]
            sim[r, :len(v["similarities"])] = v["similarities"]
        out.update({f"{tag}_keys": np.array(list(res.keys()), np.int64), f"{tag}_scores": np.array([v["score"] for v in res.values()]),
                    f"{tag}_indices": idx, f"{tag}_sims": sim, f"{tag}_cfg": np.array([k_nearest, buf], np.int64)})
        ex = r_knn.extract_test_train_indices(res)
        out[f"{tag}_extract_train"] = np.array(sorted(ex["train"]), np.int64)
    # the distance itself on explicit covariances (knn.py:6-20)
    c1, c2 = torch.rand(9, D, generator=g) + 0.05, torch.rand(n_train, D, generator=g) + 0.05
    out.update(w_cov1=c1.numpy(), w_cov2=c2.numpy(),
               w_dist=r_knn.wdist2(test.embeds[indices_test], train.embeds, c1, c2).numpy())
    np.savez_compressed(OUT / "knn_small.npz", **out)


def make_selection():
    """Acquisition scores / subset selection (bayesvlm/selection.py) on CPU tensors -> selection_small.npz.
    `python make_golden.py selection` regenerates only this file."""
    sys.path.insert(0, str(REF))
    import bayesvlm.selection as r_sel
    import bayesvlm.vlm as r_vlm

    g = torch.Generator().manual_seed(707)
    n, c = 60, 7
    mean, var = randn(g, n, c) * 3, torch.rand(n, c, generator=g) * 4 + 0.05
    class_ids = torch.randint(0, 4, (n,), generator=g)
    pl = r_vlm.ProbabilisticLogits(mean=mean, var=var)
    out = dict(mean=mean.numpy(), var=var.numpy(), class_ids=class_ids.numpy())
    for ev in ("map_alea", "comb", "comb_covar", "exp_alea"):
        torch.manual_seed(11)  # exp_alea is not seeded by the reference: seed the global generator around the call
        out[f"entropy_{ev}"] = r_sel._entropy(mean, var, ev, num_samples=25, seed=3).numpy()
    out["score_var"] = r_sel.complexity_score(pl, "var").numpy()  # 2-D var: the diagonal of the N x C matrix -> a scalar
    torch.manual_seed(12)
    out["score_map_mi"] = r_sel.complexity_score(pl, "map_mutual_info", seed=5).numpy()
    torch.manual_seed(13)
    out["score_exp_mi"] = r_sel.complexity_score(pl, "exp_mutual_info", seed=5).numpy()
    idx, val = r_sel.select_topk(pl, 9, "entropy", "map_alea", ignore_percentage=0.1, return_values=True)
    out["topk_entropy_idx"], out["topk_entropy_val"] = idx.numpy(), val.numpy()
    out["topk_cb_var"] = r_sel.select_topk_classbalanced(pl, class_ids, 10, "var").numpy()
    out["topk_cb_entropy"] = r_sel.select_topk_classbalanced(pl, class_ids, 10, "entropy", "map_alea").numpy()
    out["topk_rand"] = r_sel.select_topk_randomized(pl, 8, 1.5, "entropy", "comb", seed=4).numpy()
    out["random_cb"] = r_sel.select_random_classbalanced(var, class_ids, 10, seed=6).numpy()
    out["random"] = r_sel.select_random(pl, 12, seed=8).numpy()
    cov = torch.stack([torch.diag(v) + 0.01 for v in var[:10]])
    out["score_logdet"] = r_sel.complexity_score(r_vlm.ProbabilisticLogits(mean=mean[:10], var=cov), "logdet").numpy()
    out["score_var3d"] = r_sel.complexity_score(r_vlm.ProbabilisticLogits(mean=mean[:10], var=cov), "var").numpy()
    out["topk_var3d"] = r_sel.select_topk(r_vlm.ProbabilisticLogits(mean=mean[:10], var=cov), 4, "var").numpy()
    np.savez_compressed(OUT / "selection_small.npz", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "knn":
        make_knn()
    elif len(sys.argv) > 1 and sys.argv[1] == "selection":
        make_selection()
    elif len(sys.argv) > 1 and sys.argv[1] == "epig_online":
        make_epig_online()
    elif len(sys.argv) > 1 and sys.argv[1] == "shipped":
        make_shipped()
    else:
        main()
