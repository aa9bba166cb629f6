"""Multi-GPU behaviour ON HARDWARE (skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
  * rank-sharded kfac_ggn(distributed=True) over NCCL == the single-process result;
  * every entry point works on a device that is NOT the current one (the library switches devices itself)."""
import math
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
LS = math.log(100.0)
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem(n_cb=5, ncls=2048, d=128, d_in=192, seed=3):
    g = torch.Generator().manual_seed(seed)
    n = n_cb * ncls + 100  # remainder is dropped
    z = torch.randn(n, d, generator=g)
    return (z + 1.5 * torch.randn(n, d, generator=g), torch.randn(n, d_in, generator=g), z + 1.5 * torch.randn(n, d, generator=g))


def _worker(rank, world, port, ncls, likelihood, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from bayesvlm_b200.hessians import kfac_ggn
        from bayesvlm_b200.vlm import CLIP, SIGLIP

        emb_s, act_s, emb_t = (t.cuda(rank) for t in _problem())
        vlm = (SIGLIP(logit_scale=4.765, logit_bias=-12.93, device=f"cuda:{rank}") if likelihood == "siglip"
               else CLIP(logit_scale=LS, device=f"cuda:{rank}"))
        A, B = kfac_ggn(vlm, ncls, 5, emb_s, act_s, emb_t, f"cuda:{rank}", likelihood, distributed=True)
        if rank == 0:
            out["A"], out["B"] = A.cpu().numpy(), B.numpy()
    finally:
        dist.destroy_process_group()


@needs2
@pytest.mark.parametrize("likelihood", ["info_nce", "siglip"])
def test_sharded_kfac_over_nccl_equals_single_process(likelihood):
    import torch.multiprocessing as mp

    from bayesvlm_b200.hessians import kfac_ggn
    from bayesvlm_b200.vlm import CLIP, SIGLIP

    ncls = 2048
    emb_s, act_s, emb_t = (t.cuda(0) for t in _problem())
    vlm = (SIGLIP(logit_scale=4.765, logit_bias=-12.93, device="cuda:0") if likelihood == "siglip" else CLIP(logit_scale=LS, device="cuda:0"))
    A1, B1 = kfac_ggn(vlm, ncls, 5, emb_s, act_s, emb_t, "cuda:0", likelihood)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), ncls, likelihood, out), nprocs=2, join=True)
        A2, B2 = out["A"], out["B"]
    ra = np.linalg.norm(A2 - A1.cpu().numpy()) / np.linalg.norm(A1.cpu().numpy())
    rb = np.linalg.norm(B2 - B1.numpy()) / np.linalg.norm(B1.numpy())
    assert ra <= 1e-4 and rb <= 1e-5, (ra, rb)


@needs2
def test_entry_points_on_a_non_current_device():
    """the reference API takes a device string; nothing may depend on torch's current device."""
    from bayesvlm_b200.epig import epig_from_probs_using_matmul
    from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
    from bayesvlm_b200.hessians import compute_hessian_analytic_InfoNCE, syrk_accumulate
    from bayesvlm_b200.vlm import CLIP, EncoderResult

    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(5)
    rn = lambda *s: torch.randn(*s, generator=g)
    spd = lambda d, sc: (lambda w: (w.T @ w) / math.sqrt(4 * d) * sc)(rn(4 * d, d))
    N, C, D, d_in = 300, 40, 64, 96
    tens = dict(ie=rn(N, D), ia=rn(N, d_in), te=rn(C, D), ta=rn(C, D), Ai=torch.linalg.inv(spd(d_in, 30.0)), Bi=torch.linalg.inv(spd(D, 2.0)),
                At=torch.linalg.inv(spd(D, 30.0)), Bt=torch.linalg.inv(spd(D, 2.0)))
    probs = torch.softmax(rn(50, 16, 5), -1).half()
    res = {}
    for dev in ("cuda:0", "cuda:1"):
        t = {k: v.to(dev) for k, v in tens.items()}
        m = CLIP(logit_scale=LS, device=dev)
        m.set_covariances(KFC(t["Ai"], t["Bi"]), KFC(t["At"], t["Bt"]))
        with torch.no_grad():
            out = m(EncoderResult(t["ie"], t["ia"]), EncoderResult(t["te"], t["ta"]))
        H = compute_hessian_analytic_InfoNCE(t["ie"], t["te"], torch.tensor(LS))
        A = syrk_accumulate(t["ia"])
        p = probs.to(dev)
        s = epig_from_probs_using_matmul(p, p[:30], chunk_size=256)
        assert out.mean.device == torch.device(dev) and H.device == torch.device(dev)
        res[dev] = [x.float().cpu() for x in (out.mean, out.var, H, A, s)]
    assert torch.cuda.current_device() == 0
    # (the GGN's column sums q are accumulated with atomics, and q / max q is then rounded into the fp16 operands of the final
    #  GEMM: the order of those fp32 adds moves H by up to ~2e-5 of its largest entry from run to run, on one device as well)
    for name, tol, a, b in zip(("mean", "var", "H", "A", "epig"), (1e-5, 1e-5, 1e-4, 1e-5, 1e-5), res["cuda:0"], res["cuda:1"]):
        assert torch.equal(a, b) or float((a - b).abs().max() / a.abs().max()) <= tol, name
