"""ADVICE r1 (workspaces shared by streams): four host threads, each on its own CUDA stream, run the predictive, the InfoNCE GGN
and the SYRK concurrently through one model; results must equal the ones computed alone (scratch buffers are per stream)."""
import math
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_concurrent_threads_and_streams():
    import bench
    from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
    from bayesvlm_b200.hessians import compute_hessian_analytic_InfoNCE, syrk_accumulate
    from bayesvlm_b200.vlm import CLIP, EncoderResult

    g = torch.Generator().manual_seed(3)
    N, C, D, di, dt = 6000, 700, 512, 768, 512
    spd = lambda d, sc, lam: torch.linalg.inv(bench.surrogate_spd(g, d, sc).double() + math.sqrt(lam) * torch.eye(d, dtype=torch.float64)).float().cuda()
    m = CLIP(logit_scale=bench.LS, device="cuda")
    m.set_covariances(KFC(spd(di, 3e3, 600.0), spd(D, 20.0, 600.0)), KFC(spd(dt, 3e3, 200.0), spd(D, 20.0, 200.0)))
    txt = EncoderResult(torch.randn(C, D, generator=g).cuda(), torch.randn(C, dt, generator=g).cuda())
    imgs = [EncoderResult(torch.randn(N, D, generator=g).cuda(), torch.randn(N, di, generator=g).cuda()) for _ in range(4)]
    Xs = [torch.randn(3000, 256, generator=g).cuda() for _ in range(4)]
    with torch.no_grad():
        ref = [m(im, txt) for im in imgs]
        refH = [compute_hessian_analytic_InfoNCE(x[:900], x[900:], torch.tensor(bench.LS)) for x in Xs]
        refA = [syrk_accumulate(x) for x in Xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(4)]
    res, resH, resA = [None] * 4, [None] * 4, [None] * 4
    def work(i):
        with torch.no_grad(), torch.cuda.stream(streams[i]):
            for _ in range(20):
                res[i] = m(imgs[i], txt)
                resH[i] = compute_hessian_analytic_InfoNCE(Xs[i][:900], Xs[i][900:], torch.tensor(bench.LS))
                resA[i] = syrk_accumulate(Xs[i])
    ths = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    torch.cuda.synchronize()
    for i in range(4):
        assert torch.equal(res[i].mean, ref[i].mean) and torch.equal(res[i].var, ref[i].var)
        # (split-K / column-sum atomics: the order of the fp32 adds differs from run to run)
        assert float((resH[i] - refH[i]).abs().max() / refH[i].abs().max()) <= 1e-4
        assert float((resA[i] - refA[i]).abs().max() / refA[i].abs().max()) <= 1e-5


def test_target_cache_prepared_on_another_stream():
    """The cached target-side operands are prepared asynchronously on the stream of the first call; a second call on another
    stream, issued without any synchronisation in between, must wait for them."""
    import bench
    from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
    from bayesvlm_b200.vlm import CLIP, EncoderResult

    g = torch.Generator().manual_seed(4)
    N, C, D, di, dt = 3000, 1000, 768, 1024, 768
    spd = lambda d, sc, lam: torch.linalg.inv(bench.surrogate_spd(g, d, sc).double() + math.sqrt(lam) * torch.eye(d, dtype=torch.float64)).float().cuda()
    covs = (KFC(spd(di, 3e3, 600.0), spd(D, 20.0, 600.0)), KFC(spd(dt, 3e3, 200.0), spd(D, 20.0, 200.0)))
    img = EncoderResult(torch.randn(N, D, generator=g).cuda(), torch.randn(N, di, generator=g).cuda())
    txt = EncoderResult(torch.randn(C, D, generator=g).cuda(), torch.randn(C, dt, generator=g).cuda())
    m0 = CLIP(logit_scale=bench.LS, device="cuda")
    m0.set_covariances(*covs)
    with torch.no_grad():
        ref = m0(img, txt)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(5):
        m0._target_cache = None  # (drop the previous operands, then poison the freed blocks the next ones will be carved from)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        junk = torch.full((32 << 20,), float("nan"), device="cuda")
        del junk
        m = CLIP(logit_scale=bench.LS, device="cuda")
        m.set_covariances(*covs)
        m._sides()  # (factorisation etc.: synchronous)
        torch.cuda.synchronize()
        with torch.no_grad():
            with torch.cuda.stream(sa):
                torch.cuda._sleep(20_000_000)  # keep stream A busy so that the target preparation is still queued ...
                a = m(img, txt)
            with torch.cuda.stream(sb):       # ... when stream B asks for the cached operands
                b = m(img, txt)
        torch.cuda.synchronize()
        assert torch.equal(a.mean, ref.mean) and torch.equal(b.mean, ref.mean) and torch.equal(b.var, ref.var)
