"""The tcgen05/TMA GEMM engine in isolation (diagnostic C-ABI entry point) against torch on the SAME rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, B, fmt, split_k=1, alpha=1.0):
    from bayesvlm_b200 import _lib
    from bayesvlm_b200._lib import lib

    dev = A.device
    M, K = A.shape
    N = B.shape[0]
    kp = int(lib.bvlm_padded_k(K))
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    A16 = torch.empty((M, kp), dtype=dt, device=dev)
    B16 = torch.empty((N, kp), dtype=dt, device=dev)
    st = _lib.stream_ptr(dev)
    _lib.check(lib.bvlm_convert_rows_16(_lib.ptr(A), M, K, A.stride(0), fmt, _lib.ptr(A16), kp, st), "convert A")
    _lib.check(lib.bvlm_convert_rows_16(_lib.ptr(B), N, K, B.stride(0), fmt, _lib.ptr(B16), kp, st), "convert B")
    D = torch.full((M, N), float("nan"), dtype=torch.float32, device=dev)
    _lib.check(lib.bvlm_gemm_tn_f32(_lib.ptr(A16), M, _lib.ptr(B16), N, kp, fmt, alpha, _lib.ptr(D), D.stride(0), split_k, st),
               "gemm")
    torch.cuda.synchronize()
    assert torch.equal(A16[:, :K].float(), A.to(dt).float()) and (A16[:, K:] == 0).all()
    ref = (A16.double() @ B16.double().T) * alpha
    return D, ref


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("shape", [(128, 256, 64), (128, 256, 128), (300, 1000, 768), (50, 10, 512), (1000, 513, 2304),
                                   (4097, 77, 100)])
def test_gemm_matches_torch(shape, fmt):
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    D, ref = _gemm(A, B, fmt)
    err = (D.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert torch.isfinite(D).all()
    assert err <= 2e-5 * scale, (err, scale)


@pytest.mark.parametrize("split_k", [2, 5, 16])
def test_gemm_split_k_atomic_accumulate(split_k):
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(384, 4096, device="cuda", generator=g)
    B = torch.randn(512, 4096, device="cuda", generator=g)
    D, ref = _gemm(A, B, 1, split_k=split_k, alpha=0.5)
    assert (D.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


def test_gemm_identity_layout():
    """A = I picks rows of B^T: catches any swizzle / descriptor / TMEM lane mix-up exactly."""
    K = 256
    A = torch.eye(K, device="cuda")[:200]
    B = torch.arange(300 * K, device="cuda", dtype=torch.float32).reshape(300, K) % 1024
    D, ref = _gemm(A, B, 0)
    assert torch.equal(D.double(), ref)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(256, 256, 64), (300, 520, 192), (1000, 512, 1024), (77, 40, 128), (512, 512, 4096)])
def test_gemm_pair_engine_mn_major(shape, mode):
    """CTA-pair engine with MN-major operands (stored [K, M] / [K, N]): no transposing pass for X^T X shaped products."""
    from bayesvlm_b200 import _lib
    from bayesvlm_b200._lib import lib

    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + mode)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    B = torch.randn(N, K, device="cuda", generator=g).half()
    pad8 = lambda n: (n + 7) // 8 * 8
    if mode & 1:
        A_st = torch.zeros(K, pad8(M), device="cuda", dtype=torch.float16)
        A_st[:, :M] = A.T
    else:
        A_st = A.contiguous()
    if mode & 2:
        B_st = torch.zeros(K, pad8(N), device="cuda", dtype=torch.float16)
        B_st[:, :N] = B.T
    else:
        B_st = B.contiguous()
    D = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.bvlm_gemm_mn_f32(_lib.ptr(A_st), M, A_st.stride(0), _lib.ptr(B_st), N, B_st.stride(0), K, mode, 1.0,
                                    _lib.ptr(D), D.stride(0), _lib.stream_ptr(D.device)), "gemm_mn")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    assert torch.isfinite(D).all()
    assert (D.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
