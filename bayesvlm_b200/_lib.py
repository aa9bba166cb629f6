"""ctypes binding of libbvlm.so (C ABI declared in include/bvlm.h).

The library is built in-tree by :mod:`bayesvlm_b200.build`.  There is no CPU fallback: importing this module on a
machine without the shared library raises, and every compute entry point raises ``RuntimeError`` when the status
code is non-zero (no GPU, wrong architecture, bad arguments ...).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

import torch

_PKG_DIR = Path(__file__).resolve().parent
# BVLM_LIB: load another build of the same C ABI instead (the -DBVLM_DIAG ablation build, `python -m bayesvlm_b200.build --diag`)
LIB_PATH = Path(os.environ["BVLM_LIB"]).resolve() if os.environ.get("BVLM_LIB") else _PKG_DIR / "libbvlm.so"

PREC_X1 = 1
PREC_X2F8 = 2
PREC_X3 = 3


def _load() -> ctypes.CDLL:
    if not LIB_PATH.exists() and os.environ.get("BVLM_NO_AUTOBUILD", "0") != "1":
        from . import build as _build

        _build.build()
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -m bayesvlm_b200.build` (needs nvcc). "
            "bayesvlm_b200 has no CPU or PyTorch fallback for its kernels."
        )
    return ctypes.CDLL(str(LIB_PATH))


lib = _load()

_P = c_void_p
_I = c_int64

# name -> (restype, argtypes); mirrors include/bvlm.h one to one (tests/test_abi.py checks the header against this).
SIGNATURES = {
    "bvlm_version": (c_char_p, []),
    "bvlm_status_string": (c_char_p, [c_int]),
    "bvlm_device_check": (c_int, []),
    "bvlm_syrk_workspace_bytes": (c_size_t, [_I, _I, c_int, c_int]),
    "bvlm_syrk_f32acc": (c_int, [_P, _I, _I, _I, c_int, c_int, _P, _I, c_float, c_int, _P, c_size_t, _P]),
    "bvlm_ggn_workspace_bytes": (c_size_t, [_I, _I, _I, c_int]),
    "bvlm_ggn_infonce": (c_int, [_P, _I, _I, _P, _I, _I, _I, c_float, c_int, _P, _I, c_int, _P, c_size_t, _P]),
    "bvlm_ggn_siglip": (c_int, [_P, _I, _I, _P, _I, _I, _I, c_float, c_float, c_int, _P, _I, c_int, _P, c_size_t, _P]),
    "bvlm_padded_k": (c_int64, [_I]),
    "bvlm_padded_cols": (c_int64, [_I]),
    "bvlm_factor_prepare": (c_int, [_P, _I, _I, c_float, _P, _I, _P]),
    "bvlm_quadform_workspace_bytes": (c_size_t, [_I, _I, c_int]),
    "bvlm_quadform": (c_int, [_P, _I, _I, _I, c_int, _P, _I, _I, c_float, _P, _P, c_size_t, _P]),
    "bvlm_predictive_target_workspace_bytes": (c_size_t, [_I, _I, _I, c_int]),
    "bvlm_predictive_t8_cols": (c_int64, [_I]),
    "bvlm_predictive_target_prepare": (
        c_int,
        [_P, _I, _I, _I, _P, _I, _I, c_int, _P, _I, _I, c_float, _P, c_float, c_float, c_int, _P, _P, _P, _P, _P, c_size_t, _P],
    ),
    "bvlm_predictive_workspace_bytes": (c_size_t, [_I, _I, _I, c_int, c_int]),
    "bvlm_predictive": (
        c_int,
        [_P, _I, _I, _I, _P, _I, _I, c_int, _P, _I, _I, c_float, _P, c_float, c_float, _P, _P, _P, _P, _P, _I, c_int, _P, _P,
         _P, _I, _P, c_size_t, _P],
    ),
    "bvlm_probit_softmax": (c_int, [_P, _P, _I, _I, _I, _P, _P]),
    "bvlm_mc_softmax_accumulate": (c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "bvlm_epig_operand_k": (c_int, [_I]),
    "bvlm_epig_prepare_supported": (c_int, [_I, _I, c_int]),
    "bvlm_epig_prepare_from_noise": (c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "bvlm_epig_prepare_from_probs": (c_int, [_P, _I, _I, _I, _P, _P, _P]),
    "bvlm_epig_prepare_pair_from_noise": (c_int, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _P]),
    "bvlm_epig_joint_operands_workspace_bytes": (c_size_t, [_I, _I, _I, _I]),
    "bvlm_epig_joint_entropy_operands": (c_int, [_P, _I, _P, _I, _I, _I, _I, _P, _P, c_size_t, _P]),
    "bvlm_epig_sample_probs": (c_int, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "bvlm_epig_marginal_entropy_f16": (c_int, [_P, _I, _I, _I, _P, _P]),
    "bvlm_epig_joint_workspace_bytes": (c_size_t, [_I, _I, _I, _I]),
    "bvlm_epig_joint_entropy_f16": (c_int, [_P, _I, _P, _I, _I, _I, _I, _P, _P, c_size_t, _P]),
    "bvlm_gemm_tn_f32": (c_int, [_P, _I, _P, _I, _I, c_int, c_float, _P, _I, c_int, _P]),
    "bvlm_gemm_mn_f32": (c_int, [_P, _I, _I, _P, _I, _I, _I, c_int, c_float, _P, _I, _P]),
    "bvlm_convert_rows_16": (c_int, [_P, _I, _I, _I, c_int, _P, _I, _P]),
    "bvlm_launch_count": (c_int64, []),
    "bvlm_timing_enable": (c_int, [c_int]),
    "bvlm_timing_tag_count": (c_int, []),
    "bvlm_timing_tag_name": (c_char_p, [c_int]),
    "bvlm_timing_collect": (c_int, [_P, _P, c_int]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == ABI drift: fail at import, loudly
    _fn.restype = _res
    _fn.argtypes = _args


def status_string(rc: int) -> str:
    return lib.bvlm_status_string(int(rc)).decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"libbvlm: {what} failed with status {rc}: {status_string(rc)}")


def version() -> str:
    return lib.bvlm_version().decode()


def launch_count() -> int:
    return int(lib.bvlm_launch_count())


def timing_enable(on: bool) -> None:
    check(lib.bvlm_timing_enable(int(on)), "bvlm_timing_enable")


def timing_collect() -> dict:
    """{kernel tag: (launches, total milliseconds)} of the tensor-core launches recorded since the last collect."""
    n = int(lib.bvlm_timing_tag_count())
    launches = (c_int64 * n)()
    total = (ctypes.c_double * n)()
    check(lib.bvlm_timing_collect(ctypes.cast(launches, c_void_p), ctypes.cast(total, c_void_p), n), "bvlm_timing_collect")
    return {lib.bvlm_timing_tag_name(i).decode(): (int(launches[i]), float(total[i])) for i in range(n) if launches[i]}


# ------------------------------------------------------------------------------------------------------------------
# tensor plumbing
# ------------------------------------------------------------------------------------------------------------------
def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    """Kernel operands must already live on a CUDA device: CPU tensors raise (no CPU fallback)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}; bayesvlm_b200 kernels run on CUDA (sm_100a) only and have no CPU fallback"
        )
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def rowmajor(t: torch.Tensor) -> torch.Tensor:
    """Return a view/copy whose last dimension is contiguous (row pitch may exceed the row length)."""
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return t
    return t.contiguous()


def ptr(t) -> c_void_p:
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(device: torch.device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def run(device: torch.device, name: str, *args) -> None:
    """Call the entry point `name` with `device` as the CUDA current device (the library launches kernels, memsets and
    encodes tensor maps on the current device; the reference API takes a device argument instead) and raise on a
    non-zero status."""
    with torch.cuda.device(device):
        rc = getattr(lib, name)(*args)
    check(rc, name)


_WORKSPACES: dict = {}


def workspace(device: torch.device, nbytes: int, tag: str = "default") -> torch.Tensor:
    """Grow-only scratch buffer (uint8, 256-byte aligned by the caching allocator) per (device, stream, tag): kernels on
    different streams never share scratch, and a buffer is only ever reused in the order of the stream it was allocated on,
    so regrowing it (the old block returns to the allocator, which is stream-ordered) is safe."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            del _WORKSPACES[key]
            del buf
        with torch.cuda.device(device):
            buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


def release_workspaces() -> None:
    _WORKSPACES.clear()
