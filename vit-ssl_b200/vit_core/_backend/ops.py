"""Raw (non-autograd) Python wrappers over the C-ABI entry points.

Every function takes CUDA tensors, checks layout, and enqueues the kernel on the current stream.
Autograd glue lives in `functional.py`.
"""
from __future__ import annotations

import torch

from . import lib as _l

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU = 0, 1, 2, 3


def _p(t):
    return None if t is None else t.data_ptr()


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _l.VitsslError("vitssl_b200 ops need CUDA tensors (no CPU fallback)")


def gemm(a, b, *, a_mn=False, b_mn=False, epilogue=EPI_NONE, bias=None, aux=None, alpha=1.0,
         out_dtype=torch.bfloat16, split_k=0, dropout_p=0.0, seed=0, offset=0, out=None):
    """C = alpha * op(A) @ op(B) (+ epilogue). a, b are 2-D bf16 with unit inner stride.

    a_mn=False: a is [M,K]; True: a is [K,M].  b_mn=False: b is [N,K]; True: b is [K,N].
    """
    _l.ensure_device()
    _check_cuda(a, b, bias, aux, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, f"gemm: reduction dims differ ({K} vs {Kb})"
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    else:
        assert out.shape == (M, N) and out.stride(1) == 1
        out_dtype = out.dtype
    assert out_dtype in (torch.bfloat16, torch.float32)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    ld_aux = 0
    if aux is not None:
        assert aux.dtype == torch.bfloat16 and aux.shape == (M, N) and aux.stride(1) == 1
        ld_aux = aux.stride(0)
    _l.call(
        "vitssl_gemm_bf16", _p(a), _p(b), _p(out), M, N, K, a.stride(0), b.stride(0),
        out.stride(0), int(a_mn), int(b_mn), int(epilogue), _p(bias), _p(aux), ld_aux,
        float(alpha), int(out_dtype == torch.float32), int(split_k), float(dropout_p),
        int(seed), int(offset), _l.stream_ptr(),
    )
    return out
