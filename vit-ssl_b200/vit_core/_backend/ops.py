"""Raw (non-autograd) Python wrappers over the C-ABI entry points.

Every function takes CUDA tensors, checks layout, and enqueues the kernel on the current stream.
Autograd glue lives in `functional.py`.
"""
from __future__ import annotations

import torch

from . import lib as _l

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_BIAS_GELU_D, EPI_MUL = 0, 1, 2, 3, 4, 5

# Optional per-launch timing used by bench.py's instrumented pass: when PROFILE is a list, the
# wrapped kernels are bracketed by CUDA events on the launching stream and
# (kind, start_event, end_event, algorithmic_work) tuples are appended (work = FLOPs or bytes).
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(kind, e0, work):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROFILE.append((kind, e0, e1, work))


def _p(t):
    """Device pointer of a tensor argument (None -> NULL). A host tensor never reaches a kernel:
    there is no CPU fallback, and a host pointer would fault the whole CUDA context."""
    if t is None:
        return None
    if not t.is_cuda:
        raise _l.VitsslError("vitssl_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


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
    e0 = _prof_begin()
    _l.call(
        "vitssl_gemm_bf16", _p(a), _p(b), _p(out), M, N, K, a.stride(0), b.stride(0),
        out.stride(0), int(a_mn), int(b_mn), int(epilogue), _p(bias), _p(aux), ld_aux,
        float(alpha), int(out_dtype == torch.float32), int(split_k), float(dropout_p),
        int(seed), int(offset), _l.stream_ptr(),
    )
    _prof_end(f"gemm|{M}x{N}x{K}|a_mn={int(a_mn)} b_mn={int(b_mn)} epi={int(epilogue)}", e0, 2.0 * M * N * K)
    return out


def gemm_rowsum(a, b, *, a_mn=False, b_mn=False, alpha=1.0, split_k=0, out=None, rowsum=None):
    """fp32 C = alpha * op(A) @ op(B) and, from the same kernel, rowsum[m] += alpha * sum_k op(A)[m, k]
    (with a = dY stored [tokens, features] and a_mn=True: weight gradient + bias gradient of a Linear).
    `out` / `rowsum` are accumulated into when given with split_k=-2 (caller-zeroed); otherwise fresh
    zeroed buffers are returned."""
    _l.ensure_device()
    _check_cuda(a, b, out, rowsum)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    K, M = a.shape if a_mn else a.shape[::-1]
    N = b.shape[1] if b_mn else b.shape[0]
    if out is None:
        out = torch.zeros((M, N), device=a.device, dtype=torch.float32)
    if rowsum is None:
        rowsum = torch.zeros((M,), device=a.device, dtype=torch.float32)
    assert out.dtype == torch.float32 and rowsum.dtype == torch.float32 and rowsum.is_contiguous()
    e0 = _prof_begin()
    _l.call("vitssl_gemm_bf16_rowsum", _p(a), _p(b), _p(out), _p(rowsum), M, N, K, a.stride(0), b.stride(0),
            out.stride(0), int(a_mn), int(b_mn), float(alpha), int(split_k), _l.stream_ptr())
    _prof_end(f"gemm|{M}x{N}x{K}|a_mn={int(a_mn)} b_mn={int(b_mn)} epi=0", e0, 2.0 * M * N * K)
    return out, rowsum


def add_layernorm_fwd(x, branch, gamma, beta, *, eps=1e-5, dropout_p=0.0, seed=0, offset=0):
    """x: fp32 [..., D] rows (last dim contiguous, uniform row pitch); branch: bf16 dense or None.

    Returns (x_out fp32 dense | x itself when branch is None, y bf16 | None, mean, rstd).
    """
    _l.ensure_device()
    _check_cuda(x, branch, gamma, beta)
    D = x.shape[-1]
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    x2 = x.reshape(-1, D) if x.is_contiguous() else x
    assert x2.dim() == 2, "strided LayerNorm input must be 2-D"
    rows, ldx = x2.shape[0], x2.stride(0)
    x_out = None
    if branch is not None:
        assert branch.dtype == torch.bfloat16 and branch.is_contiguous() and branch.numel() == rows * D
        x_out = torch.empty((rows, D), device=x.device, dtype=torch.float32)
    y = mean = rstd = None
    if gamma is not None:
        y = torch.empty((rows, D), device=x.device, dtype=torch.bfloat16)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    e0 = _prof_begin()
    _l.call("vitssl_add_layernorm_fwd", _p(x2), ldx, _p(branch), _p(x_out), _p(gamma), _p(beta),
            _p(y), _p(mean), _p(rstd), rows, D, float(eps), float(dropout_p), int(seed), int(offset),
            _l.stream_ptr())
    _prof_end("add_layernorm", e0, rows * D * (4 + (6 if branch is not None else 0) + (2 if gamma is not None else 0)))
    return (x_out if x_out is not None else x2), y, mean, rstd


def add_layernorm_bwd(dy, x, mean, rstd, gamma, dres, *, want_dx=True, want_dbranch=False,
                      dropout_p=0.0, seed=0, offset=0, dx_out=None):
    """Returns (dx fp32 | None, dbranch bf16 | None, dgamma | None, dbeta | None)."""
    _l.ensure_device()
    ref = dy if dy is not None else dres
    D = ref.shape[-1]
    rows = ref.numel() // D if ref.is_contiguous() else ref.shape[0]
    dev = ref.device
    ldx = 0
    if dy is not None:
        assert dy.dtype == torch.bfloat16 and dy.is_contiguous()
        assert x.dtype == torch.float32 and x.stride(-1) == 1
        ldx = x.stride(-2) if x.dim() >= 2 else D
    ld_dres = 0
    if dres is not None:
        assert dres.dtype == torch.float32 and dres.stride(-1) == 1
        ld_dres = dres.stride(-2) if dres.dim() >= 2 else D
    dx = None
    ld_dx = 0
    if dx_out is not None:
        dx = dx_out
        ld_dx = dx.stride(-2)
    elif want_dx:
        dx = torch.empty((rows, D), device=dev, dtype=torch.float32)
        ld_dx = D
    dbranch = torch.empty((rows, D), device=dev, dtype=torch.bfloat16) if want_dbranch else None
    dgamma = dbeta = None
    if dy is not None:
        dgamma = torch.empty(D, device=dev, dtype=torch.float32)
        dbeta = torch.empty(D, device=dev, dtype=torch.float32)
    e0 = _prof_begin()
    _l.call("vitssl_add_layernorm_bwd", _p(dy), _p(x), ldx, _p(mean), _p(rstd), _p(gamma),
            _p(dres), ld_dres, _p(dx), ld_dx, _p(dbranch), _p(dgamma), _p(dbeta), rows, D,
            float(dropout_p), int(seed), int(offset), _l.stream_ptr())
    _prof_end("add_layernorm", e0, rows * D * ((6 if dy is not None else 0) + (4 if dres is not None else 0)
                                               + (4 if dx is not None else 0) + (2 if dbranch is not None else 0)))
    return dx, dbranch, dgamma, dbeta


def attention_supported(Sq, Sk, d):
    return bool(_l.lib().vitssl_attention_supported(Sq, Sk, d))


def attention_fwd(q, k, v, H, scale, want_lse=True):
    """q: bf16 [B,Sq,H*64] view (last dim contiguous, token pitch uniform); k,v likewise.
    Returns (out bf16 [B,Sq,H*64], lse | None, out_lo | None): out_lo = bf16(O - bf16(O)) is produced with
    the LSE (training) and goes to `attention_bwd`, whose delta = rowsum(O * dO) needs O to 16 bits."""
    _l.ensure_device()
    B, Sq, _ = q.shape
    Sk = k.shape[1]
    for t_, S_ in ((q, Sq), (k, Sk), (v, Sk)):
        assert t_.dtype == torch.bfloat16 and t_.stride(2) == 1 and t_.stride(0) == S_ * t_.stride(1)
    out = torch.empty((B, Sq, H * 64), device=q.device, dtype=torch.bfloat16)
    lse = torch.empty((B, H, Sq), device=q.device, dtype=torch.float32) if want_lse else None
    out_lo = torch.empty_like(out) if want_lse else None
    _l.call("vitssl_attention_fwd", _p(q), _p(k), _p(v), q.stride(1), k.stride(1), v.stride(1),
            _p(out), _p(out_lo), out.stride(1), _p(lse), B, H, Sq, Sk, float(scale), _l.stream_ptr())
    return out, lse, out_lo


def attention_bwd(q, k, v, out, d_out, lse, H, scale, dq, dk, dv, out_lo=None):
    """dq/dk/dv: preallocated bf16 [B,S,H*64] views (may alias slices of one [B,S,3D] buffer)."""
    _l.ensure_device()
    B, Sq, _ = q.shape
    Sk = k.shape[1]
    assert out.is_contiguous() and d_out.is_contiguous() and d_out.dtype == torch.bfloat16
    assert out_lo is None or (out_lo.is_contiguous() and out_lo.shape == out.shape and out_lo.dtype == torch.bfloat16)
    delta = torch.empty((B, H, Sq), device=q.device, dtype=torch.float32)
    _l.call("vitssl_attention_bwd", _p(q), _p(k), _p(v), q.stride(1), k.stride(1), v.stride(1),
            _p(out), _p(out_lo), _p(d_out), out.stride(1), _p(lse), _p(delta), _p(dq), dq.stride(1), _p(dk),
            dk.stride(1), _p(dv), dv.stride(1), B, H, Sq, Sk, float(scale), _l.stream_ptr())


def _strides_bhsd(t):
    # t: [B,H,S,d] view with unit stride on d
    assert t.stride(3) == 1
    return [t.stride(0), t.stride(1), t.stride(2)]


def attention_generic_fwd(q, k, v, scale, want_probs=False, want_lse=True):
    """q,k,v: bf16 [B,H,S,d] views (any strides, unit stride on d). Returns (out [B,H,Sq,d], probs, lse)."""
    import ctypes
    _l.ensure_device()
    B, H, Sq, d = q.shape
    Sk = k.shape[2]
    out = torch.empty((B, H, Sq, d), device=q.device, dtype=torch.bfloat16)
    probs = torch.empty((B, H, Sq, Sk), device=q.device, dtype=torch.float32) if want_probs else None
    lse = torch.empty((B, H, Sq), device=q.device, dtype=torch.float32) if want_lse else None
    st = _strides_bhsd(q) + _strides_bhsd(k) + _strides_bhsd(v) + _strides_bhsd(out)
    arr = (ctypes.c_int64 * 12)(*st)
    _l.call("vitssl_attention_generic_fwd", _p(q), _p(k), _p(v), ctypes.addressof(arr), _p(out),
            _p(probs), _p(lse), B, H, Sq, Sk, d, float(scale), _l.stream_ptr())
    return out, probs, lse


def attention_generic_bwd(q, k, v, out, d_out, lse, scale):
    """Returns (dq [B,H,Sq,d] bf16, dk, dv as fp32 [B,Sk,H,d])."""
    import ctypes
    _l.ensure_device()
    B, H, Sq, d = q.shape
    Sk = k.shape[2]
    q = q.contiguous(); out = out.contiguous(); d_out = d_out.contiguous()
    dq = torch.empty_like(q)
    dk = torch.zeros((B, Sk, H, d), device=q.device, dtype=torch.float32)
    dv = torch.zeros((B, Sk, H, d), device=q.device, dtype=torch.float32)
    st = _strides_bhsd(q) + _strides_bhsd(k) + _strides_bhsd(v) + _strides_bhsd(out)
    arr = (ctypes.c_int64 * 12)(*st)
    _l.call("vitssl_attention_generic_bwd", _p(q), _p(k), _p(v), ctypes.addressof(arr), _p(out),
            _p(d_out), _p(lse), _p(dq), _p(dk), _p(dv), B, H, Sq, Sk, d, float(scale),
            _l.stream_ptr())
    return dq, dk, dv


# ----------------------------------------------------------------------------------------
# multi-tensor helpers
# ----------------------------------------------------------------------------------------
def _ptr_array(tensors):
    import ctypes
    return (ctypes.c_void_p * len(tensors))(*[_p(t) for t in tensors])


def _numel_array(tensors):
    import ctypes
    return (ctypes.c_int64 * len(tensors))(*[t.numel() for t in tensors])


def multi_cast_bf16(srcs, dsts):
    """fp32 -> bf16 for lists of contiguous tensors (dst may be slices of a larger buffer)."""
    import ctypes
    if not srcs:
        return
    _l.ensure_device()
    for s, d in zip(srcs, dsts):
        assert s.dtype == torch.float32 and d.dtype == torch.bfloat16
        assert s.is_contiguous() and d.is_contiguous() and s.numel() == d.numel()
    a, b, n = _ptr_array(srcs), _ptr_array(dsts), _numel_array(srcs)
    _l.call("vitssl_multi_cast_bf16", ctypes.addressof(a), ctypes.addressof(b), ctypes.addressof(n),
            len(srcs), _l.stream_ptr())


def cast_bf16(x):
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    multi_cast_bf16([x.contiguous()], [out])
    return out


def multi_ema(teacher, student, momentum):
    import ctypes
    if not teacher:
        return
    _l.ensure_device()
    for t, s in zip(teacher, student):
        assert t.dtype == torch.float32 and s.dtype == torch.float32
        assert t.is_contiguous() and s.is_contiguous() and t.numel() == s.numel()
    a, b, n = _ptr_array(teacher), _ptr_array(student), _numel_array(teacher)
    _l.call("vitssl_multi_ema", ctypes.addressof(a), ctypes.addressof(b), ctypes.addressof(n),
            len(teacher), float(momentum), _l.stream_ptr())


def multi_ema_shadow(teacher, student, shadows, momentum):
    """`multi_ema` that also writes each teacher tensor's bf16 shadow (shadows[i] or None)."""
    import ctypes
    if not teacher:
        return
    _l.ensure_device()
    for t, s_, sh in zip(teacher, student, shadows):
        assert t.dtype == torch.float32 and s_.dtype == torch.float32
        assert t.is_contiguous() and s_.is_contiguous() and t.numel() == s_.numel()
        assert sh is None or (sh.dtype == torch.bfloat16 and sh.is_contiguous() and sh.numel() == t.numel())
    a, b, c, n = _ptr_array(teacher), _ptr_array(student), _ptr_array(shadows), _numel_array(teacher)
    _l.call("vitssl_multi_ema_shadow", ctypes.addressof(a), ctypes.addressof(b), ctypes.addressof(c),
            ctypes.addressof(n), len(teacher), float(momentum), _l.stream_ptr())


def colsum_bf16(x):
    """x: bf16 [rows, cols] (unit inner stride) -> fp32 [cols]."""
    _l.ensure_device()
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    out = torch.empty(x.shape[1], device=x.device, dtype=torch.float32)
    _l.call("vitssl_colsum_bf16", _p(x), x.stride(0), x.shape[0], x.shape[1], _p(out), _l.stream_ptr())
    return out


def im2col_bf16(img, p, out=None):
    """img: fp32 [B,C,H,W], or raw uint8 image bytes (value = byte / 255, torchvision ToTensor).
    `out`: optional contiguous bf16 [B*N, C*p*p] destination (a slice of a larger patch buffer: several
    views are packed into one GEMM operand without a concatenated copy of the images)."""
    _l.ensure_device()
    assert img.dtype in (torch.float32, torch.uint8) and img.is_contiguous() and img.dim() == 4
    B, C, H, W = img.shape
    rows, cols = B * (H // p) * (W // p), C * p * p
    if out is None:
        out = torch.empty((rows, cols), device=img.device, dtype=torch.bfloat16)
    assert out.dtype == torch.bfloat16 and out.shape == (rows, cols) and out.is_contiguous()
    name = "vitssl_im2col_bf16" if img.dtype == torch.float32 else "vitssl_im2col_u8_bf16"
    _l.call(name, _p(img), _p(out), B, C, H, W, p, _l.stream_ptr())
    return out


def mean_tokens(x):
    """fp32 [B,S,D] -> [B,D] mean over tokens (ssl/simmim/model.py:91-93)."""
    _l.ensure_device()
    assert x.dtype == torch.float32 and x.dim() == 3 and x.is_contiguous()
    B, S, D = x.shape
    out = torch.empty((B, D), device=x.device, dtype=torch.float32)
    _l.call("vitssl_mean_tokens_f32", _p(x), _p(out), B, S, D, _l.stream_ptr())
    return out


def knn_cosine(val, train, train_labels, k, num_classes, want_neighbors=False):
    """Cosine k-NN vote on the GPU; val [Nv,D], train [Nt,D] fp32, train_labels int [Nt] in [0, num_classes).
    Returns (pred int32 [Nv], neighbors int32 [Nv,k] | None)."""
    _l.ensure_device()
    val = val.float().contiguous()
    train = train.float().contiguous()
    lab = train_labels.to(device=val.device, dtype=torch.int32).contiguous()
    Nv, D = val.shape
    Nt = train.shape[0]
    assert train.shape[1] == D and lab.numel() == Nt
    val_n, train_n = torch.empty_like(val), torch.empty_like(train)
    sims = torch.empty((Nv, Nt), device=val.device, dtype=torch.float32)
    pred = torch.empty((Nv,), device=val.device, dtype=torch.int32)
    nbr = torch.empty((Nv, k), device=val.device, dtype=torch.int32) if want_neighbors else None
    _l.call("vitssl_knn_cosine", _p(val), _p(train), _p(lab), _p(val_n), _p(train_n), _p(sims), _p(pred), _p(nbr),
            Nv, Nt, D, int(k), int(num_classes), _l.stream_ptr())
    return pred, nbr


def gather_patches_f32(img, rows_idx, p):
    _l.ensure_device()
    assert img.dtype in (torch.float32, torch.uint8) and img.is_contiguous() and rows_idx.dtype == torch.int32
    B, C, H, W = img.shape
    out = torch.empty((rows_idx.numel(), C * p * p), device=img.device, dtype=torch.float32)
    name = "vitssl_gather_patches_f32" if img.dtype == torch.float32 else "vitssl_gather_patches_u8_f32"
    _l.call(name, _p(img), _p(rows_idx), _p(out), rows_idx.numel(), C, H, W, p, _l.stream_ptr())
    return out


def interp_rows_fwd(src, idx, w):
    """dst[i,:] = sum_t w[i,t] * src[idx[i,t],:]; src fp32 [n_in, D], idx int32 / w fp32 [n_out, taps]."""
    _l.ensure_device()
    assert src.dtype == torch.float32 and src.is_contiguous() and src.dim() == 2
    assert idx.dtype == torch.int32 and w.dtype == torch.float32 and idx.shape == w.shape and idx.is_contiguous() and w.is_contiguous()
    n_out, taps = idx.shape
    dst = torch.empty((n_out, src.shape[1]), device=src.device, dtype=torch.float32)
    _l.call("vitssl_interp_rows_fwd", _p(src), _p(idx), _p(w), _p(dst), n_out, src.shape[1], taps, _l.stream_ptr())
    return dst


def interp_rows_bwd(ddst, idx, w, n_in):
    _l.ensure_device()
    assert ddst.dtype == torch.float32 and ddst.is_contiguous()
    n_out, taps = idx.shape
    dsrc = torch.empty((n_in, ddst.shape[1]), device=ddst.device, dtype=torch.float32)
    _l.call("vitssl_interp_rows_bwd", _p(ddst), _p(idx), _p(w), _p(dsrc), n_in, n_out, ddst.shape[1], taps,
            _l.stream_ptr())
    return dsrc


RANDPERM_MAX_N = 1024


def randperm_offset_per_call(n: int) -> int:
    """Philox offset one `torch.randperm(n, device=cuda)` call consumes (Randperm.cu)."""
    return int(_l.lib().vitssl_randperm_offset_per_call(int(n)))


def simmim_mask(B, N, n_keep, device, want_perm=True):
    """B sequential `torch.randperm(N, device=cuda)[:n_keep]` draws replayed bit-exactly in one
    launch from the device's default generator, which is advanced exactly as those calls would
    (ssl/simmim/masking.py:22-25). Returns (perm int64 [B,n_keep] | None, bool_mask [B,N],
    rows int32 [B*n_keep], inv int32 [B*N])."""
    _l.ensure_device()
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[index]
    seed, offset = int(gen.initial_seed()), int(gen.get_offset())
    gen.set_offset(offset + B * randperm_offset_per_call(N))
    perm = torch.empty((B, n_keep), device=dev, dtype=torch.int64) if want_perm else None
    bool_mask = torch.empty((B, N), device=dev, dtype=torch.bool)
    rows = torch.empty((B * n_keep,), device=dev, dtype=torch.int32)
    inv = torch.empty((B * N,), device=dev, dtype=torch.int32)
    _l.call("vitssl_simmim_mask", _p(perm), _p(bool_mask), _p(rows), _p(inv), B, N, n_keep,
            seed & 0xFFFFFFFFFFFFFFFF, offset, _l.stream_ptr())
    return perm, bool_mask, rows, inv


def embed_tokens_fwd(proj, cls, pos, mask_u8, mask_token, B, N, D):
    _l.ensure_device()
    S = N + (1 if cls is not None else 0)
    x = torch.empty((B, S, D), device=proj.device, dtype=torch.float32)
    _l.call("vitssl_embed_tokens_fwd", _p(proj), _p(cls), _p(pos), _p(mask_u8), _p(mask_token), _p(x),
            B, N, D, _l.stream_ptr())
    return x


def embed_tokens_bwd(dx, mask_u8, B, N, D, has_cls, want_dmask_token):
    _l.ensure_device()
    S = N + (1 if has_cls else 0)
    assert dx.dtype == torch.float32 and dx.shape == (B, S, D) and dx.stride(2) == 1
    dproj = torch.empty((B * N, D), device=dx.device, dtype=torch.bfloat16)
    dpos = torch.empty((S, D), device=dx.device, dtype=torch.float32)
    dmt = torch.empty(D, device=dx.device, dtype=torch.float32) if want_dmask_token else None
    _l.call("vitssl_embed_tokens_bwd", _p(dx), dx.stride(0), dx.stride(1), _p(mask_u8), _p(dproj),
            _p(dpos), _p(dmt), B, N, D, int(has_cls), _l.stream_ptr())
    return dproj, dpos, dmt


def gather_rows_bf16(x2d, idx):
    _l.ensure_device()
    assert x2d.dtype == torch.float32 and x2d.dim() == 2 and x2d.stride(1) == 1 and idx.dtype == torch.int32
    out = torch.empty((idx.numel(), x2d.shape[1]), device=x2d.device, dtype=torch.bfloat16)
    _l.call("vitssl_gather_rows_bf16", _p(x2d), x2d.stride(0), _p(idx), _p(out), idx.numel(), x2d.shape[1],
            _l.stream_ptr())
    return out


def scatter_rows_f32(dy, inv_idx, rows):
    _l.ensure_device()
    assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and inv_idx.dtype == torch.int32
    D = dy.shape[1]
    dx = torch.empty((rows, D), device=dy.device, dtype=torch.float32)
    _l.call("vitssl_scatter_rows_f32", _p(dy), _p(inv_idx), _p(dx), rows, D, _l.stream_ptr())
    return dx


def l1_loss_fwd(pred, target, want_sign=True):
    _l.ensure_device()
    assert pred.dtype == torch.bfloat16 and target.dtype == torch.float32
    assert pred.is_contiguous() and target.is_contiguous() and pred.numel() == target.numel()
    sign = torch.empty_like(pred) if want_sign else None
    loss = torch.empty((), device=pred.device, dtype=torch.float32)
    _l.call("vitssl_l1_loss_fwd", _p(pred), _p(target), _p(sign), _p(loss), pred.numel(), _l.stream_ptr())
    return loss, sign


def l1_loss_bwd(sign, grad_out):
    """d(pred) = sign * grad_out / n, grad_out a device scalar (any float dtype)."""
    _l.ensure_device()
    go = grad_out.reshape(()).to(torch.float32).contiguous()
    dpred = torch.empty_like(sign)
    _l.call("vitssl_l1_loss_bwd", _p(sign), _p(go), _p(dpred), sign.numel(), _l.stream_ptr())
    return dpred


def adamw_step(params, grads, exp_avgs, exp_avg_sqs, shadows, steps, lr, beta1, beta2, eps, weight_decay,
               grad_scale=None, found_inf=None):
    """Fused multi-tensor AdamW (csrc/optimizer.cu). All lists hold contiguous fp32 CUDA tensors of
    equal sizes per index; shadows[i] is a contiguous bf16 destination or None; steps[i] a 0-dim fp32
    device counter. grad_scale / found_inf: 0-dim fp32 device tensors or None."""
    import ctypes
    if not params:
        return
    _l.ensure_device()
    for p_, g_, m_, v_ in zip(params, grads, exp_avgs, exp_avg_sqs):
        assert p_.dtype == torch.float32 and g_.dtype == torch.float32 and p_.is_contiguous() and g_.is_contiguous()
        assert m_.is_contiguous() and v_.is_contiguous() and p_.numel() == g_.numel() == m_.numel() == v_.numel()
    sh = None
    if shadows is not None and any(t is not None for t in shadows):
        for p_, t in zip(params, shadows):
            assert t is None or (t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() == p_.numel())
        sh = _ptr_array(shadows)
    a, g, m, v, st, n = (_ptr_array(params), _ptr_array(grads), _ptr_array(exp_avgs), _ptr_array(exp_avg_sqs),
                         _ptr_array(steps), _numel_array(params))
    _l.call("vitssl_adamw_step", ctypes.addressof(a), ctypes.addressof(g), ctypes.addressof(m), ctypes.addressof(v),
            ctypes.addressof(sh) if sh is not None else None, ctypes.addressof(st), ctypes.addressof(n), len(params),
            float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), _p(grad_scale), _p(found_inf),
            _l.stream_ptr())


def l2norm_fwd(x):
    _l.ensure_device()
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 2
    y = torch.empty_like(x)
    inv = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
    _l.call("vitssl_l2norm_fwd", _p(x), _p(y), _p(inv), x.shape[0], x.shape[1], _l.stream_ptr())
    return y, inv


def l2norm_bwd(x, inv, dy):
    _l.ensure_device()
    dx = torch.empty_like(x)
    _l.call("vitssl_l2norm_bwd", _p(x), _p(inv), _p(dy), _p(dx), x.shape[0], x.shape[1], _l.stream_ptr())
    return dx


def weight_norm_fwd(v, g, out=None):
    _l.ensure_device()
    assert v.dtype == torch.float32 and v.is_contiguous() and g.is_contiguous() and g.numel() == v.shape[0]
    w = out if out is not None else torch.empty(v.shape, device=v.device, dtype=torch.bfloat16)
    inv = torch.empty(v.shape[0], device=v.device, dtype=torch.float32)
    _l.call("vitssl_weight_norm_fwd", _p(v), _p(g), _p(w), _p(inv), v.shape[0], v.shape[1], _l.stream_ptr())
    return w, inv


def weight_norm_bwd(dw, v, g, inv):
    _l.ensure_device()
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    dg = torch.empty_like(g)
    dv = torch.empty_like(v)
    _l.call("vitssl_weight_norm_bwd", _p(dw), _p(v), _p(g), _p(inv), _p(dg), _p(dv), v.shape[0], v.shape[1],
            _l.stream_ptr())
    return dg, dv


def center_ema(center, colsum, momentum, inv_rows):
    _l.ensure_device()
    out = torch.empty_like(center)
    _l.call("vitssl_center_ema", _p(center), _p(colsum), _p(out), center.numel(), float(momentum),
            float(inv_rows), _l.stream_ptr())
    return out


def dino_loss_fwd(teacher, student, center, teacher_temp, student_temp):
    _l.ensure_device()
    G, B, K = teacher.shape
    V = student.shape[0]
    assert teacher.dtype == torch.bfloat16 and student.dtype == torch.bfloat16
    assert teacher.is_contiguous() and student.is_contiguous() and center.dtype == torch.float32
    loss = torch.empty((), device=teacher.device, dtype=torch.float32)
    t_stats = torch.empty((G, B, 2), device=teacher.device, dtype=torch.float32)
    s_lse = torch.empty((V, B), device=teacher.device, dtype=torch.float32)
    _l.call("vitssl_dino_loss_fwd", _p(teacher), _p(student), _p(center), _p(loss), _p(t_stats), _p(s_lse),
            G, V, B, K, float(teacher_temp), float(student_temp), _l.stream_ptr())
    return loss, t_stats, s_lse


def dino_loss_bwd(teacher, student, center, t_stats, s_lse, grad_out, teacher_temp, student_temp):
    _l.ensure_device()
    G, B, K = teacher.shape
    V = student.shape[0]
    dstudent = torch.empty_like(student)
    go = grad_out.reshape(()).to(torch.float32).contiguous()
    _l.call("vitssl_dino_loss_bwd", _p(teacher), _p(student), _p(center), _p(t_stats), _p(s_lse), _p(go),
            _p(dstudent), G, V, B, K, float(teacher_temp), float(student_temp), _l.stream_ptr())
    return dstudent


# ----------------------------------------------------------------------------------------
# whole encoder stack in one C call (vitssl_encoder_stack_fwd / _bwd)
# ----------------------------------------------------------------------------------------
import ctypes as _ct

_PP = _ct.POINTER(_ct.c_void_p)


class _EncFwdArgs(_ct.Structure):
    _fields_ = ([(n, _ct.c_int64) for n in ("B", "S", "D", "H", "F", "L")]
                + [("dropout_p", _ct.c_float), ("eps", _ct.c_float), ("seed", _ct.c_uint64),
                   ("x_in", _ct.c_void_p), ("out", _ct.c_void_p), ("y1", _ct.c_void_p), ("y2", _ct.c_void_p * 2)]
                + [(n, _PP) for n in ("wqkv", "wo", "w1", "w2", "b1", "b2", "g1", "be1", "g2", "be2",
                                      "xs", "mean1", "rstd1", "xn1", "qkv", "ctx", "lse", "ctx_lo",
                                      "xmid", "mean2", "rstd2", "xn2", "u", "h")])


class _EncBwdArgs(_ct.Structure):
    _fields_ = ([("fwd", _ct.POINTER(_EncFwdArgs)), ("gout", _ct.c_void_p), ("dx", _ct.c_void_p)]
                + [(n, _ct.c_void_p) for n in ("dbranch", "du", "dxn", "dctx", "dqkv", "delta")]
                + [("gs", _ct.c_void_p * 2)]
                + [(n, _PP) for n in ("dwqkv", "dwo", "dw1", "db1", "dw2", "db2", "dg1", "dbe1", "dg2", "dbe2")]
                + [("l_begin", _ct.c_int64), ("l_end", _ct.c_int64)])


def _layer_ptrs(t, L, per_layer=True):
    """ctypes array of L device pointers into tensor t: slice l of its first dim, or t itself."""
    base = _p(t)
    step = t.stride(0) * t.element_size() if per_layer else 0
    return (_ct.c_void_p * L)(*[base + l * step for l in range(L)])


def _list_ptrs(ts):
    return (_ct.c_void_p * len(ts))(*[_p(t) for t in ts])


def encoder_stack_supported(S, D, H):
    return D == H * 64 and D % 8 == 0 and attention_supported(S, S, 64)


class EncoderStackState:
    """Buffers and the C argument block of one forward call (kept alive for its backward)."""
    __slots__ = ("args", "keep", "dims")


def encoder_stack_fwd(x, weights, params, H, p, seed, need_grad, eps=1e-5):
    """x fp32 [B,S,D] contiguous; weights[l] = (wqkv, wo, w1, w2) bf16 shadows; params = flat list of
    12 fp32 parameters per block (functional.block_params order). Returns (out fp32 [B,S,D], state)."""
    _l.ensure_device()
    B, S, D = x.shape
    L = len(weights)
    F_ = weights[0][2].shape[0]
    M = B * S
    dev = x.device
    bf, f32 = torch.bfloat16, torch.float32
    n = L if need_grad else 1          # saved activations: one slot per layer, or one reused slot
    per = need_grad
    xs = torch.empty((max(n - 1, 1), M, D), device=dev, dtype=f32)
    stats = torch.empty((4, n, M), device=dev, dtype=f32)
    xn1 = torch.empty((n, M, D), device=dev, dtype=bf)
    qkv = torch.empty((n, M, 3 * D), device=dev, dtype=bf)
    ctx = torch.empty((n, M, D), device=dev, dtype=bf)
    ctx_lo = torch.empty((n, M, D), device=dev, dtype=bf) if need_grad else None  # rounding residual of ctx (backward only)
    lse = torch.empty((n, B * H * S), device=dev, dtype=f32)
    xmid = torch.empty((n, M, D), device=dev, dtype=f32)
    xn2 = torch.empty((n, M, D), device=dev, dtype=bf)
    uh = torch.empty((2, n, M, F_), device=dev, dtype=bf)
    ytmp = torch.empty((3, M, D), device=dev, dtype=bf)
    out = torch.empty((B, S, D), device=dev, dtype=f32)
    a = _EncFwdArgs()
    a.B, a.S, a.D, a.H, a.F, a.L = B, S, D, H, F_, L
    a.dropout_p, a.eps, a.seed = float(p), float(eps), int(seed)
    a.x_in, a.out = _p(x), _p(out)
    a.y1 = _p(ytmp[0])
    a.y2[0], a.y2[1] = _p(ytmp[1]), _p(ytmp[2])
    keep = [x, out, xs, stats, xn1, qkv, ctx, lse, xmid, xn2, uh, ytmp, weights, params, ctx_lo]
    arrays = {}
    for i, name in enumerate(("wqkv", "wo", "w1", "w2")):
        arrays[name] = _list_ptrs([w[i] for w in weights])
    for i, name in ((5, "b1"), (7, "b2"), (8, "g1"), (9, "be1"), (10, "g2"), (11, "be2")):
        arrays[name] = _list_ptrs([params[12 * l + i] for l in range(L)])
    # xs[l] for l >= 1 lives in slot l-1 (layer 0 reads x_in)
    xs_base, xs_step = _p(xs), (xs.stride(0) * 4 if per else 0)
    arrays["xs"] = (_ct.c_void_p * L)(*[xs_base + max(l - 1, 0) * xs_step for l in range(L)])
    for i, name in enumerate(("mean1", "rstd1", "mean2", "rstd2")):
        arrays[name] = _layer_ptrs(stats[i], L, per)
    for name, t in (("xn1", xn1), ("qkv", qkv), ("ctx", ctx), ("lse", lse), ("xmid", xmid), ("xn2", xn2),
                    ("u", uh[0]), ("h", uh[1])):
        arrays[name] = _layer_ptrs(t, L, per)
    if ctx_lo is not None:
        arrays["ctx_lo"] = _layer_ptrs(ctx_lo, L, per)
    for name, arr in arrays.items():
        setattr(a, name, _ct.cast(arr, _PP))
    keep.append(arrays)
    _l.call("vitssl_encoder_stack_fwd", _ct.addressof(a), _l.stream_ptr())
    st = EncoderStackState()
    st.args, st.keep, st.dims = a, keep, (B, S, D, H, F_, L)
    return out, st


def encoder_stack_last_qkv(st):
    """bf16 [B,S,3D] fused QKV activations of the LAST block (return_attn=True path)."""
    B, S, D, H, F_, L = st.dims
    qkv = st.keep[5]
    return qkv[qkv.shape[0] - 1].view(B, S, 3 * D)


_STACK_GRAD_SHAPES = lambda D, F_: (("dwqkv", (3 * D, D)), ("dwo", (D, D)), ("dw1", (F_, D)), ("db1", (F_,)),
                                    ("dw2", (D, F_)), ("db2", (D,)), ("dg1", (D,)), ("dbe1", (D,)), ("dg2", (D,)),
                                    ("dbe2", (D,)))


HOST_TRACE = None  # developer diagnostic: set to a list to collect host times of the stack's C calls


class EncoderStackBackward:
    """Backward of one `encoder_stack_fwd` call, runnable as descending layer ranges.

    Every parameter gradient of the stack lives in ONE zeroed fp32 buffer, layer-major: the C side
    accumulates into it (split-K TMA reduce-adds, fused bias column sums, LayerNorm dgamma/dbeta), so
    a single fill replaces ~125 per-kernel memsets per step, a range of layers is one contiguous
    slice (data parallelism all-reduces it while the next range computes), and a second pass over
    the same blocks (`flat=` given) simply keeps accumulating."""

    def __init__(self, st, gout, flat=None):
        B, S, D, H, F_, L = st.dims
        M = B * S
        dev = gout.device
        bf, f32 = torch.bfloat16, torch.float32
        self.st, self.L, self.D = st, L, D
        self.dx = torch.empty((B, S, D), device=dev, dtype=f32)
        tmp_d = torch.empty((3, M, D), device=dev, dtype=bf)        # dbranch, dxn, dctx
        du = torch.empty((M, F_), device=dev, dtype=bf)
        dqkv = torch.empty((M, 3 * D), device=dev, dtype=bf)
        gs = torch.empty((2, M, D), device=dev, dtype=f32)
        delta = torch.empty((B * H * S,), device=dev, dtype=f32)
        shapes = _STACK_GRAD_SHAPES(D, F_)
        self.per_layer = sum(int(torch.Size(shp).numel()) for _, shp in shapes)
        if flat is None:
            flat = torch.zeros((L * self.per_layer,), device=dev, dtype=f32)
        assert flat.numel() == L * self.per_layer
        self.flat = flat
        g = {k: [] for k, _ in shapes}
        for l in range(L):
            off = l * self.per_layer
            for k, shp in shapes:
                n = int(torch.Size(shp).numel())
                g[k].append(flat[off:off + n].view(shp))
                off += n
        self.g = g
        b = _EncBwdArgs()
        b.fwd = _ct.pointer(st.args)
        b.gout, b.dx = _p(gout), _p(self.dx)
        b.dbranch, b.dxn, b.dctx = _p(tmp_d[0]), _p(tmp_d[1]), _p(tmp_d[2])
        b.du, b.dqkv, b.delta = _p(du), _p(dqkv), _p(delta)
        b.gs[0], b.gs[1] = _p(gs[0]), _p(gs[1])
        arrays = {name: (_ct.c_void_p * L)(*[_p(t) for t in views]) for name, views in g.items()}
        for name, arr in arrays.items():
            setattr(b, name, _ct.cast(arr, _PP))
        self.args = b
        self.keep = (gout, tmp_d, du, dqkv, gs, delta, arrays)

    def run(self, l_begin, l_end):
        self.args.l_begin, self.args.l_end = int(l_begin), int(l_end)
        if HOST_TRACE is not None:
            import time as _t
            t0 = _t.perf_counter()
            _l.call("vitssl_encoder_stack_bwd", _ct.addressof(self.args), _l.stream_ptr())
            HOST_TRACE.append(("stack_bwd_c_call_ms", round((_t.perf_counter() - t0) * 1e3, 2)))
            return
        _l.call("vitssl_encoder_stack_bwd", _ct.addressof(self.args), _l.stream_ptr())

    def flat_slice(self, l_begin, l_end):
        return self.flat[l_begin * self.per_layer:l_end * self.per_layer]

    def param_grads(self):
        """Gradient views in functional.block_params order (12 per layer)."""
        D, g, out = self.D, self.g, []
        for l in range(self.L):
            wq = g["dwqkv"][l]
            out += [wq[:D], wq[D:2 * D], wq[2 * D:], g["dwo"][l], g["dw1"][l], g["db1"][l], g["dw2"][l],
                    g["db2"][l], g["dg1"][l], g["dbe1"][l], g["dg2"][l], g["dbe2"][l]]
        return out


def encoder_stack_bwd(st, gout, n_chunks=1, on_chunk=None):
    """gout fp32 [B,S,D] contiguous -> (dx fp32 [B,S,D], per-layer gradient views dict name -> list of L).
    The backward may run as `n_chunks` C calls over descending layer ranges; after each one
    `on_chunk(flat_slice, l_lo, l_hi)` is called with that range's contiguous gradient slice."""
    run = EncoderStackBackward(st, gout)
    L = run.L
    n_chunks = max(1, min(int(n_chunks), L))
    bounds = [round(i * L / n_chunks) for i in range(n_chunks + 1)]
    for c in reversed(range(n_chunks)):
        run.run(bounds[c], bounds[c + 1])
        if on_chunk is not None:
            on_chunk(run.flat_slice(bounds[c], bounds[c + 1]), bounds[c], bounds[c + 1])
    return run.dx, run.g
