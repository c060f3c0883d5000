"""Autograd glue: torch.autograd.Function wrappers that sequence the sm_100a kernels.

Numerical contract (matches the reference under `autocast(bf16)`): parameters,
the residual stream, LayerNorm statistics, losses and all gradients of parameters are fp32;
GEMM / attention operands and their outputs are bf16 with fp32 accumulation.
"""
from __future__ import annotations

import math
import weakref
from types import SimpleNamespace
from typing import List, Optional, Sequence

import torch

from . import dp, ops
from .ops import EPI_BIAS, EPI_BIAS_GELU_D, EPI_MUL, EPI_NONE


# ----------------------------------------------------------------------------------------
# bf16 weight shadows
# ----------------------------------------------------------------------------------------
def _cache_of(owner) -> dict:
    c = owner.__dict__.get("_vitssl_cache")
    if c is None:
        c = {}
        object.__setattr__(owner, "_vitssl_cache", c)
    return c


# id(param) -> (weakref to the cache dict's owner, key, bf16 destination view): lets the fused
# optimizer (vit_core/optim.py) write the shadow in its update pass and mark it fresh
_SHADOW_REG: dict = {}


def bf16_shadows(requests):
    """requests: list of (owner_module, key, [fp32 params]) -> list of bf16 tensors.

    Each shadow is the row-wise concatenation of its params (viewed 2-D). Stale shadows (parameter
    version or storage changed) are refreshed with ONE multi-tensor cast launch for the whole list.
    """
    outs, srcs, dsts = [], [], []
    for owner, key, params in requests:
        cache = _cache_of(owner)
        sig = tuple((p.data_ptr(), p._version) for p in params)
        ent = cache.get(key)
        dev = params[0].device
        if ent is not None and ent[1] == sig and ent[0].device == dev:
            outs.append(ent[0])
            continue
        rows = sum(p.shape[0] for p in params)
        cols = params[0].numel() // params[0].shape[0]
        if ent is not None and ent[0].shape == (rows, cols) and ent[0].device == dev:
            buf = ent[0]
        else:
            buf = torch.empty((rows, cols), device=dev, dtype=torch.bfloat16)
        r = 0
        views = []
        for p in params:
            pd = p.detach()
            srcs.append(pd if pd.is_contiguous() else pd.contiguous())
            dsts.append(buf[r:r + p.shape[0]])
            views.append(dsts[-1])
            r += p.shape[0]
        cache[key] = (buf, sig, tuple(weakref.ref(p) for p in params))
        owner_ref = weakref.ref(owner)
        for p, v in zip(params, views):
            _SHADOW_REG[id(p)] = (owner_ref, key, v, weakref.ref(p))
        outs.append(buf)
    if srcs:
        ops.multi_cast_bf16(srcs, dsts)
    return outs


def shadow_destination(param):
    """bf16 view that shadows `param` in some module's GEMM-operand cache, or None."""
    ent = _SHADOW_REG.get(id(param))
    if ent is None:
        return None
    owner, key, view, pref = ent[0](), ent[1], ent[2], ent[3]()
    if owner is None or pref is not param:
        _SHADOW_REG.pop(id(param), None)
        return None
    cur = owner.__dict__.get("_vitssl_cache", {}).get(key)
    if cur is None or cur[0].data_ptr() > view.data_ptr() or view.device != param.device or not param.is_contiguous():
        return None
    if view.data_ptr() + view.numel() * 2 > cur[0].data_ptr() + cur[0].numel() * 2:
        return None
    return view


def mark_shadows_fresh(params):
    """After an in-place update that also wrote the shadows (fused optimizer): re-sign the cache
    entries of `params` with their current versions so the next forward skips the cast."""
    seen = set()
    for p in params:
        ent = _SHADOW_REG.get(id(p))
        if ent is None:
            continue
        owner = ent[0]()
        if owner is None or (id(owner), ent[1]) in seen:
            continue
        seen.add((id(owner), ent[1]))
        cache = owner.__dict__.get("_vitssl_cache", {})
        cur = cache.get(ent[1])
        if cur is None:
            continue
        ps = [r() for r in cur[2]]
        if any(q is None for q in ps):
            continue
        cache[ent[1]] = (cur[0], tuple((q.data_ptr(), q._version) for q in ps), cur[2])


def _new_seed() -> int:
    # CPU generator: deterministic under torch.manual_seed, no device sync
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _as_bf16(x: torch.Tensor) -> torch.Tensor:
    if x.dtype == torch.bfloat16:
        return x if x.is_contiguous() else x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    return ops.cast_bf16(x if x.is_contiguous() else x.contiguous())


def autocast_out(y: torch.Tensor) -> torch.Tensor:
    """Linear outputs are bf16 inside autocast (like the reference) and fp32 outside it."""
    if torch.is_autocast_enabled():
        return y
    return y.float() if y.dtype != torch.float32 else y


# ----------------------------------------------------------------------------------------
# attention core shared by the encoder stack and MultiHeadedAttention
# ----------------------------------------------------------------------------------------
def _attn_fwd(q, k, v, H, want_lse):
    """q: [B,Sq,D] bf16 view, k/v: [B,Sk,D] views (unit inner stride). Returns (ctx [B,Sq,D], lse,
    ctx_lo): ctx_lo (tcgen05 path, training) is the bf16 rounding residual of ctx for the backward."""
    B, Sq, D = q.shape
    Sk = k.shape[1]
    dk = D // H
    scale = 1.0 / math.sqrt(dk)
    if ops.attention_supported(Sq, Sk, dk):
        return ops.attention_fwd(q, k, v, H, scale, want_lse=want_lse)
    qh, kh, vh = (t.unflatten(2, (H, dk)).transpose(1, 2) for t in (q, k, v))
    out, _, lse = ops.attention_generic_fwd(qh, kh, vh, scale, want_probs=False, want_lse=want_lse)
    return out.transpose(1, 2).reshape(B, Sq, D), lse, None


def _attn_bwd(q, k, v, ctx_, dctx, lse, H, dq, dk_, dv, ctx_lo=None):
    B, Sq, D = q.shape
    Sk = k.shape[1]
    dk = D // H
    scale = 1.0 / math.sqrt(dk)
    if ops.attention_supported(Sq, Sk, dk) and Sq <= 256:
        ops.attention_bwd(q, k, v, ctx_, dctx, lse, H, scale, dq, dk_, dv, out_lo=ctx_lo)
        return
    qh, kh, vh = (t.unflatten(2, (H, dk)).transpose(1, 2) for t in (q, k, v))
    oh = ctx_.unflatten(2, (H, dk)).transpose(1, 2)
    doh = dctx.unflatten(2, (H, dk)).transpose(1, 2)
    gq, gk, gv = ops.attention_generic_bwd(qh, kh, vh, oh, doh, lse, scale)
    dq.copy_(gq.transpose(1, 2).reshape(B, Sq, D))
    dk_.copy_(gk.reshape(B, Sk, D))
    dv.copy_(gv.reshape(B, Sk, D))


def attention_probs(q, k, H):
    """fp32 probabilities [B,H,Sq,Sk] (return_attn=True path, attention.py:24-25)."""
    B, Sq, D = q.shape
    dk = D // H
    qh, kh = (t.unflatten(2, (H, dk)).transpose(1, 2) for t in (q, k))
    _, probs, _ = ops.attention_generic_fwd(qh, kh, kh, 1.0 / math.sqrt(dk), want_probs=True, want_lse=False)
    return probs


# ----------------------------------------------------------------------------------------
# encoder stack: L pre-LN blocks in one autograd node (encoder_block.py:32-53)
# params per block (12): wq wk wv wo | w1 b1 w2 b2 | g1 be1 g2 be2
# dropout sites per block l: 3l (after out-proj), 3l+1 (after GELU), 3l+2 (after FFN)
# ----------------------------------------------------------------------------------------
import os as _os

# layer chunks of the stack backward under data parallelism (3 layers each for L = 12 at the default 4)
DP_STACK_CHUNKS = int(_os.environ.get("VITSSL_DP_CHUNKS", "4"))


def _take_saved(ctx):
    """The tensors a node stored as `ctx.saved`, released from the node: like autograd's own saved tensors
    they are gone after the first backward, so a loss object that outlives its step (the trainers keep
    `loss` until the next iteration assigns it) does not pin this step's buffers during the next forward."""
    saved, ctx.saved = ctx.saved, None
    if saved is None:
        raise RuntimeError("vit_core: trying to backward through the graph a second time (the node's saved "
                           "buffers were released by the first backward; retain_graph is not supported by the fused nodes)")
    return saved


class _EncoderStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, meta, *params):
        L, H = meta.L, meta.H
        B, S, D = x.shape
        M = B * S
        need_grad = any(ctx.needs_input_grad)
        p = meta.p if meta.training else 0.0
        seed = _new_seed() if p > 0 else 0
        x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
        if ops.encoder_stack_supported(S, D, H):
            # one C call enqueues every kernel of the L blocks (csrc/encoder.cu)
            out, st = ops.encoder_stack_fwd(x, meta.weights, params, H, p, seed, need_grad)
            probs = None
            if meta.return_attn:
                q3 = ops.encoder_stack_last_qkv(st)
                probs = attention_probs(q3[..., :D], q3[..., D:2 * D], H)
            if need_grad:
                ctx.stack_state = st
                ctx.meta, ctx.params = meta, params
                ctx.shape = (B, S, D)
            if probs is not None:
                ctx.mark_non_differentiable(probs)
                return out, probs
            return out
        ctx.stack_state = None
        stream = x.view(M, D)
        branch = None
        saved = []
        last_qkv = None
        for l in range(L):
            wqkv, wo, w1, w2 = meta.weights[l]
            _, _, _, _, _, b1, _, b2, g1, be1, g2, be2 = params[12 * l:12 * l + 12]
            F_ = w1.shape[0]
            xs, xn1, mean1, rstd1 = ops.add_layernorm_fwd(
                stream, branch, g1, be1, dropout_p=p if branch is not None else 0.0, seed=seed,
                offset=max(3 * l - 1, 0))
            qkv = ops.gemm(xn1, wqkv)
            qkv3 = qkv.view(B, S, 3 * D)
            ctx_, lse, ctx_lo = _attn_fwd(qkv3[..., :D], qkv3[..., D:2 * D], qkv3[..., 2 * D:], H, need_grad)
            y1 = ops.gemm(ctx_.view(M, D), wo)
            xmid, xn2, mean2, rstd2 = ops.add_layernorm_fwd(xs, y1, g2, be2, dropout_p=p, seed=seed,
                                                            offset=3 * l)
            u = torch.empty((M, F_), device=x.device, dtype=torch.bfloat16)
            h = ops.gemm(xn2, w1, epilogue=EPI_BIAS_GELU_D, bias=b1, aux=u, dropout_p=p, seed=seed,
                         offset=3 * l + 1)  # u <- mask/(1-p) * gelu'(pre-activation): the whole backward factor
            y2 = ops.gemm(h, w2, epilogue=EPI_BIAS, bias=b2)
            stream, branch = xmid, y2
            last_qkv = qkv3
            if need_grad:
                saved.append((xs, mean1, rstd1, xn1, qkv3, ctx_, lse, xmid, mean2, rstd2, xn2, u, h, ctx_lo))
        out, _, _, _ = ops.add_layernorm_fwd(stream, branch, None, None, dropout_p=p, seed=seed,
                                             offset=3 * L - 1)
        probs = None
        if meta.return_attn:
            probs = attention_probs(last_qkv[..., :D], last_qkv[..., D:2 * D], H)
        if need_grad:
            ctx.saved = saved
            ctx.meta = meta
            ctx.params = params
            ctx.p, ctx.seed = p, seed
            ctx.shape = (B, S, D)
        out = out.view(B, S, D)
        if probs is not None:
            ctx.mark_non_differentiable(probs)
            return out, probs
        return out

    @staticmethod
    def backward(ctx, gout, *_):
        g = gout if (gout.dtype == torch.float32 and gout.is_contiguous()) else gout.float().contiguous()
        if ctx.stack_state is not None:
            st, ctx.stack_state = ctx.stack_state, None
            L = ctx.meta.L
            D = ctx.shape[2]
            # data parallel: the stack runs as a few layer chunks, and each chunk's gradient slice
            # starts its all-reduce on the communication stream while the next chunk computes
            sync = dp.sync_for(ctx.params) if all(ctx.needs_input_grad[2:]) else None
            run = ops.EncoderStackBackward(st, g)
            bounds = dp_chunk_bounds(L, DP_STACK_CHUNKS) if sync is not None else [0, L]
            n_chunks = len(bounds) - 1
            for c in reversed(range(n_chunks)):
                run.run(bounds[c], bounds[c + 1])
                if sync is not None:
                    sync.prereduce(run.flat_slice(bounds[c], bounds[c + 1]), ctx.params[12 * bounds[c]:12 * bounds[c + 1]])
            if sync is not None:
                sync.join()  # only the last chunk's all-reduce is still in flight here
            dx, grads = run.dx, run.param_grads()
            for i, need in enumerate(ctx.needs_input_grad[2:]):
                if not need:
                    grads[i] = None
            return (dx if ctx.needs_input_grad[0] else None, None, *grads)
        meta, params, saved = ctx.meta, ctx.params, _take_saved(ctx)
        L, H = meta.L, meta.H
        B, S, D = ctx.shape
        M = B * S
        p, seed = ctx.p, ctx.seed
        gs = g.view(M, D)
        _, dbranch, _, _ = ops.add_layernorm_bwd(None, None, None, None, None, gs, want_dx=False,
                                                 want_dbranch=True, dropout_p=p, seed=seed,
                                                 offset=3 * L - 1)
        grads: List[Optional[torch.Tensor]] = [None] * (12 * L)
        for l in reversed(range(L)):
            wqkv, wo, w1, w2 = meta.weights[l]
            g1, g2 = params[12 * l + 8], params[12 * l + 10]
            xs, mean1, rstd1, xn1, qkv3, ctx_, lse, xmid, mean2, rstd2, xn2, u, h, ctx_lo = saved[l]
            saved[l] = None
            dy2 = dbranch
            du = ops.gemm(dy2, w2, b_mn=True, epilogue=EPI_MUL, aux=u)
            dW2 = ops.gemm(dy2, h, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1)
            db2 = ops.colsum_bf16(dy2)
            dxn2 = ops.gemm(du, w1, b_mn=True)
            dW1 = ops.gemm(du, xn2, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1)
            db1 = ops.colsum_bf16(du)
            del du, u, h
            gs, dy1, dg2, dbe2 = ops.add_layernorm_bwd(dxn2, xmid, mean2, rstd2, g2, gs,
                                                       want_dbranch=True, dropout_p=p, seed=seed,
                                                       offset=3 * l)
            dctx = ops.gemm(dy1, wo, b_mn=True)
            dWo = ops.gemm(dy1, ctx_.view(M, D), a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1)
            dqkv = torch.empty((B, S, 3 * D), device=g.device, dtype=torch.bfloat16)
            _attn_bwd(qkv3[..., :D], qkv3[..., D:2 * D], qkv3[..., 2 * D:], ctx_, dctx.view(B, S, D), lse, H,
                      dqkv[..., :D], dqkv[..., D:2 * D], dqkv[..., 2 * D:], ctx_lo=ctx_lo)
            dqkv2 = dqkv.view(M, 3 * D)
            dxn1 = ops.gemm(dqkv2, wqkv, b_mn=True)
            dWqkv = ops.gemm(dqkv2, xn1, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1)
            gs, dbranch, dg1, dbe1 = ops.add_layernorm_bwd(dxn1, xs, mean1, rstd1, g1, gs,
                                                           want_dbranch=(l > 0), dropout_p=p if l > 0 else 0.0,
                                                           seed=seed, offset=max(3 * l - 1, 0))
            grads[12 * l:12 * l + 12] = [dWqkv[:D], dWqkv[D:2 * D], dWqkv[2 * D:], dWo, dW1, db1, dW2, db2,
                                         dg1, dbe1, dg2, dbe2]
        # data parallel: this path pre-reduces its contributions exactly like the fused one, so a
        # parameter shared by a fused node and a per-op node (DINO global / local passes with
        # different S) is averaged once per contribution and the hooks skip it consistently
        sync = dp.sync_for(params) if all(ctx.needs_input_grad[2:]) else None
        if sync is not None:
            flat = torch.cat([g_.reshape(-1) for g_ in grads])
            sync.prereduce(flat, params)
            sync.join()
            off = 0
            for i, g_ in enumerate(grads):
                grads[i] = flat[off:off + g_.numel()].view(g_.shape)
                off += g_.numel()
        for i, need in enumerate(ctx.needs_input_grad[2:]):
            if not need:
                grads[i] = None
        dx = gs.view(B, S, D) if ctx.needs_input_grad[0] else None
        return (dx, None, *grads)


class _EncoderStackMultiFn(torch.autograd.Function):
    """Several token batches (different B and S) through the SAME L blocks as ONE autograd node, e.g.
    DINO's global-crop and local-crop student passes (ssl/dino/model.py:117-118). Every pass's
    parameter gradients accumulate into one zeroed layer-major buffer on the C side, so autograd
    never adds per-parameter gradients of the two passes (149 tiny launches per step before), and
    under data parallelism a layer chunk is all-reduced once, after all passes contributed to it."""

    @staticmethod
    def forward(ctx, meta, n_in, *args):
        xs, params = args[:n_in], args[n_in:]
        need_grad = any(ctx.needs_input_grad)
        p = meta.p if meta.training else 0.0
        outs, states = [], []
        for x in xs:
            x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
            seed = _new_seed() if p > 0 else 0
            out, st = ops.encoder_stack_fwd(x, meta.weights, params, meta.H, p, seed, need_grad)
            outs.append(out)
            states.append(st)
        if need_grad:
            ctx.states, ctx.meta, ctx.params, ctx.n_in = states, meta, params, n_in
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        states, ctx.states = ctx.states, None
        n_in, L = ctx.n_in, ctx.meta.L
        runs, flat = [], None
        for st, g in zip(states, gouts):
            g = g if (g.dtype == torch.float32 and g.is_contiguous()) else g.float().contiguous()
            r = ops.EncoderStackBackward(st, g, flat=flat)
            flat = r.flat
            runs.append(r)
        sync = dp.sync_for(ctx.params) if all(ctx.needs_input_grad[2 + n_in:]) else None
        bounds = dp_chunk_bounds(L, DP_STACK_CHUNKS) if sync is not None else [0, L]
        n_chunks = len(bounds) - 1
        for c in reversed(range(n_chunks)):
            for r in runs:
                r.run(bounds[c], bounds[c + 1])
            if sync is not None:
                sync.prereduce(runs[0].flat_slice(bounds[c], bounds[c + 1]), ctx.params[12 * bounds[c]:12 * bounds[c + 1]])
        if sync is not None:
            sync.join()
        grads = runs[0].param_grads()
        for i, need in enumerate(ctx.needs_input_grad[2 + n_in:]):
            if not need:
                grads[i] = None
        dxs = [r.dx if ctx.needs_input_grad[2 + i] else None for i, r in enumerate(runs)]
        return (None, None, *dxs, *grads)


def block_params(blk) -> list:
    a, f = blk.self_attention, blk.feed_forward
    return [a.w_query.weight, a.w_key.weight, a.w_value.weight, a.final_linear.weight,
            f.linear_in.weight, f.linear_in.bias, f.linear_out.weight, f.linear_out.bias,
            blk.layer_norm1.weight, blk.layer_norm1.bias, blk.layer_norm2.weight, blk.layer_norm2.bias]


def encoder_stack(blocks: Sequence, x: torch.Tensor, return_attn: bool = False):
    """Run a sequence of EncoderBlock modules as one fused autograd node.

    Returns (x_out fp32 [B,S,D], probs of the LAST block or None) — vit.py:35-45.
    """
    blocks = list(blocks)
    if not blocks:
        return x, None
    reqs = []
    for blk in blocks:
        a, f = blk.self_attention, blk.feed_forward
        reqs += [(a, "wqkv", [a.w_query.weight, a.w_key.weight, a.w_value.weight]),
                 (a, "wo", [a.final_linear.weight]),
                 (f, "w1", [f.linear_in.weight]),
                 (f, "w2", [f.linear_out.weight])]
    sh = bf16_shadows(reqs)
    b0 = blocks[0]
    meta = SimpleNamespace(
        L=len(blocks), H=b0.self_attention.num_heads, p=float(b0.drop1.p),
        training=bool(b0.training), return_attn=bool(return_attn),
        weights=[tuple(sh[4 * i:4 * i + 4]) for i in range(len(blocks))])
    params = []
    for blk in blocks:
        params += block_params(blk)
    out = _EncoderStackFn.apply(x, meta, *params)
    if return_attn:
        return out[0], out[1]
    return out, None


def dp_chunk_bounds(L: int, n_chunks: int):
    """Layer boundaries of the data-parallel stack backward (processed from the top chunk down). The
    chunk that runs LAST — the lowest layers — is a single layer: its all-reduce is the one nothing
    is left to overlap with, so it should be the smallest (7 MB instead of 22 MB for ViT-S/16)."""
    n_chunks = max(1, min(int(n_chunks), L))
    if n_chunks < 3 or L < n_chunks + 1:
        return [round(i * L / n_chunks) for i in range(n_chunks + 1)]
    rest = [1 + round(i * (L - 1) / (n_chunks - 1)) for i in range(n_chunks)]
    return [0] + rest


def encoder_stack_multi(blocks: Sequence, xs: Sequence[torch.Tensor]):
    """`encoder_stack` for several token batches sharing the blocks; returns a list of outputs."""
    blocks = list(blocks)
    xs = list(xs)
    H = blocks[0].self_attention.num_heads if blocks else 0
    if len(xs) < 2 or not blocks or not all(ops.encoder_stack_supported(x.shape[1], x.shape[2], H) for x in xs):
        return [encoder_stack(blocks, x)[0] for x in xs]
    reqs = []
    for blk in blocks:
        a, f = blk.self_attention, blk.feed_forward
        reqs += [(a, "wqkv", [a.w_query.weight, a.w_key.weight, a.w_value.weight]),
                 (a, "wo", [a.final_linear.weight]),
                 (f, "w1", [f.linear_in.weight]),
                 (f, "w2", [f.linear_out.weight])]
    sh = bf16_shadows(reqs)
    b0 = blocks[0]
    meta = SimpleNamespace(L=len(blocks), H=H, p=float(b0.drop1.p), training=bool(b0.training),
                           weights=[tuple(sh[4 * i:4 * i + 4]) for i in range(len(blocks))])
    params = []
    for blk in blocks:
        params += block_params(blk)
    return list(_EncoderStackMultiFn.apply(meta, len(xs), *xs, *params))


# ----------------------------------------------------------------------------------------
# generic MLP: chain of Linear(+GELU)(+dropout after GELU). Used by FeedForwardBlock
# (feed_forward.py:26-28), the DINO head MLP (ssl/dino/head.py:10-16) and plain linears.
# ----------------------------------------------------------------------------------------
class _MLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, meta, *params):
        # params: (w, b) per layer (b may be None); meta.weights: bf16 shadows; meta.gelu: flags
        n = len(meta.gelu)
        lead = x.shape[:-1]
        xb = _as_bf16(x.reshape(-1, x.shape[-1]))
        need_grad = any(ctx.needs_input_grad)
        p = meta.p if meta.training else 0.0
        seed = _new_seed() if p > 0 else 0
        acts, pre = [xb], []
        cur = xb
        for i in range(n):
            w, b = meta.weights[i], params[2 * i + 1]
            last = i == n - 1
            out_dtype = meta.out_dtype if last else torch.bfloat16
            if meta.gelu[i]:
                u = torch.empty((cur.shape[0], w.shape[0]), device=cur.device, dtype=torch.bfloat16)
                cur = ops.gemm(cur, w, epilogue=EPI_BIAS_GELU_D, bias=b, aux=u, dropout_p=p, seed=seed, offset=i)
                pre.append(u)
            else:
                cur = ops.gemm(cur, w, epilogue=EPI_BIAS if b is not None else EPI_NONE, bias=b,
                               out_dtype=out_dtype)
                pre.append(None)
            if not last:
                acts.append(cur)
        if need_grad:
            ctx.acts, ctx.pre, ctx.meta, ctx.p, ctx.seed = acts, pre, meta, p, seed
            ctx.in_dtype, ctx.in_shape = x.dtype, x.shape
        return cur.view(*lead, cur.shape[-1])

    @staticmethod
    def backward(ctx, gout):
        meta, acts, pre = ctx.meta, ctx.acts, ctx.pre
        n = len(meta.gelu)
        p, seed = ctx.p, ctx.seed
        dy = _as_bf16(gout.reshape(-1, gout.shape[-1]))
        grads = [None] * (2 * n)
        for i in reversed(range(n)):
            w = meta.weights[i]
            # dy is the gradient of layer i's linear output (pre-activation) here
            if ctx.needs_input_grad[2 + 2 * i]:
                grads[2 * i] = ops.gemm(dy, acts[i], a_mn=True, b_mn=True, out_dtype=torch.float32,
                                        split_k=-1).view(meta.wshapes[i])
            if ctx.needs_input_grad[3 + 2 * i]:
                grads[2 * i + 1] = ops.colsum_bf16(dy)
            if i > 0:
                if meta.gelu[i - 1]:
                    dy = ops.gemm(dy, w, b_mn=True, epilogue=EPI_MUL, aux=pre[i - 1])
                else:
                    dy = ops.gemm(dy, w, b_mn=True)
            elif ctx.needs_input_grad[0]:
                dy = ops.gemm(dy, w, b_mn=True, out_dtype=torch.float32 if ctx.in_dtype == torch.float32
                              else torch.bfloat16)
        dx = dy.view(ctx.in_shape).to(ctx.in_dtype) if ctx.needs_input_grad[0] else None
        return (dx, None, *grads)


def mlp(x, owner_layers, gelu_flags, *, dropout_p=0.0, training=False, out_dtype=torch.bfloat16):
    """owner_layers: list of modules with .weight [N,K...] and optional .bias."""
    reqs = [(m, "w", [m.weight]) for m in owner_layers]
    sh = bf16_shadows(reqs)
    meta = SimpleNamespace(gelu=list(gelu_flags), weights=sh, p=float(dropout_p), training=bool(training),
                           out_dtype=out_dtype, wshapes=[m.weight.shape for m in owner_layers])
    params = []
    for m in owner_layers:
        params += [m.weight, getattr(m, "bias", None)]
    return _MLPFn.apply(x, meta, *params)


# ----------------------------------------------------------------------------------------
# MultiHeadedAttention with arbitrary query / key / value inputs (attention.py:61-106)
# ----------------------------------------------------------------------------------------
class _MHAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_in, k_in, v_in, meta, wq, wk, wv, wo):
        H = meta.H
        B, Sq, D = q_in.shape
        Sk = k_in.shape[1]
        same = meta.same
        wqkv, wob = meta.weights
        qb = _as_bf16(q_in.reshape(-1, D))
        if same:
            qkv = ops.gemm(qb, wqkv).view(B, Sq, 3 * D)
            q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
            kb = vb = qb
        else:
            kb = _as_bf16(k_in.reshape(-1, D))
            vb = _as_bf16(v_in.reshape(-1, D))
            q = ops.gemm(qb, wqkv[:D]).view(B, Sq, D)
            k = ops.gemm(kb, wqkv[D:2 * D]).view(B, Sk, D)
            v = ops.gemm(vb, wqkv[2 * D:]).view(B, Sk, D)
        ctx_, lse, ctx_lo = _attn_fwd(q, k, v, H, True)
        out = ops.gemm(ctx_.view(-1, D), wob).view(B, Sq, D)
        probs = attention_probs(q, k, H) if meta.return_attn else None
        ctx.saved = (qb, kb, vb, q, k, v, ctx_, lse, ctx_lo)
        ctx.meta = meta
        ctx.in_dtype = q_in.dtype
        if probs is not None:
            ctx.mark_non_differentiable(probs)
            return out, probs
        return out

    @staticmethod
    def backward(ctx, gout, *_):
        meta = ctx.meta
        H = meta.H
        qb, kb, vb, q, k, v, ctx_, lse, ctx_lo = _take_saved(ctx)
        wqkv, wob = meta.weights
        B, Sq, D = q.shape
        Sk = k.shape[1]
        dy = _as_bf16(gout.reshape(-1, D))
        dctx = ops.gemm(dy, wob, b_mn=True).view(B, Sq, D)
        dWo = ops.gemm(dy, ctx_.view(-1, D), a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1)
        f32 = torch.float32
        if meta.same:
            dqkv = torch.empty((B, Sq, 3 * D), device=dy.device, dtype=torch.bfloat16)
            _attn_bwd(q, k, v, ctx_, dctx, lse, H, dqkv[..., :D], dqkv[..., D:2 * D], dqkv[..., 2 * D:], ctx_lo=ctx_lo)
            d2 = dqkv.view(-1, 3 * D)
            dx = ops.gemm(d2, wqkv, b_mn=True, out_dtype=f32).view(B, Sq, D)
            dW = ops.gemm(d2, qb, a_mn=True, b_mn=True, out_dtype=f32, split_k=-1)
            third = dx / 3.0  # the same tensor was passed three times: autograd sums the three slots
            return third, third, third, None, dW[:D], dW[D:2 * D], dW[2 * D:], dWo
        dq = torch.empty((B, Sq, D), device=dy.device, dtype=torch.bfloat16)
        dk = torch.empty((B, Sk, D), device=dy.device, dtype=torch.bfloat16)
        dv = torch.empty((B, Sk, D), device=dy.device, dtype=torch.bfloat16)
        _attn_bwd(q, k, v, ctx_, dctx, lse, H, dq, dk, dv, ctx_lo=ctx_lo)
        outs = []
        for g_, w_, xb_, S_ in ((dq, wqkv[:D], qb, Sq), (dk, wqkv[D:2 * D], kb, Sk), (dv, wqkv[2 * D:], vb, Sk)):
            g2 = g_.view(-1, D)
            outs.append((ops.gemm(g2, w_, b_mn=True, out_dtype=f32).view(B, S_, D),
                         ops.gemm(g2, xb_, a_mn=True, b_mn=True, out_dtype=f32, split_k=-1)))
        return (outs[0][0], outs[1][0], outs[2][0], None, outs[0][1], outs[1][1], outs[2][1], dWo)


def multi_head_attention(mod, query, key, value, return_attn=False):
    wqkv, wo = bf16_shadows([(mod, "wqkv", [mod.w_query.weight, mod.w_key.weight, mod.w_value.weight]),
                             (mod, "wo", [mod.final_linear.weight])])
    same = (query is key) and (key is value)
    meta = SimpleNamespace(H=mod.num_heads, same=same, weights=(wqkv, wo), return_attn=bool(return_attn))
    out = _MHAFn.apply(query, key, value, meta, mod.w_query.weight, mod.w_key.weight, mod.w_value.weight,
                       mod.final_linear.weight)
    if return_attn:
        return out[0], out[1]
    return out, None


# ----------------------------------------------------------------------------------------
# ScaledDotProductAttention on already-projected heads (attention.py:5-27), any head dim
# ----------------------------------------------------------------------------------------
class _SDPAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, return_attn):
        lead = q.shape[:-2]
        Sq, d = q.shape[-2:]
        Sk = k.shape[-2]
        qb = _as_bf16(q.reshape(-1, 1, Sq, d))
        kb = _as_bf16(k.reshape(-1, 1, Sk, d))
        vb = _as_bf16(v.reshape(-1, 1, Sk, d))
        scale = 1.0 / math.sqrt(d)
        out, probs, lse = ops.attention_generic_fwd(qb, kb, vb, scale, want_probs=return_attn, want_lse=True)
        ctx.saved = (qb, kb, vb, out, lse)
        ctx.scale, ctx.dt, ctx.shapes = scale, q.dtype, (q.shape, k.shape, v.shape)
        o = out.view(*lead, Sq, d).to(q.dtype)
        if return_attn:
            probs = probs.view(*lead, Sq, Sk)
            ctx.mark_non_differentiable(probs)
            return o, probs
        return o

    @staticmethod
    def backward(ctx, gout, *_):
        qb, kb, vb, out, lse = _take_saved(ctx)
        go = _as_bf16(gout.reshape(out.shape))
        dq, dk, dv = ops.attention_generic_bwd(qb, kb, vb, out, go, lse, ctx.scale)
        qs, ks, vs = ctx.shapes
        return (dq.view(qs).to(ctx.dt), dk.transpose(1, 2).reshape(ks).to(ctx.dt),
                dv.transpose(1, 2).reshape(vs).to(ctx.dt), None)


def scaled_dot_product_attention(q, k, v, return_attn=False):
    out = _SDPAFn.apply(q, k, v, bool(return_attn))
    if return_attn:
        return out[0], out[1]
    return out, None


# ----------------------------------------------------------------------------------------
# stand-alone LayerNorm (MLPHead.norm, mlp_head.py:9,13): fp32 rows in, bf16 out
# ----------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        D = x.shape[-1]
        x2 = x if (x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32) else x.float().reshape(-1, D).contiguous()
        _, y, mean, rstd = ops.add_layernorm_fwd(x2, None, gamma, beta, eps=eps)
        ctx.saved = (x2, mean, rstd, gamma)
        ctx.in_shape, ctx.in_dtype = x.shape, x.dtype
        return y.view(*x.shape[:-1], D)

    @staticmethod
    def backward(ctx, gout):
        x2, mean, rstd, gamma = _take_saved(ctx)
        dy = _as_bf16(gout.reshape(-1, gout.shape[-1]))
        dx, _, dg, db = ops.add_layernorm_bwd(dy, x2, mean, rstd, gamma, None)
        return dx.view(ctx.in_shape).to(ctx.in_dtype), dg, db, None


def layer_norm(x, ln_module):
    return _LayerNormFn.apply(x, ln_module.weight, ln_module.bias, float(ln_module.eps))


# ----------------------------------------------------------------------------------------
# patch embedding: im2col + projection GEMM + token assembly
# (patch_embedding.py:50-63, 90-96, 122-128; ssl/simmim/model.py:43-49)
# ----------------------------------------------------------------------------------------
def _as_image(img):
    """Images reach the kernels as contiguous fp32 [B,C,H,W] (ToTensor output, what the reference's
    loaders yield) or as raw uint8 bytes, which the kernels scale by 1/255 themselves (§8(f)3)."""
    if img.dtype == torch.uint8:
        return img if img.is_contiguous() else img.contiguous()
    return img if (img.dtype == torch.float32 and img.is_contiguous()) else img.float().contiguous()


class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, meta, weight, bias, cls, pos, mask_token):
        p = meta.patch
        D = weight.shape[0]
        if meta.views is not None:
            # several equally-shaped image batches (DINO's crops of one resolution, model.py:114-115):
            # each is unfolded straight into its slice of ONE patch matrix - no torch.cat of the images
            views = [_as_image(v) for v in meta.views]
            Bv, C, Hh, Ww = views[0].shape
            N = (Hh // p) * (Ww // p)
            B = Bv * len(views)
            patches = torch.empty((B * N, C * p * p), device=views[0].device, dtype=torch.bfloat16)
            for i, v in enumerate(views):
                ops.im2col_bf16(v, p, out=patches[i * Bv * N:(i + 1) * Bv * N])
        else:
            img = _as_image(img)
            B, C, Hh, Ww = img.shape
            N = (Hh // p) * (Ww // p)
            patches = ops.im2col_bf16(img, p)
        proj = ops.gemm(patches, meta.w_bf16, epilogue=EPI_BIAS, bias=bias)
        cls_v = cls.reshape(-1).contiguous() if cls is not None else None
        pos_v = pos.reshape(-1, D).contiguous()
        mt_v = mask_token.reshape(-1).contiguous() if mask_token is not None else None
        x = ops.embed_tokens_fwd(proj, cls_v, pos_v, meta.mask_u8 if mask_token is not None else None, mt_v, B, N, D)
        ctx.saved = (patches,)
        ctx.meta, ctx.dims = meta, (B, N, D)
        ctx.shapes = (weight.shape, None if cls is None else cls.shape, pos.shape,
                      None if mask_token is None else mask_token.shape)
        return x

    @staticmethod
    def backward(ctx, gx):
        (patches,) = _take_saved(ctx)
        meta = ctx.meta
        B, N, D = ctx.dims
        wshape, cshape, pshape, mshape = ctx.shapes
        g = gx if gx.dtype == torch.float32 else gx.float()
        if g.stride(2) != 1:
            g = g.contiguous()
        has_cls = cshape is not None
        dproj, dpos, dmt = ops.embed_tokens_bwd(g, meta.mask_u8 if mshape is not None else None, B, N, D,
                                                has_cls, mshape is not None)
        try:  # bias gradient from the weight-gradient kernel (ones-operand MMA) where the shape allows
            dW, db = ops.gemm_rowsum(dproj, patches, a_mn=True, b_mn=True, split_k=-1)
        except ops._l.VitsslError:
            dW = ops.gemm(dproj, patches, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1)
            db = ops.colsum_bf16(dproj)
        dcls = dpos[0].reshape(cshape).clone() if has_cls else None
        return (None, None, dW.view(wshape), db, dcls, dpos.view(pshape), None if dmt is None else dmt.view(mshape))


def embed_patches(img, owner, weight, bias, cls, pos, patch, mask_u8=None, mask_token=None):
    """[B,C,H,W] image -> fp32 tokens [B, N(+1), D] = (CLS | mask_token | projection) + pos.
    Shape errors surface here as ValueError, like the reference's broadcast failure in
    `x += positional_embedding` (patch_embedding.py:63,95; ssl/simmim/model.py:49) — the kernels
    index `pos` by token and would otherwise read out of bounds."""
    views = None
    if isinstance(img, (list, tuple)):
        views = list(img)
        if not views or any(v.dim() != 4 or v.shape != views[0].shape or v.dtype != views[0].dtype for v in views):
            raise ValueError("a list of views must hold equally-shaped [B,C,H,W] batches of one dtype")
        img = views[0]
        if len(views) == 1:
            views = None
    if img.dim() != 4:
        raise ValueError(f"expected images [B,C,H,W], got {tuple(img.shape)}")
    B, C, Hh, Ww = img.shape
    D = weight.shape[0]
    if Hh % patch or Ww % patch:
        raise ValueError(f"Image dimensions H={Hh}, W={Ww} must be divisible by patch_size={patch}")
    if C * patch * patch != weight.numel() // D:
        raise ValueError(f"patch features C*p*p = {C * patch * patch} do not match the projection ({weight.numel() // D})")
    S = (Hh // patch) * (Ww // patch) + (1 if cls is not None else 0)
    if pos.numel() != S * D:
        raise ValueError(f"positional embedding {tuple(pos.shape)} does not match {S} tokens of width {D} "
                         f"(image {Hh}x{Ww}, patch {patch})")
    if cls is not None and cls.numel() != D:
        raise ValueError(f"cls_token {tuple(cls.shape)} does not have {D} features")
    if mask_token is not None and mask_token.numel() != D:
        raise ValueError(f"mask_token {tuple(mask_token.shape)} does not have {D} features")
    (w_bf16,) = bf16_shadows([(owner, "wproj", [weight])])
    meta = SimpleNamespace(patch=int(patch), w_bf16=w_bf16, mask_u8=mask_u8, views=views)
    return _EmbedFn.apply(img, meta, weight, bias, cls, pos, mask_token)


# ----------------------------------------------------------------------------------------
# bicubic resize of the positional-embedding grid (patch_embedding.py:26-48) as a sparse row
# interpolation: a 16-tap (index, weight) table per output token, built once per grid pair with
# torch's upsample_bicubic2d arithmetic (align_corners=False: src = (dst + 0.5) * in/out - 0.5,
# cubic-convolution coefficients with A = -0.75, border indices clamped). Row 0 (CLS) passes through.
# ----------------------------------------------------------------------------------------
_BICUBIC_TABLES: dict = {}


def bicubic_tables(in_h, in_w, out_h, out_w, device, with_cls=True):
    key = (in_h, in_w, out_h, out_w, str(device), with_cls)
    hit = _BICUBIC_TABLES.get(key)
    if hit is not None:
        return hit
    A = -0.75

    def axis(n_in, n_out):
        scale = n_in / n_out
        idx, wts = [], []
        for o in range(n_out):
            src = scale * (o + 0.5) - 0.5
            f = math.floor(src)
            t = src - f
            c1 = lambda x: ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0
            c2 = lambda x: ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A
            wts.append([c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)])
            idx.append([min(max(f - 1 + k, 0), n_in - 1) for k in range(4)])
        return idx, wts

    iy, wy = axis(in_h, out_h)
    ix, wx = axis(in_w, out_w)
    off = 1 if with_cls else 0
    idx, w = [], []
    if with_cls:
        idx.append([0] * 16)
        w.append([1.0] + [0.0] * 15)
    for oy in range(out_h):
        for ox in range(out_w):
            idx.append([off + iy[oy][a] * in_w + ix[ox][b] for a in range(4) for b in range(4)])
            w.append([wy[oy][a] * wx[ox][b] for a in range(4) for b in range(4)])
    out = (torch.tensor(idx, dtype=torch.int32, device=device), torch.tensor(w, dtype=torch.float32, device=device))
    _BICUBIC_TABLES[key] = out
    return out


class _InterpRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx, w):
        s2 = src if (src.dtype == torch.float32 and src.is_contiguous()) else src.float().contiguous()
        ctx.tables, ctx.n_in, ctx.dt = (idx, w), s2.shape[0], src.dtype
        return ops.interp_rows_fwd(s2, idx, w)

    @staticmethod
    def backward(ctx, g):
        idx, w = ctx.tables
        g = g if (g.dtype == torch.float32 and g.is_contiguous()) else g.float().contiguous()
        return ops.interp_rows_bwd(g, idx, w, ctx.n_in).to(ctx.dt), None, None


def interpolate_pos_embedding(pos, grid_in, grid_out):
    """pos [1, 1 + gh*gw, D] (CLS row first) -> [1, 1 + oh*ow, D], bicubic over the patch grid."""
    idx, w = bicubic_tables(grid_in[0], grid_in[1], grid_out[0], grid_out[1], pos.device)
    return _InterpRowsFn.apply(pos.reshape(-1, pos.shape[-1]), idx, w).unsqueeze(0)


# ----------------------------------------------------------------------------------------
# SimMIM: masked-row gather and fused L1 loss
# ----------------------------------------------------------------------------------------
class _GatherRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, inv_idx):
        B, S, D = x.shape
        x2 = x.reshape(B * S, D)
        ctx.saved = (inv_idx,)
        ctx.shape = (B, S, D)
        return ops.gather_rows_bf16(x2, idx)

    @staticmethod
    def backward(ctx, gy):
        (inv_idx,) = _take_saved(ctx)
        B, S, D = ctx.shape
        return ops.scatter_rows_f32(_as_bf16(gy), inv_idx, B * S).view(B, S, D), None, None


def gather_rows(x, idx, inv_idx):
    return _GatherRowsFn.apply(x, idx, inv_idx)


class _L1LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        pb = _as_bf16(pred)
        tg = target if (target.dtype == torch.float32 and target.is_contiguous()) else target.float().contiguous()
        loss, sign = ops.l1_loss_fwd(pb, tg, want_sign=ctx.needs_input_grad[0])
        ctx.saved = (sign,)
        ctx.n, ctx.dt = pb.numel(), pred.dtype
        return loss

    @staticmethod
    def backward(ctx, go):
        (sign,) = _take_saved(ctx)
        d = ops.l1_loss_bwd(sign, go)  # one pass: sign * (go / n), go read on the device
        return (d if ctx.dt == torch.bfloat16 else d.to(ctx.dt)), None


def l1_loss(pred, target):
    """mean |pred - target| with the sign tensor kept for backward (nn.L1Loss(mean) semantics)."""
    return _L1LossFn.apply(pred, target)


# ----------------------------------------------------------------------------------------
# DINO head tail: F.normalize + weight-normed Linear (ssl/dino/head.py:17,21-22)
# ----------------------------------------------------------------------------------------
class _NormLinearWNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, g, v, bias, out_dtype):
        zb = _as_bf16(z.reshape(-1, z.shape[-1]))
        zn, inv_z = ops.l2norm_fwd(zb)
        w, inv_v = ops.weight_norm_fwd(v.detach().contiguous(), g.detach().reshape(-1).contiguous())
        logits = ops.gemm(zn, w, epilogue=EPI_BIAS, bias=bias, out_dtype=out_dtype)
        need = any(ctx.needs_input_grad)
        if need:
            ctx.saved = (zb, zn, inv_z, w, inv_v, g, v)
            ctx.in_dtype, ctx.in_shape = z.dtype, z.shape
        return logits

    @staticmethod
    def backward(ctx, gl):
        zb, zn, inv_z, w, inv_v, g, v = _take_saved(ctx)
        dl = _as_bf16(gl)
        dzn = ops.gemm(dl, w, b_mn=True)
        dz = ops.l2norm_bwd(zb, inv_z, dzn)
        dW = ops.gemm(dl, zn, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1)
        dbias = ops.colsum_bf16(dl)
        dg, dv = ops.weight_norm_bwd(dW, v.detach().contiguous(), g.detach().reshape(-1).contiguous(), inv_v)
        return dz.view(ctx.in_shape).to(ctx.in_dtype), dg.view(g.shape), dv, dbias, None


def normalize_wn_linear(z, g, v, bias, out_dtype=torch.bfloat16):
    return _NormLinearWNFn.apply(z, g, v, bias, out_dtype)


# ----------------------------------------------------------------------------------------
# DINO loss (ssl/dino/loss.py:13-29) — factorised single-pass kernels
# ----------------------------------------------------------------------------------------
class _DinoLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher, student, center, tt, ts):
        tb, sb = _as_bf16(teacher), _as_bf16(student)
        c = center.reshape(-1).float().contiguous()
        loss, t_stats, s_lse = ops.dino_loss_fwd(tb, sb, c, tt, ts)
        ctx.saved = (tb, sb, c, t_stats, s_lse)
        ctx.temps, ctx.dt = (tt, ts), student.dtype
        return loss

    @staticmethod
    def backward(ctx, go):
        tb, sb, c, t_stats, s_lse = _take_saved(ctx)
        tt, ts = ctx.temps
        ds = ops.dino_loss_bwd(tb, sb, c, t_stats, s_lse, go, tt, ts)
        return None, ds.to(ctx.dt), None, None, None


def dino_loss(teacher, student, center, teacher_temp, student_temp):
    return _DinoLossFn.apply(teacher.detach(), student, center.detach(), float(teacher_temp), float(student_temp))
