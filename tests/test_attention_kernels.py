"""GPU parity of the attention kernels against the oracle's float64 attention
(attention.py:5-27): tcgen05 forward/backward (d_k = 64, S <= 256) and the generic SIMT path."""
import math

import pytest
import torch

from oracle import vit_ref

pytestmark = pytest.mark.gpu


def _ops():
    from vit_core._backend import ops
    return ops


def _ref(qkv, B, S, H):
    D = H * 64
    q, k, v = [t.reshape(B, S, H, 64).transpose(1, 2) for t in qkv.double().split(D, dim=-1)]
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    ctx, probs = vit_ref.scaled_dot_product_attention(q, k, v)
    return q, k, v, ctx, probs


# short sequences are packed several images per 128-row tile (block-diagonal attention): B = 7, S = 37 -> groups of
# 3 with a remainder of 1; B = 5, S = 50 -> pairs; B = 9, S = 17 -> groups of 7; B = 4, S = 64 -> exactly two per tile
@pytest.mark.parametrize("B,S,H", [(2, 196, 6), (3, 197, 6), (2, 37, 3), (1, 256, 2), (2, 128, 1), (5, 16, 4), (2, 144, 6), (1, 1, 1), (2, 129, 2),
                                   (7, 37, 2), (5, 50, 1), (9, 17, 2), (4, 64, 1), (3, 65, 1)])
def test_attention_tcgen05_fwd_bwd(B, S, H):
    ops = _ops()
    D = H * 64
    g = torch.Generator().manual_seed(B * 1000 + S)
    qkv = (torch.randn(B, S, 3 * D, generator=g)).to(torch.bfloat16)
    d_out = (torch.randn(B, S, D, generator=g)).to(torch.bfloat16)
    q, k, v, ctx_ref, _ = _ref(qkv, B, S, H)
    (ctx_ref.transpose(1, 2).reshape(B, S, D) * d_out.double()).sum().backward()

    qkv_c = qkv.cuda()
    qv, kv, vv = qkv_c[..., :D], qkv_c[..., D:2 * D], qkv_c[..., 2 * D:]
    scale = 1.0 / math.sqrt(64)
    out, lse, out_lo = ops.attention_fwd(qv, kv, vv, H, scale)
    ref_ctx = ctx_ref.detach().transpose(1, 2).reshape(B, S, D)
    err = (out.double().cpu() - ref_ctx).abs().max().item()
    assert err < 2e-2 * max(1.0, ref_ctx.abs().max().item()), err
    # log-sum-exp of the scaled scores
    scores = torch.matmul(q.detach(), k.detach().transpose(-2, -1)) * scale
    assert torch.allclose(lse.double().cpu(), torch.logsumexp(scores, dim=-1), atol=1e-3)

    # hi + lo carries the context to ~16 bits (the backward's delta is evaluated from it)
    e_hl = ((out.double() + out_lo.double()).cpu() - ref_ctx).abs().max().item()
    assert e_hl < 1e-2 * max(1.0, ref_ctx.abs().max().item()) and e_hl <= err
    dqkv = torch.empty_like(qkv_c)
    ops.attention_bwd(qv, kv, vv, out, d_out.cuda(), lse, H, scale,
                      dqkv[..., :D], dqkv[..., D:2 * D], dqkv[..., 2 * D:], out_lo=out_lo)
    for name, got, ref in (("dq", dqkv[..., :D], q.grad), ("dk", dqkv[..., D:2 * D], k.grad), ("dv", dqkv[..., 2 * D:], v.grad)):
        ref = ref.transpose(1, 2).reshape(B, S, D)
        e = (got.double().cpu() - ref).abs().max().item()
        assert e < 3e-2 * max(1e-3, ref.abs().max().item()), (name, e, ref.abs().max().item())


def test_attention_cross_lengths_tcgen05():
    ops = _ops()
    B, Sq, Sk, H = 2, 70, 200, 2
    D = H * 64
    g = torch.Generator().manual_seed(3)
    q = torch.randn(B, Sq, D, generator=g).to(torch.bfloat16)
    k = torch.randn(B, Sk, D, generator=g).to(torch.bfloat16)
    v = torch.randn(B, Sk, D, generator=g).to(torch.bfloat16)
    qh, kh, vh = [t.double().reshape(B, -1, H, 64).transpose(1, 2) for t in (q, k, v)]
    ref, _ = vit_ref.scaled_dot_product_attention(qh, kh, vh)
    out, _, _ = ops.attention_fwd(q.cuda(), k.cuda(), v.cuda(), H, 0.125)
    ref = ref.transpose(1, 2).reshape(B, Sq, D)
    assert (out.double().cpu() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,Sq,Sk,H", [(60, 196, 196, 6), (200, 37, 37, 3), (160, 70, 200, 2), (160, 200, 70, 2),
                                        (155, 128, 128, 1), (33, 256, 256, 5)])
def test_attention_bwd_persistent_many_items(B, Sq, Sk, H):
    """More (batch, head) items than SMs: every CTA of the persistent backward kernel walks several
    items, so tile slots are recycled across items (1x1, 1x2, 2x1 and 2x2 tilings, cross lengths)."""
    ops = _ops()
    D = H * 64
    g = torch.Generator().manual_seed(B + Sq + 7 * Sk)
    q = torch.randn(B, Sq, D, generator=g).to(torch.bfloat16)
    k = torch.randn(B, Sk, D, generator=g).to(torch.bfloat16)
    v = torch.randn(B, Sk, D, generator=g).to(torch.bfloat16)
    d_out = torch.randn(B, Sq, D, generator=g).to(torch.bfloat16)
    qh, kh, vh = [t.double().reshape(B, -1, H, 64).transpose(1, 2).requires_grad_(True) for t in (q, k, v)]
    ref, _ = vit_ref.scaled_dot_product_attention(qh, kh, vh)
    (ref.transpose(1, 2).reshape(B, Sq, D) * d_out.double()).sum().backward()
    qc, kc, vc = q.cuda(), k.cuda(), v.cuda()
    out, lse, out_lo = ops.attention_fwd(qc, kc, vc, H, 0.125)
    dq, dk, dv = torch.empty_like(qc), torch.empty_like(kc), torch.empty_like(vc)
    for rep in range(2):  # twice: the second launch must not depend on leftover state
        dq.fill_(7.0); dk.fill_(7.0); dv.fill_(7.0)
        ops.attention_bwd(qc, kc, vc, out, d_out.cuda(), lse, H, 0.125, dq, dk, dv, out_lo=out_lo if rep == 0 else None)
        for name, got, r in (("dq", dq, qh.grad), ("dk", dk, kh.grad), ("dv", dv, vh.grad)):
            r = r.transpose(1, 2).reshape(B, -1, D)
            e = (got.double().cpu() - r).abs().max().item()
            assert e < 3e-2 * max(1e-3, r.abs().max().item()), (name, rep, e, r.abs().max().item())


@pytest.mark.parametrize("S,H", [(37, 3), (197, 6)])
def test_attention_bwd_on_tokens_with_a_common_component(S, H):
    """Real activations: every token's value vector shares a large common part and the upstream
    gradient sits on one row (a CLS-token loss). Then dP - delta cancels to a small difference and
    the accuracy of delta = rowsum(O * dO) decides the accuracy of dQ: taken from the bf16 context
    alone it was ~10x worse than the reference under autocast (4.8e-2 against 5.1e-3 on a late ViT-Ti
    block); from hi + lo it stays within a small factor of it (the rest is the forward's bf16
    probabilities, which the reference's fp32 softmax backward does not see)."""
    ops = _ops()
    B, D = 6, H * 64
    g = torch.Generator().manual_seed(S)
    common = torch.randn(1, 1, 3 * D, generator=g) * 2.0
    qkv = (common + 0.5 * torch.randn(B, S, 3 * D, generator=g)).to(torch.bfloat16)
    d_out = torch.zeros(B, S, D)
    d_out[:, 0] = torch.randn(B, D, generator=g) * 0.02
    d_out = d_out.to(torch.bfloat16)
    q, k, v, ctx_ref, _ = _ref(qkv, B, S, H)
    (ctx_ref.transpose(1, 2).reshape(B, S, D) * d_out.double()).sum().backward()
    qkv_c = qkv.cuda()
    qv, kv, vv = qkv_c[..., :D], qkv_c[..., D:2 * D], qkv_c[..., 2 * D:]
    out, lse, out_lo = ops.attention_fwd(qv, kv, vv, H, 0.125)
    dqkv = torch.empty_like(qkv_c)
    ops.attention_bwd(qv, kv, vv, out, d_out.cuda(), lse, H, 0.125, dqkv[..., :D], dqkv[..., D:2 * D], dqkv[..., 2 * D:], out_lo=out_lo)
    # the reference's arithmetic under autocast on the same inputs (attention.py:20-23), as the yardstick
    qt, kt, vt = (t.view(B, S, H, 64).transpose(1, 2).detach().clone().requires_grad_(True) for t in (qv, kv, vv))
    pr = torch.softmax(((qt @ kt.transpose(-1, -2)) / 8.0).float(), -1)
    (pr.to(torch.bfloat16) @ vt).backward(d_out.cuda().view(B, S, H, 64).transpose(1, 2))

    def rl2(a, b):
        return ((a.double().cpu() - b).norm() / b.norm()).item()
    for name, got, yard, ref in (("dq", dqkv[..., :D], qt.grad, q.grad), ("dk", dqkv[..., D:2 * D], kt.grad, k.grad), ("dv", dqkv[..., 2 * D:], vt.grad, v.grad)):
        ref = ref.transpose(1, 2).reshape(B, S, D)
        e, ey = rl2(got, ref), rl2(yard.transpose(1, 2).reshape(B, S, D), ref)
        assert e <= max(1e-2, 3.0 * ey), (name, e, ey)


@pytest.mark.parametrize("B,H,Sq,Sk,d", [(4, 1, 10, 10, 10), (4, 8, 10, 12, 8), (2, 6, 300, 300, 64), (3, 4, 17, 17, 32)])
def test_attention_generic_fwd_bwd(B, H, Sq, Sk, d):
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    q = torch.randn(B, H, Sq, d, generator=g).to(torch.bfloat16)
    k = torch.randn(B, H, Sk, d, generator=g).to(torch.bfloat16)
    v = torch.randn(B, H, Sk, d, generator=g).to(torch.bfloat16)
    d_out = torch.randn(B, H, Sq, d, generator=g).to(torch.bfloat16)
    qr, kr, vr = [t.double().requires_grad_(True) for t in (q, k, v)]
    ref, probs_ref = vit_ref.scaled_dot_product_attention(qr, kr, vr)
    (ref * d_out.double()).sum().backward()
    scale = 1.0 / math.sqrt(d)
    out, probs, lse = ops.attention_generic_fwd(q.cuda(), k.cuda(), v.cuda(), scale, want_probs=True)
    assert (out.double().cpu() - ref.detach()).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    assert torch.allclose(probs.double().cpu(), probs_ref.detach(), atol=1e-5)
    dq, dk, dv = ops.attention_generic_bwd(q.cuda(), k.cuda(), v.cuda(), out, d_out.cuda(), lse, scale)
    assert (dq.double().cpu() - qr.grad).abs().max().item() < 3e-2 * qr.grad.abs().max().item()
    assert (dk.double().cpu().transpose(1, 2) - kr.grad).abs().max().item() < 3e-2 * kr.grad.abs().max().item()
    assert (dv.double().cpu().transpose(1, 2) - vr.grad).abs().max().item() < 3e-2 * vr.grad.abs().max().item()
