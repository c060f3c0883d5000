"""Developer tool (GPU box): how far are the attention gradients of (a) the tcgen05 kernels, (b) the
generic SIMT kernels and (c) plain torch ops composed the way autocast(bf16) runs the reference
(attention.py:20-23) from a float64 evaluation of the same bf16 inputs? Prints relative L2 errors."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core._backend import ops

torch.manual_seed(0)
bf = torch.bfloat16


def rl2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


for (B, H, S, std) in [(8, 3, 37, 1.0), (8, 3, 37, 0.3), (12, 6, 37, 1.0), (4, 6, 196, 1.0), (4, 6, 197, 1.0), (4, 6, 197, 0.3)]:
    D = H * 64
    qkv = (torch.randn(B, S, 3 * D, device="cuda") * std).to(bf)
    do = (torch.randn(B, S, D, device="cuda") * 0.5).to(bf)
    scale = 1.0 / 8.0
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    # float64 truth on the same bf16 inputs
    qd, kd, vd = (t.double().view(B, S, H, 64).transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v))
    p = torch.softmax(qd @ kd.transpose(-1, -2) * scale, -1)
    od = p @ vd
    od.backward(do.double().view(B, S, H, 64).transpose(1, 2))
    ref = [od.transpose(1, 2).reshape(B, S, D)] + [t.grad.transpose(1, 2).reshape(B, S, D) for t in (qd, kd, vd)]
    # (a) tcgen05 kernels
    ctx, lse, ctx_lo = ops.attention_fwd(q, k, v, H, scale)
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(q, k, v, ctx, do, lse, H, scale, dqkv[..., :D], dqkv[..., D:2 * D], dqkv[..., 2 * D:], out_lo=ctx_lo)
    a = [ctx, dqkv[..., :D], dqkv[..., D:2 * D], dqkv[..., 2 * D:]]
    # (b) generic kernels
    qh, kh, vh = (t.unflatten(2, (H, 64)).transpose(1, 2) for t in (q, k, v))
    og, _, lg = ops.attention_generic_fwd(qh, kh, vh, scale, want_probs=False, want_lse=True)
    gq, gk, gv = ops.attention_generic_bwd(qh, kh, vh, og, do.unflatten(2, (H, 64)).transpose(1, 2), lg, scale)
    b = [og.transpose(1, 2).reshape(B, S, D), gq.transpose(1, 2).reshape(B, S, D), gk.reshape(B, S, D), gv.reshape(B, S, D)]
    # (c) torch ops as autocast runs the reference: bf16 matmuls, fp32 softmax, bf16 probabilities into PV
    qt, kt, vt = (t.view(B, S, H, 64).transpose(1, 2).detach().clone().requires_grad_(True) for t in (q, k, v))
    sc = (qt @ kt.transpose(-1, -2)) / math.sqrt(64)
    pr = torch.softmax(sc.float(), -1)
    ot = pr.to(bf) @ vt
    ot.backward(do.view(B, S, H, 64).transpose(1, 2))
    c = [ot.transpose(1, 2).reshape(B, S, D)] + [t.grad.transpose(1, 2).reshape(B, S, D) for t in (qt, kt, vt)]
    print(f"B={B} H={H} S={S} std={std}")
    for name, got in (("tcgen05", a), ("generic", b), ("torch-autocast", c)):
        print(f"   {name:15s} out {rl2(got[0], ref[0]):.2e}  dq {rl2(got[1], ref[1]):.2e}  dk {rl2(got[2], ref[2]):.2e}  dv {rl2(got[3], ref[3]):.2e}")
