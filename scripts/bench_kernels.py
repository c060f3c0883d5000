"""Developer tool (GPU box): CUDA-event timing of the non-GEMM hot kernels at the ViT-S/16 SimMIM
B=256 shapes (attention fwd/bwd, fused add+LayerNorm fwd/bwd). `ONLY=<substr>`, `REPS=n`."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core._backend import ops

B = int(os.environ.get("B", 256))
S = int(os.environ.get("S", 196))
H, D = int(os.environ.get("H", 6)), int(os.environ.get("D", 384))
REPS = int(os.environ.get("REPS", 20))
ONLY = os.environ.get("ONLY", "")
P = float(os.environ.get("P", 0.1))
dev, bf = "cuda", torch.bfloat16
NB = 3
M = B * S
scale = 1.0 / math.sqrt(64)

qkv = [(torch.randn(B, S, 3 * D, device=dev) * 0.5).to(bf) for _ in range(NB)]
dctx = [(torch.randn(B, S, D, device=dev) * 0.5).to(bf) for _ in range(NB)]
dqkv = [torch.empty(B, S, 3 * D, device=dev, dtype=bf) for _ in range(NB)]
ctxs, lses, los = [], [], []
for i in range(NB):
    c, l, lo = ops.attention_fwd(qkv[i][..., :D], qkv[i][..., D:2 * D], qkv[i][..., 2 * D:], H, scale)
    ctxs.append(c); lses.append(l); los.append(lo)
xs = [torch.randn(M, D, device=dev) for _ in range(NB)]
br = [(torch.randn(M, D, device=dev)).to(bf) for _ in range(NB)]
g, be = torch.randn(D, device=dev), torch.randn(D, device=dev)
lnout = [ops.add_layernorm_fwd(xs[i], br[i], g, be, dropout_p=P, seed=1, offset=0) for i in range(NB)]

cases = []


def add(name, fn, work, unit):
    if ONLY and ONLY not in name:
        return
    cases.append((name, fn, work, unit))


attn_flops = 4.0 * B * H * S * S * 64
add("attn fwd", lambda i: ops.attention_fwd(qkv[i][..., :D], qkv[i][..., D:2 * D], qkv[i][..., 2 * D:], H, scale), attn_flops, "TFLOP/s")
add("attn bwd", lambda i: ops.attention_bwd(qkv[i][..., :D], qkv[i][..., D:2 * D], qkv[i][..., 2 * D:], ctxs[i], dctx[i], lses[i], H, scale,
                                             dqkv[i][..., :D], dqkv[i][..., D:2 * D], dqkv[i][..., 2 * D:], out_lo=los[i]), 2.5 * attn_flops, "TFLOP/s")
add("ln fwd add+ln", lambda i: ops.add_layernorm_fwd(xs[i], br[i], g, be, dropout_p=P, seed=1, offset=0), M * D * 12.0, "GB/s")
add("ln fwd ln only", lambda i: ops.add_layernorm_fwd(xs[i], None, g, be), M * D * 6.0, "GB/s")
add("ln bwd full", lambda i: ops.add_layernorm_bwd(br[i], lnout[i][0], lnout[i][2], lnout[i][3], g, xs[i], want_dbranch=True,
                                                   dropout_p=P, seed=1, offset=0), M * D * 16.0, "GB/s")
add("colsum", lambda i: ops.colsum_bf16(br[i]), M * D * 2.0, "GB/s")

for name, fn, work, unit in cases:
    for i in range(3):
        fn(i % NB)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REPS):
        fn(i % NB)
    e1.record()
    torch.cuda.synchronize()
    us_ = e0.elapsed_time(e1) / REPS * 1e3
    rate = work / us_ * (1e-6 if unit == "TFLOP/s" else 1e-3)
    print(f"{name:20s} {us_:8.1f} us  {rate:8.1f} {unit}", flush=True)
