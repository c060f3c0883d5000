"""Developer tool (GPU box): CUDA-event timing of the GEMM shapes of one ViT-S/16 SimMIM step
(B=256 -> M=50176). Inputs rotate through buffers larger than L2. `ONLY=<substr>` filters cases,
`REPS=n` sets repetitions (default 20)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core._backend import ops
from vit_core._backend.ops import EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_GELU_D, EPI_DGELU, EPI_MUL, EPI_NONE

M = int(os.environ.get("M", 50176))
REPS = int(os.environ.get("REPS", 20))
ONLY = os.environ.get("ONLY", "")
P = float(os.environ.get("P", 0.1))
dev = "cuda"
bf = torch.bfloat16


def rnd(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).to(bf)


cases = []  # name, callable factory
D, F = 384, 1536
NB = 3  # rotating buffer sets


def add(name, fn, flops):
    if ONLY and ONLY not in name:
        return
    cases.append((name, fn, flops))


xs = [rnd(M, D) for _ in range(NB)]
hs = [rnd(M, F) for _ in range(NB)]
us = [torch.empty(M, F, device=dev, dtype=bf) for _ in range(NB)]
qkvs = [rnd(M, 3 * D) for _ in range(NB)]
w1, w2, wqkv, wo = rnd(F, D), rnd(D, F), rnd(3 * D, D), rnd(D, D)
b1, b2 = torch.randn(F, device=dev), torch.randn(D, device=dev)

add("ffn1 fwd gelu  MxFxD", lambda i: ops.gemm(xs[i], w1, epilogue=EPI_BIAS_GELU, bias=b1, aux=us[i], dropout_p=P, seed=1, offset=1), 2.0 * M * F * D)
add("ffn1 fwd gelu_d MxFxD", lambda i: ops.gemm(xs[i], w1, epilogue=EPI_BIAS_GELU_D, bias=b1, aux=us[i], dropout_p=P, seed=1, offset=1), 2.0 * M * F * D)
add("ffn2 dgrad mul MxFxD", lambda i: ops.gemm(xs[i], w2, b_mn=True, epilogue=EPI_MUL, aux=hs[i]), 2.0 * M * F * D)
add("ffn1 fwd gelu p=0", lambda i: ops.gemm(xs[i], w1, epilogue=EPI_BIAS_GELU, bias=b1, aux=us[i]), 2.0 * M * F * D)
add("ffn1 fwd bias  MxFxD", lambda i: ops.gemm(xs[i], w1, epilogue=EPI_BIAS, bias=b1), 2.0 * M * F * D)
add("ffn1 fwd none  MxFxD", lambda i: ops.gemm(xs[i], w1), 2.0 * M * F * D)
add("ffn2 fwd bias  MxDxF", lambda i: ops.gemm(hs[i], w2, epilogue=EPI_BIAS, bias=b2), 2.0 * M * F * D)
add("ffn2 dgrad dgelu MxFxD", lambda i: ops.gemm(xs[i], w2, b_mn=True, epilogue=EPI_DGELU, aux=hs[i], dropout_p=P, seed=1, offset=1), 2.0 * M * F * D)
add("ffn2 dgrad dgelu p=0", lambda i: ops.gemm(xs[i], w2, b_mn=True, epilogue=EPI_DGELU, aux=hs[i]), 2.0 * M * F * D)
add("ffn1 dgrad none MxDxF", lambda i: ops.gemm(hs[i], w1, b_mn=True), 2.0 * M * F * D)
add("ffn wgrad FxD (k=M)", lambda i: ops.gemm(hs[i], xs[i], a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1), 2.0 * M * F * D)
add("ffn wgrad DxF (k=M)", lambda i: ops.gemm(xs[i], hs[i], a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1), 2.0 * M * F * D)
add("qkv fwd Mx3DxD", lambda i: ops.gemm(xs[i], wqkv), 2.0 * M * 3 * D * D)
add("qkv dgrad MxDx3D", lambda i: ops.gemm(qkvs[i], wqkv, b_mn=True), 2.0 * M * 3 * D * D)
add("qkv wgrad 3DxD", lambda i: ops.gemm(qkvs[i], xs[i], a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1), 2.0 * M * 3 * D * D)
add("out fwd MxDxD", lambda i: ops.gemm(xs[i], wo), 2.0 * M * D * D)
add("out dgrad MxDxD", lambda i: ops.gemm(xs[i], wo, b_mn=True), 2.0 * M * D * D)
add("out wgrad DxD", lambda i: ops.gemm(xs[i], xs[(i + 1) % NB], a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=-1), 2.0 * M * D * D)

for name, fn, flops in cases:
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
    print(f"{name:28s} {us_:8.1f} us  {flops / us_ * 1e-6:8.1f} TFLOP/s", flush=True)
