"""Developer tool (GPU box): SimMIM pretraining step of another architecture (BASELINE.json
configs[3]: ViT-B/16, batch 128 per GPU) with the same step body as bench.py; CUDA-event timing.
ARCH=vit_b|vit_s, B=<batch>."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core.ssl.simmim import SimMIMViT

ARCH = os.environ.get("ARCH", "vit_b")
D, L, H, F_ = {"vit_b": (768, 12, 12, 3072), "vit_s": (384, 12, 6, 1536)}[ARCH]
B = int(os.environ.get("B", 128))
torch.manual_seed(0)
m = SimMIMViT(num_blocks=L, input_shape=(3, 224, 224), embed_dim=D, patch_size=16, num_heads=H, mlp_dim=F_, dropout=0.1,
              mask_ratio=0.6).cuda().train()
opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-3, fused=True)
scaler = torch.amp.GradScaler("cuda")
xs = [torch.rand(B, 3, 224, 224, device="cuda") for _ in range(3)]


def step(x):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = m.reconstruction_loss(x)
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    return loss


for i in range(4):
    step(xs[i % 3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for i in range(n):
    loss = step(xs[i % 3])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
N, P = 196, 768
n_m = int(N * 0.6)
blk = 8 * N * D * D + 4 * N * N * D + 4 * N * D * F_
flops = 3 * (L * blk + 2 * n_m * D * P) + 2 * (2 * N * P * D)
print(f"SimMIM {ARCH} B={B}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} img/s  {B / ms * 1e3 * flops / 1e12:.0f} TFLOP/s model "
      f"({B / ms * 1e3 * flops / 1e12 / 1387.4:.3f} of sustained bf16 peak)  loss {loss.item():.4f}")
