"""Developer tool (GPU box): one DINO ViT-S/16 pretraining step (BASELINE.json configs[2] shape:
2x224 global + 6x96 local crops, K=65536 head, EMA teacher) timed with CUDA events, plus a
torch.profiler kernel table. B per GPU via env B (default 64)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core.ssl.dino import DINOViT
from vit_core.ssl.dino.loss import DINOLoss

B = int(os.environ.get("B", 64))
K = int(os.environ.get("K", 65536))
torch.manual_seed(0)
m = DINOViT(num_blocks=12, input_shape=(3, 224, 224), embed_dim=384, patch_size=16, num_heads=6, mlp_dim=1536,
            dropout=0.1, output_dim=K, center_momentum=0.9).cuda().train()
crit = DINOLoss(0.04, 0.1)
opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-3, fused=True)
scaler = torch.amp.GradScaler("cuda")
views = [torch.rand(B, 3, 224, 224, device="cuda") for _ in range(2)] + [torch.rand(B, 3, 96, 96, device="cuda") for _ in range(6)]


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        t, s = m(views, 2)
        loss = crit(t.view(2, B, K), s.view(8, B, K), m.center)
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    m.momentum_update_teacher(0.996)
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"DINO ViT-S/16 B={B} K={K}: {ms:.2f} ms/step  {B / ms * 1e3:.1f} img/s  loss {loss.item():.5f}  "
      f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "prof_dino.txt"), "w") as f:
    f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=80))
