"""Developer tool (GPU box): per-kernel breakdown of one SimMIM ViT-S/16 training step.
Writes gpurun_out/prof_table.txt (torch.profiler, CUPTI) and gpurun_out/gemm_shapes.txt (CUDA-event
timing per GEMM shape from the ops.PROFILE hook). Not a benchmark: numbers taken under a profiler
are for attribution only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core._backend import ops
from vit_core.ssl.simmim import SimMIMViT

B = int(os.environ.get("B", 256))
torch.manual_seed(0)
m = SimMIMViT(num_blocks=12, input_shape=(3, 224, 224), embed_dim=384, patch_size=16, num_heads=6, mlp_dim=1536,
              dropout=0.1, mask_ratio=0.6).cuda().train()
opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-3, fused=True)
scaler = torch.amp.GradScaler("cuda")
x = torch.rand(B, 3, 224, 224, device="cuda")


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = m.reconstruction_loss(x)
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()


for _ in range(3):
    step()
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

import time
t0 = time.perf_counter()
for _ in range(5):
    step()
t_cpu = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 5

ops.PROFILE = []
step()
torch.cuda.synchronize()
recs, ops.PROFILE = ops.PROFILE, None
agg = {}
for k, a, b, w in recs:
    t = a.elapsed_time(b)
    c = agg.setdefault(k, [0, 0.0, 0.0])
    c[0] += 1; c[1] += t; c[2] += w
with open(os.path.join(ROOT, "gpurun_out", "gemm_shapes.txt"), "w") as f:
    f.write(f"cpu launch time per step {t_cpu*1e3:.2f} ms, wall per step {t_all*1e3:.2f} ms\n")
    for k, (n, t, w) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        unit = "TFLOP/s" if k.startswith("gemm") else "GB/s"
        rate = w / (t * 1e-3) / (1e12 if k.startswith("gemm") else 1e9)
        f.write(f"{k:60s} n={n:3d} total={t:8.3f} ms  avg={t/n*1e3:8.1f} us  {rate:8.1f} {unit}\n")

from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
with open(os.path.join(ROOT, "gpurun_out", "prof_table.txt"), "w") as f:
    f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
print(open(os.path.join(ROOT, "gpurun_out", "gemm_shapes.txt")).read())
