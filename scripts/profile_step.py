"""Developer tool (GPU box): per-kernel breakdown (torch.profiler / CUPTI) of one training step of the
bench workloads through the drop-in API. WORKLOAD=simmim|dino, OPT=vitssl|torch, B=<batch>.
Writes gpurun_out/prof_<workload>.txt. Not a benchmark: numbers under a profiler are for attribution."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core.optim import FusedAdamW

WL = os.environ.get("WORKLOAD", "simmim")
OPT = os.environ.get("OPT", "vitssl")
B = int(os.environ.get("B", 256 if WL == "simmim" else 128))
arch = dict(embed_dim=384, num_blocks=12, num_heads=6, mlp_dim=1536, patch_size=16)
torch.manual_seed(0)
if WL == "simmim":
    from vit_core.ssl.simmim import SimMIMViT
    m = SimMIMViT(input_shape=(3, 224, 224), dropout=0.1, mask_ratio=0.6, **arch).cuda().train()
    crit = torch.nn.L1Loss()
    batch = [torch.rand(B, 3, 224, 224, device="cuda")]
else:
    from vit_core.ssl.dino import DINOViT
    from vit_core.ssl.dino.loss import DINOLoss
    m = DINOViT(input_shape=(3, 224, 224), dropout=0.1, output_dim=65536, center_momentum=0.9, **arch).cuda().train()
    crit = DINOLoss(0.04, 0.1)
    batch = [torch.rand(B, 3, 224, 224, device="cuda") for _ in range(2)] + [torch.rand(B, 3, 96, 96, device="cuda") for _ in range(6)]
params = [p for p in m.parameters() if p.requires_grad]
opt = FusedAdamW(params, lr=1e-4, weight_decay=1e-3) if OPT == "vitssl" else torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-3, fused=True)
scaler = torch.amp.GradScaler("cuda")


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        if WL == "simmim":
            pred, tgt = m(batch[0])
            loss = crit(pred, tgt)
        else:
            t, s = m(batch, 2)
            loss = crit(t.view(2, B, -1), s.view(8, B, -1), m.center)
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    if WL == "dino":
        m.momentum_update_teacher(0.996)
    return loss


for _ in range(4):
    step()
torch.cuda.synchronize()
for rep in range(3):  # wall time of individual steps: catches one-off stalls (allocator growth, lazy loads)
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"step {rep}: host issue {1e3 * (t1 - t0):.2f} ms, wall {1e3 * (t2 - t0):.2f} ms", flush=True)

from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = os.path.join(ROOT, "gpurun_out", f"prof_{WL}_{OPT}.txt")
with open(out, "w") as f:
    f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    f.write("\n\n==== by host time ====\n")
    f.write(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
print(open(out).read()[:200])
