"""Developer tool (GPU box): where does the GPU idle inside a step? Runs the SimMIM (or DINO) step of the
bench through the drop-in API under torch.profiler, with and without the per-step `loss.item()` the
trainers do, and lists the largest gaps between consecutive kernels with their neighbours.
WORKLOAD=simmim|dino. Numbers under a profiler are for attribution only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
from vit_core.optim import FusedAdamW

WL = os.environ.get("WORKLOAD", "simmim")
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
opt = FusedAdamW(params, lr=1e-4, weight_decay=1e-3)
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


def timed(n, sync_each):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        l = step()
        if sync_each:
            l.item()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for _ in range(5):
    step()
print(f"ms/step free-running {timed(20, False):.3f}   with loss.item() every step {timed(20, True):.3f}", flush=True)

from torch.profiler import ProfilerActivity, profile
for sync_each in (False, True):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            l = step()
            if sync_each:
                l.item()
        torch.cuda.synchronize()
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                 if str(e.device_type).endswith("CUDA") and e.time_range.end > e.time_range.start), key=lambda x: x[0])
    busy = sum(e - s for s, e, _ in ks)
    span = ks[-1][1] - ks[0][0]
    gaps = [(ks[i + 1][0] - ks[i][1], ks[i][2][:60], ks[i + 1][2][:60]) for i in range(len(ks) - 1)]
    tot_gap = sum(g for g, _, _ in gaps if g > 0)
    print(f"\n== sync_each={sync_each}: {len(ks)} device ops over 3 steps, span {span / 3e3:.3f} ms/step, busy {busy / 3e3:.3f}, gaps {tot_gap / 3e3:.3f}")
    small = sum(g for g, _, _ in gaps if 0 < g <= 5)
    print(f"   gaps <= 5 us: {small / 3e3:.3f} ms/step ({sum(1 for g, _, _ in gaps if 0 < g <= 5) // 3} per step); larger ones:")
    for g, a, b in sorted(gaps, reverse=True)[:18]:
        print(f"   {g:8.1f} us   after {a:<60s} before {b}")
