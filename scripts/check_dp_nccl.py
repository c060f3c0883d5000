"""GPU box, torchrun --nproc-per-node N: data-parallel SimMIM step over NCCL. Checks that the
gradients every rank ends up with equal the mean of the per-rank single-GPU gradients (the chunked
stack backward + in-node all-reduce path), and that they are identical across ranks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-ssl_b200"))
import torch
import torch.distributed as dist
from vit_core._backend import dp
from vit_core.ssl.simmim import SimMIMViT

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = dict(num_blocks=6, input_shape=(3, 64, 64), embed_dim=128, patch_size=8, num_heads=2, mlp_dim=512, dropout=0.0,
           mask_ratio=0.6)
torch.manual_seed(0)
model = SimMIMViT(**cfg).cuda().train()
torch.manual_seed(100 + rank)
x = torch.rand(8, 3, 64, 64, device="cuda")


def grads(m, seed):
    m.zero_grad(set_to_none=True)
    torch.cuda.manual_seed(seed)  # same mask draw in both runs
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = m.reconstruction_loss(x)
    (loss * 1024.0).backward()
    torch.cuda.synchronize()
    return [p.grad.detach().clone() for p in m.parameters()]


local_g = grads(model, 5 + rank)          # before attaching: plain single-GPU gradients of this rank's shard
want = []
for g in local_g:
    t = g.clone()
    dist.all_reduce(t)
    want.append(t / world)
sync = dp.attach(model)
got = grads(model, 5 + rank)
assert sync.reduced_bytes > 0
worst = 0.0
for a, b in zip(got, want):
    worst = max(worst, ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item())
chk = torch.stack([g.double().sum() for g in got])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(c, allc[0]) for c in allc)
if rank == 0:
    print(f"dp check (SimMIM): world {world}, worst rel diff vs mean of local grads {worst:.2e}, identical across ranks: {same}")
    assert worst < 1e-4 and same

# ---- DINO: the student stack runs twice per step (global and local crops share parameters), the
# teacher is frozen, the center is averaged across ranks inside the forward
from vit_core.ssl.dino import DINOViT
from vit_core.ssl.dino.loss import DINOLoss

torch.manual_seed(1)
dcfg = dict(num_blocks=4, input_shape=(3, 64, 64), embed_dim=128, patch_size=8, num_heads=2, mlp_dim=256, dropout=0.0,
            output_dim=512, center_momentum=0.9)
dino = DINOViT(**dcfg).cuda().train()
crit = DINOLoss(0.04, 0.1)
torch.manual_seed(200 + rank)
Bd = 4
views = [torch.rand(Bd, 3, 64, 64, device="cuda") for _ in range(2)] + [torch.rand(Bd, 3, 32, 32, device="cuda") for _ in range(4)]
center0 = dino.center.clone()


def dino_grads(m):
    m.zero_grad(set_to_none=True)
    m.center = center0.clone()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        t, s_ = m(views, 2)
        # the same (pre-update) center on every rank and in both runs, so only the gradient path differs
        loss = crit(t.view(2, Bd, -1), s_.view(6, Bd, -1), center0)
    (loss * 1024.0).backward()
    torch.cuda.synchronize()
    return [p.grad.detach().clone() for p in m.parameters() if p.requires_grad]


import vit_core._backend.dp as _dp  # keep the first run local: no auto-attach
_maybe = _dp.maybe_attach
_dp.maybe_attach = lambda module: None
local_g = dino_grads(dino)
want = []
for g_ in local_g:
    t_ = g_.clone()
    dist.all_reduce(t_)
    want.append(t_ / world)
_dp.maybe_attach = _maybe
sync = dp.attach(dino)
got = dino_grads(dino)
worst = 0.0
for a, b in zip(got, want):
    worst = max(worst, ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item())
chk = torch.stack([g_.double().sum() for g_ in got])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(c, allc[0]) for c in allc)
if rank == 0:
    print(f"dp check (DINO, 2 global + 4 local crops): worst rel diff {worst:.2e}, identical across ranks: {same}")
    assert worst < 1e-3 and same
dist.destroy_process_group()
