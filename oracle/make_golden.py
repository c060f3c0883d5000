"""Generate tests/golden/*.pt by running the REAL reference (kristi700/ViT-SSL, /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
Every case: seed -> construct the reference module -> float64 forward/backward on CPU with
dropout = 0 -> save weights (fp32), inputs, outputs, loss and gradients (fp64). The committed
fixtures pin both the oracle restatement (tests/test_oracle_golden.py) and the CUDA path
(tests/test_models.py).
"""
import os
import sys

REF = "/root/reference"
assert os.path.isdir(REF), "reference checkout not found"
sys.path = [p for p in sys.path if "vit-ssl_b200" not in p]
sys.path.insert(0, REF)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import vit_core  # noqa: E402  (the reference package)
from vit_core.attention import MultiHeadedAttention  # noqa: E402
from vit_core.encoder_block import EncoderBlock  # noqa: E402
from vit_core.feed_forward import FeedForwardBlock  # noqa: E402
from vit_core.patch_embedding import (ConvolutionalPatchEmbedding, DynamicPatchEmbedding,  # noqa: E402
                                      ManualPatchEmbedding)
from vit_core.ssl.dino.loss import DINOLoss  # noqa: E402
from vit_core.ssl.dino.model import DINOViT  # noqa: E402
from vit_core.ssl.simmim.model import SimMIMViT  # noqa: E402
from vit_core.vit import ViT  # noqa: E402

assert vit_core.__file__.startswith(REF), vit_core.__file__
sys.path.insert(1, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.cases import build_dino_case, digest  # noqa: E402
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def sd32(m):
    return {k: v.detach().clone().float() for k, v in m.state_dict().items()}


def grads64(m):
    # stored as fp32 to keep the fixtures small (values computed in fp64)
    return {k: p.grad.detach().float() for k, p in m.named_parameters() if p.grad is not None}


def save(name, obj):
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def case_encoder_block():
    torch.manual_seed(101)
    m = EncoderBlock(d_model=128, num_heads=2, mlp_dim=256, dropout=0.0)
    w = sd32(m)
    m = m.double()
    x = torch.rand(2, 37, 128).double().requires_grad_(True)
    dy = torch.randn(2, 37, 128).double()
    y, probs = m(x, return_attn=True)
    (y * dy).sum().backward()
    save("encoder_block", dict(cfg=dict(d_model=128, num_heads=2, mlp_dim=256), weights=w, x=x.detach().float(),
                               dy=dy.float(), y=y.detach(), probs=probs.detach(), dx=x.grad.clone(), grads=grads64(m)))


def case_mha_cross():
    torch.manual_seed(102)
    m = MultiHeadedAttention(d_model=128, num_heads=2)
    w = sd32(m)
    m = m.double()
    q = torch.randn(2, 10, 128).double().requires_grad_(True)
    k = torch.randn(2, 12, 128).double().requires_grad_(True)
    v = torch.randn(2, 12, 128).double().requires_grad_(True)
    out, probs = m(q, k, v, return_attn=True)
    dy = torch.randn(2, 10, 128).double()
    (out * dy).sum().backward()
    save("mha_cross", dict(weights=w, q=q.detach().float(), k=k.detach().float(), v=v.detach().float(), dy=dy.float(),
                           out=out.detach(), probs=probs.detach(), dq=q.grad.clone(), dk=k.grad.clone(),
                           dv=v.grad.clone(), grads=grads64(m)))


def case_ffn():
    torch.manual_seed(103)
    m = FeedForwardBlock(d_model=64, d_ff=128, dropout=0.0)
    w = sd32(m)
    m = m.double()
    x = torch.randn(4, 10, 64).double().requires_grad_(True)
    y = m(x)
    dy = torch.randn(4, 10, 64).double()
    (y * dy).sum().backward()
    save("ffn", dict(weights=w, x=x.detach().float(), dy=dy.float(), y=y.detach(), dx=x.grad.clone(), grads=grads64(m)))


def case_patch_embeddings():
    torch.manual_seed(104)
    x = torch.rand(3, 3, 32, 32)
    out = {}
    for name, cls in (("conv", ConvolutionalPatchEmbedding), ("manual", ManualPatchEmbedding)):
        m = cls((3, 32, 32), 64, 8)
        w = sd32(m)
        y = m.double()(x.double())
        out[name] = dict(weights=w, y=y.detach())
    m = DynamicPatchEmbedding((3, 32, 32), 64, 8)
    w = sd32(m)
    md = m.double()
    x16 = torch.rand(3, 3, 16, 16)
    out["dynamic"] = dict(weights=w, y=md(x.double()).detach(), x_local=x16, y_local=md(x16.double()).detach())
    out["x"] = x
    save("patch_embeddings", out)


def case_vit():
    torch.manual_seed(105)
    cfg = dict(num_classes=10, num_blocks=2, input_shape=(3, 32, 32), embed_dim=128, patch_size=8, num_heads=2,
               mlp_dim=256, dropout=0.0)
    m = ViT(**cfg)
    w = sd32(m)
    m = m.double()
    x = torch.rand(4, 3, 32, 32)
    labels = torch.randint(0, 10, (4,))
    logits, probs = m(x.double(), return_attn=True)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    save("vit", dict(cfg=cfg, weights=w, x=x, labels=labels, logits=logits.detach(), probs=probs.detach(),
                     loss=loss.detach(), grads=grads64(m)))


def case_simmim():
    torch.manual_seed(106)
    cfg = dict(num_blocks=2, input_shape=(3, 32, 32), embed_dim=128, patch_size=8, num_heads=2, mlp_dim=256,
               dropout=0.0, mask_ratio=0.6)
    m = SimMIMViT(**cfg)
    w = sd32(m)
    m = m.double()
    x = torch.rand(5, 3, 32, 32)
    # replay the generator to record the permutations the reference draws (masking.py:22-25)
    state = torch.get_rng_state()
    perms = torch.stack([torch.randperm(16) for _ in range(5)])
    torch.set_rng_state(state)
    pred, targets, bool_mask = m(x.double(), return_bool_mask=True)
    loss = torch.nn.L1Loss()(pred, targets)
    loss.backward()
    feats = m.inference_forward(x.double())
    save("simmim", dict(cfg=cfg, weights=w, x=x, perms=perms, bool_mask=bool_mask.squeeze(-1).clone(),
                        pred=pred.detach(), targets=targets.detach(), loss=loss.detach(), grads=grads64(m),
                        inference=feats.detach()))


def case_dino():
    cfg, m, views, B = build_dino_case(DINOViT)
    wdig = {k: digest(v) for k, v in m.state_dict().items()}
    m = m.double()
    teacher, student = m([v.double() for v in views], 2)
    K = teacher.shape[1]
    crit = DINOLoss(0.04, 0.1)
    loss = crit(teacher.view(2, B, K), student.view(4, B, K), m.center)
    loss.backward()
    grads = {k: digest(p.grad) for k, p in m.named_parameters() if p.grad is not None}
    center_after = m.center.detach().clone()
    m.momentum_update_teacher(0.996)
    teacher_after = {k: digest(v) for k, v in m.state_dict().items() if k.startswith("teacher_")}
    feats = m.inference_forward(views[0].double())
    save("dino", dict(cfg=cfg, weight_digests=wdig, teacher=teacher.detach(), student=student.detach(),
                      center_after=center_after, loss=loss.detach(), grad_digests=grads,
                      teacher_after_digests=teacher_after, inference=feats.detach(), temps=(0.04, 0.1),
                      momentum=0.996))


def case_dino_loss():
    torch.manual_seed(108)
    G, V, B, K = 2, 8, 4, 1024
    teacher = (torch.randn(G, B, K) * 2).double()
    student = (torch.randn(V, B, K) * 2).double().requires_grad_(True)
    center = torch.randn(1, K).double() * 0.1
    loss = DINOLoss(0.05, 0.1)(teacher, student, center)
    loss.backward()
    save("dino_loss", dict(teacher=teacher.float(), student=student.detach().float(), center=center.float(),
                           loss=loss.detach(), dstudent=student.grad.clone(), temps=(0.05, 0.1)))


if __name__ == "__main__":
    for fn in (case_encoder_block, case_mha_cross, case_ffn, case_patch_embeddings, case_vit, case_simmim,
               case_dino, case_dino_loss):
        fn()
