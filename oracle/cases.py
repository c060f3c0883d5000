"""ORACLE support — deterministic test-case construction shared by `oracle/make_golden.py`
(run with the REAL reference classes) and `tests/` (run with this repo's classes). Test
infrastructure only; never imported by the product."""
import torch


def digest(t, n=256):
    """Compact fingerprint of a large tensor: shape, fp64 norm, sum and a fixed strided sample."""
    f = t.detach().double().reshape(-1)
    step = max(1, f.numel() // n)
    return dict(shape=tuple(t.shape), norm=f.norm().item(), sum=f.sum().item(), step=step,
                sample=f[::step][:n].clone())



def build_dino_case(DINOViT_cls):
    """Deterministic construction shared with tests/test_models.py (which calls it with OUR DINOViT):
    same seed + same construction order => identical fp32 weights, so the 37 MB of head weights
    need not be stored; the fixture keeps per-tensor digests to prove the weights matched."""
    torch.manual_seed(107)
    cfg = dict(num_blocks=2, input_shape=(3, 32, 32), embed_dim=128, patch_size=8, num_heads=2, mlp_dim=256,
               dropout=0.0, output_dim=512, center_momentum=0.9)
    m = DINOViT_cls(**cfg)
    with torch.no_grad():  # make teacher != student and center != 0 so the case is not degenerate
        for p in m.teacher_backbone.parameters():
            p.add_(0.01 * torch.randn_like(p))
        for p in m.teacher_head.parameters():
            p.add_(0.01 * torch.randn_like(p))
        m.center.copy_(0.1 * torch.randn_like(m.center))
    B = 3
    views = [torch.rand(B, 3, 32, 32) for _ in range(2)] + [torch.rand(B, 3, 16, 16) for _ in range(2)]
    return cfg, m, views, B


