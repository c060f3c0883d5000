"""GPU parity of the fused residual-add + LayerNorm kernels (C-ABI vitssl_add_layernorm_{fwd,bwd})
against the oracle's float64 LayerNorm (encoder_block.py:40-52)."""
import pytest
import torch

from oracle import vit_ref

pytestmark = pytest.mark.gpu


def _ops():
    from vit_core._backend import ops
    return ops


@pytest.mark.parametrize("rows,D", [(4 * 196, 384), (300, 768), (37, 192), (64, 128), (50, 64), (9, 1024), (33, 100)])
@pytest.mark.parametrize("with_branch", [False, True])
def test_add_layernorm_fwd_bwd(rows, D, with_branch):
    ops = _ops()
    g = torch.Generator().manual_seed(rows * 7 + D)
    x = torch.randn(rows, D, generator=g)
    br = (torch.randn(rows, D, generator=g) * 0.5).to(torch.bfloat16) if with_branch else None
    gamma = torch.rand(D, generator=g) + 0.5
    beta = torch.randn(D, generator=g) * 0.1
    dy = (torch.randn(rows, D, generator=g)).to(torch.bfloat16)
    dres = torch.randn(rows, D, generator=g)

    # oracle in float64
    xr = x.double().requires_grad_(True)
    brr = br.double().requires_grad_(True) if with_branch else None
    gr, br_ = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    xo = xr + brr if with_branch else xr
    y_ref = vit_ref.layer_norm(xo, gr, br_)
    (y_ref * dy.double()).sum().backward(retain_graph=True)
    (xo * dres.double()).sum().backward()

    xo_k, y, mean, rstd = ops.add_layernorm_fwd(x.cuda(), br.cuda() if with_branch else None,
                                                gamma.cuda(), beta.cuda())
    assert (y.double().cpu() - y_ref.detach()).abs().max().item() < 2e-2 * max(1.0, y_ref.abs().max().item())
    if with_branch:
        assert torch.allclose(xo_k.cpu().double(), xo.detach(), atol=1e-6)
    dx, dbr, dgamma, dbeta = ops.add_layernorm_bwd(dy.cuda(), xo_k, mean, rstd, gamma.cuda(), dres.cuda(),
                                                   want_dbranch=with_branch)
    ref_dx = xr.grad
    assert (dx.cpu().double() - ref_dx).abs().max().item() < 1e-4 * max(1.0, ref_dx.abs().max().item())
    if with_branch:
        assert (dbr.cpu().double() - brr.grad).abs().max().item() < 1e-2 * brr.grad.abs().max().item()
    assert (dgamma.cpu().double() - gr.grad).abs().max().item() < 1e-4 * max(1.0, gr.grad.abs().max().item())
    assert (dbeta.cpu().double() - br_.grad).abs().max().item() < 1e-4 * max(1.0, br_.grad.abs().max().item())


def test_add_only_and_strided_rows():
    ops = _ops()
    B, S, D = 6, 37, 192
    x = torch.randn(B, S, D, device="cuda")
    gamma = torch.rand(D, device="cuda") + 0.5
    beta = torch.randn(D, device="cuda")
    cls = x[:, 0]  # rows with pitch S*D
    _, y, mean, rstd = ops.add_layernorm_fwd(cls, None, gamma, beta)
    ref = vit_ref.layer_norm(cls.double().cpu(), gamma.double().cpu(), beta.double().cpu())
    assert (y.double().cpu() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    # add-only: x_out = x + branch, no LayerNorm
    br = torch.randn(B * S, D, device="cuda").to(torch.bfloat16)
    xo, y2, _, _ = ops.add_layernorm_fwd(x, br, None, None)
    assert y2 is None
    assert torch.allclose(xo, x.reshape(-1, D) + br.float(), atol=1e-6)


def test_dropout_mask_is_shared_between_fwd_and_bwd():
    ops = _ops()
    rows, D, p = 512, 384, 0.1
    x = torch.zeros(rows, D, device="cuda")
    br = torch.ones(rows, D, device="cuda", dtype=torch.bfloat16)
    xo, _, _, _ = ops.add_layernorm_fwd(x, br, None, None, dropout_p=p, seed=5, offset=11)
    keep = xo != 0
    frac = 1.0 - keep.float().mean().item()
    assert abs(frac - p) < 0.01
    assert torch.allclose(xo[keep], torch.full_like(xo[keep], 1 / (1 - p)), atol=1e-2)
    dres = torch.ones(rows, D, device="cuda")
    dx, dbr, _, _ = ops.add_layernorm_bwd(None, None, None, None, None, dres, want_dbranch=True,
                                          dropout_p=p, seed=5, offset=11)
    assert torch.equal(dbr != 0, keep)
    assert torch.equal(dx, dres)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols", [(50176, 384), (1000, 1536), (777, 768), (513, 1152), (64, 8), (300, 2048),
                                       (33, 65536), (129, 2304), (5, 16)])
def test_colsum_bf16_matches_torch(rows, cols):
    """bias gradients (colsum over rows of a bf16 matrix): dense fast path, wide and strided fallbacks"""
    from vit_core._backend import ops
    torch.manual_seed(rows + cols)
    x = torch.randn(rows, cols, device="cuda").bfloat16()
    ref = x.double().sum(0)
    got = ops.colsum_bf16(x)
    assert got.dtype == torch.float32 and got.shape == (cols,)
    assert (got.double() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    xs = torch.randn(rows, cols + 8, device="cuda").bfloat16()[:, :cols]   # row pitch != cols
    assert (ops.colsum_bf16(xs).double() - xs.double().sum(0)).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
