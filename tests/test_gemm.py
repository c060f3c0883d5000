"""GPU parity of the tcgen05 GEMM (C-ABI `vitssl_gemm_bf16`) against a float64 matmul of the same
bf16 inputs. Covers the three operand layouts used by forward / dgrad / wgrad, every tile width,
ragged edges, the fused epilogues and split-K."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from vit_core._backend import ops
    return ops


def _ref(a, b, a_mn, b_mn):
    A = a.double().t() if a_mn else a.double()
    B = b.double() if b_mn else b.double().t()
    return A @ B


def _mk(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * 0.5).to(torch.bfloat16)


def _err(got, ref):
    return ((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-9)).item()


LAYOUTS = [(False, False), (False, True), (True, True), (True, False)]


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize(
    "M,N,K",
    [
        (128, 64, 64), (128, 128, 64), (128, 192, 128), (128, 256, 64), (256, 256, 256),
        (392, 384, 384), (1000, 1152, 384), (296, 768, 192), (50, 64, 64), (4 * 196, 1536, 384),
        (2048, 2048, 2048), (300, 104, 72),
    ],
)
def test_gemm_layouts(M, N, K, a_mn, b_mn):
    ops = _ops()
    a = _mk((K, M) if a_mn else (M, K), 1)
    b = _mk((K, N) if b_mn else (N, K), 2)
    ref = _ref(a, b, a_mn, b_mn)
    out32 = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    torch.cuda.synchronize()
    e32 = _err(out32, ref)
    assert e32 < 2e-5, f"fp32-out rel err {e32}"
    out16 = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.bfloat16)
    e16 = _err(out16, ref)
    assert e16 < 6e-3, f"bf16-out rel err {e16}"


@pytest.mark.parametrize("split", [-1, 2, 3, 7])
def test_gemm_wgrad_split_k(split):
    ops = _ops()
    tokens, n_out, k_in = 4 * 196 * 3, 384, 192
    dy = _mk((tokens, n_out), 3)
    x = _mk((tokens, k_in), 4)
    ref = dy.double().t() @ x.double()
    out = ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=split)
    assert _err(out, ref) < 2e-5


def test_gemm_bias_and_alpha():
    ops = _ops()
    a, b = _mk((500, 384), 5), _mk((768, 384), 6)
    bias = torch.randn(768, device="cuda")
    ref = 0.25 * _ref(a, b, False, False) + bias.double()
    out = ops.gemm(a, b, epilogue=ops.EPI_BIAS, bias=bias, alpha=0.25, out_dtype=torch.float32)
    assert _err(out, ref) < 2e-5


def test_gemm_bias_gelu_and_dgelu():
    ops = _ops()
    M, N, K = 777, 1536, 384
    a, w = _mk((M, K), 7), _mk((N, K), 8)
    bias = torch.randn(N, device="cuda") * 0.1
    aux = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    h = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, aux=aux)
    u_ref = (_ref(a, w, False, False) + bias.double()).float().to(torch.bfloat16)
    # pre-activation is the bf16-rounded accumulator (allow 1 ulp flips)
    assert _err(aux, u_ref.double()) < 8e-3
    h_ref = torch.nn.functional.gelu(aux.double())
    assert (h.double() - h_ref).abs().max().item() < 2e-2 * max(1.0, h_ref.abs().max().item()) * 0.5
    assert _err(h, h_ref) < 6e-3
    # dgelu: dU = (dH @ W2) * gelu'(u)
    dh_src, w2 = _mk((M, 384), 9), _mk((384, N), 10)  # dY [M,384] @ W2[384,N] (b_mn)
    du = ops.gemm(dh_src, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux=aux)
    u = aux.double()
    gp = 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)
    du_ref = (dh_src.double() @ w2.double()) * gp
    assert _err(du, du_ref) < 6e-3


def test_gemm_dropout_epilogue_is_consistent():
    ops = _ops()
    M, N, K = 512, 1024, 256
    a, w = _mk((M, K), 11), _mk((N, K), 12)
    bias = torch.zeros(N, device="cuda")
    aux = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    p = 0.25
    h0 = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, aux=aux)
    hd = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, aux=aux, dropout_p=p, seed=123, offset=7)
    hd2 = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, aux=aux, dropout_p=p, seed=123, offset=7)
    assert torch.equal(hd, hd2)  # same (seed, offset) -> same mask
    nz = h0 != 0
    dropped = (hd == 0) & nz
    frac = dropped.sum().item() / nz.sum().item()
    assert abs(frac - p) < 0.01, frac
    kept = (~dropped) & (h0.abs() > 1e-2)
    ratio = (hd.float()[kept] / h0.float()[kept])
    assert (ratio - 1 / (1 - p)).abs().max().item() < 0.02
    # the backward epilogue regenerates the same mask
    dh_src, w2 = _mk((M, 128), 13), _mk((128, N), 14)
    du0 = ops.gemm(dh_src, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux=aux)
    dud = ops.gemm(dh_src, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux=aux, dropout_p=p, seed=123, offset=7)
    nz2 = du0 != 0
    assert torch.equal((dud == 0) & nz2 & nz, dropped & nz2)


@pytest.mark.parametrize("M,N,K", [(777, 1536, 384), (4116, 3072, 768), (50, 64, 64)])
def test_gemm_gelu_with_saved_backward_factor_and_mul_epilogue(M, N, K):
    """BIAS_GELU_D saves mask/(1-p) * gelu'(u) instead of u; MUL multiplies the dgrad accumulator by it
    (the pair the encoder stack uses): same h as BIAS_GELU, and du equals the DGELU result to bf16
    rounding of the saved factor; with dropout both see the SAME mask without regenerating it."""
    ops = _ops()
    a, w = _mk((M, K), 21), _mk((N, K), 22)
    bias = torch.randn(N, device="cuda") * 0.1
    u = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    gf = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    h_old = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, aux=u)
    h_new = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU_D, bias=bias, aux=gf)
    assert torch.equal(h_old, h_new)
    gf = gf.view(torch.float16)                                    # the saved factor is stored as fp16
    ud = u.double()
    gp = 0.5 * (1 + torch.erf(ud / math.sqrt(2))) + ud * torch.exp(-0.5 * ud * ud) / math.sqrt(2 * math.pi)
    assert (gf.double() - gp).abs().max().item() <= 6e-4          # fp16 rounding of a value in [-0.13, 1.13] + Phi polynomial
    dy, w2 = _mk((M, K), 23), _mk((K, N), 24)
    du = ops.gemm(dy, w2, b_mn=True, epilogue=ops.EPI_MUL, aux=gf.view(torch.bfloat16))
    du_ref = (dy.double() @ w2.double()) * gp
    assert _err(du, du_ref) < 8e-3
    assert _err(du, (dy.double() @ w2.double()) * gf.double()) < 5e-3   # exactly acc * saved factor, rounded once
    du_old = ops.gemm(dy, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux=u)
    assert ((du.double() - du_old.double()).norm() / du_old.double().norm()).item() < 4e-3
    if N % 8 == 0:
        p = 0.25
        hd_old = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, aux=u, dropout_p=p, seed=77, offset=5)
        hd = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU_D, bias=bias, aux=gf.view(torch.bfloat16), dropout_p=p, seed=77, offset=5)
        assert torch.equal(hd, hd_old)
        nz = (h_new != 0) & (gp.abs() > 1e-3)
        dropped = (hd == 0) & nz
        assert torch.equal((gf == 0) & nz, dropped)                    # the saved factor carries the same mask
        kept = (~dropped) & nz & (gp.abs() > 5e-2)
        assert ((gf.double()[kept] / gp[kept]) - 1 / (1 - p)).abs().max().item() < 0.02
        dud = ops.gemm(dy, w2, b_mn=True, epilogue=ops.EPI_MUL, aux=gf.view(torch.bfloat16))
        dud_old = ops.gemm(dy, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux=u, dropout_p=p, seed=77, offset=5)
        assert ((dud.double() - dud_old.double()).norm() / dud_old.double().norm()).item() < 4e-3


def test_gemm_unaligned_falls_back_to_simt():
    ops = _ops()
    # 10-class head: wgrad has a [B,10] operand whose pitch is not 16-byte aligned
    dy = _mk((64, 10), 15)
    x = _mk((64, 384), 16)
    out = ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32)
    assert _err(out, dy.double().t() @ x.double()) < 2e-5
    w = _mk((10, 384), 17)
    out2 = ops.gemm(dy, w, b_mn=True, out_dtype=torch.float32)  # dgrad: lda = 10
    assert _err(out2, dy.double() @ w.double()) < 2e-5


def test_gemm_rejects_bad_arguments():
    ops = _ops()
    from vit_core._backend.lib import VitsslError
    a, b = _mk((128, 64), 1), _mk((128, 64), 2)
    with pytest.raises(VitsslError):
        ops.gemm(a, b, epilogue=ops.EPI_BIAS)  # bias missing


@pytest.mark.parametrize("M,N,K", [(4100, 768, 768), (2048, 1536, 512), (1024, 3072, 768)])
def test_gemm_cta_pair_mode_epilogues(M, N, K):
    """N, K >= 512 and M >= 1024 run as CTA pairs (tcgen05 cta_group::2, 256-row tiles): every
    epilogue, a ragged last tile, MN-major operands and split-K wgrad must match the reference."""
    ops = _ops()
    a, w = _mk((M, K), 21), _mk((N, K), 22)
    bias = torch.randn(N, device="cuda") * 0.1
    ref = _ref(a, w, False, False)
    assert _err(ops.gemm(a, w, out_dtype=torch.float32), ref) < 2e-5
    assert _err(ops.gemm(a, w, epilogue=ops.EPI_BIAS, bias=bias, out_dtype=torch.float32), ref + bias.double()) < 2e-5
    aux = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    h = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, aux=aux)
    u_ref = (ref + bias.double()).float().to(torch.bfloat16)
    assert _err(aux, u_ref.double()) < 8e-3
    assert _err(h, torch.nn.functional.gelu(aux.double())) < 6e-3
    dy, w2 = _mk((M, K), 23), _mk((K, N), 24)            # dgrad-shaped: dY[M,K] @ W2[K,N] (b_mn)
    du = ops.gemm(dy, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux=aux, dropout_p=0.0)
    u = aux.double()
    gp = 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)
    assert _err(du, (dy.double() @ w2.double()) * gp) < 6e-3
    hd = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, aux=aux, dropout_p=0.25, seed=5, offset=3)
    dud = ops.gemm(dy, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux=aux, dropout_p=0.25, seed=5, offset=3)
    live = (h != 0) & (du != 0)
    assert torch.equal((hd == 0) & live, (dud == 0) & live)   # fwd and bwd regenerate the same mask
    # wgrad with the token axis as reduction: dW[N, K] = dY^T X, both operands MN-major, split-K
    g, x = _mk((M, N), 25), _mk((M, K), 26)
    for split in (0, -1, 3):
        dw = ops.gemm(g, x, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=split)
        assert _err(dw, g.double().t() @ x.double()) < 2e-5, split


@pytest.mark.parametrize("Mtok,Nout,Kin", [(4096, 384, 1536), (4096, 1536, 384), (1000, 200, 512), (777, 64, 64), (2048, 128, 256)])
def test_gemm_rowsum_fuses_the_bias_gradient(Mtok, Nout, Kin):
    """dW[Nout, Kin] = dY^T X and db[Nout] = colsum(dY) from one kernel (ones-operand MMA into a
    spare TMEM accumulator): split-K variants, ragged row tiles, accumulation into caller buffers."""
    ops = _ops()
    dy, x = _mk((Mtok, Nout), 31), _mk((Mtok, Kin), 32)
    dw_ref = dy.double().t() @ x.double()
    db_ref = dy.double().sum(0)
    for split in (0, -1, 3):
        dw, db = ops.gemm_rowsum(dy, x, a_mn=True, b_mn=True, split_k=split)
        assert _err(dw, dw_ref) < 2e-5, split
        assert _err(db, db_ref) < 2e-5, split
    # accumulate into caller-zeroed buffers (the encoder backward's mode), twice -> 2x
    dw = torch.zeros((Nout, Kin), device="cuda"); db = torch.zeros((Nout,), device="cuda")
    if Mtok >= 1024:  # split_k = -2 accumulates only when the problem is actually split
        for _ in range(2):
            ops.gemm_rowsum(dy, x, a_mn=True, b_mn=True, split_k=-2, out=dw, rowsum=db)
        assert _err(db, 2 * db_ref) < 2e-5
