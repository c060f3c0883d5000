"""GPU tests of the round-2 additions: fused AdamW (+GradScaler hand-over, bf16 shadows), EMA with
shadows, uint8 image kernels, bicubic positional-embedding kernel, multi-pass encoder-stack node,
nn.L1Loss dispatch, the C-side launch profile, shape validation and DINO-loss generality."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import vit_ref
from parity_utils import rel, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------
# fused AdamW vs torch.optim.AdamW (utils/train_utils.py:25-29; simmim_trainer.py:69-71)
# ------------------------------------------------------------------------------------------
def _param_sets(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(384, 384), (1536,), (7, 13), (3,), (1, 196, 384), (65, 33, 3), (2048, 384)]
    return [torch.randn(*s, generator=g) for s in shapes]


@pytest.mark.parametrize("wd,lr", [(1e-3, 1e-4), (0.0, 3e-3), (0.05, 1e-6)])
def test_fused_adamw_matches_torch_adamw(wd, lr):
    from vit_core.optim import FusedAdamW
    init = _param_sets()
    ours = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    ref = [torch.nn.Parameter(t.clone().cuda().double()) for t in init]        # float64 torch AdamW as ground truth
    ref32 = [torch.nn.Parameter(t.clone().cuda()) for t in init]               # float32 torch AdamW as yardstick
    o1 = FusedAdamW(ours, lr=lr, weight_decay=wd)
    o2 = torch.optim.AdamW(ref, lr=lr, weight_decay=wd)
    o3 = torch.optim.AdamW(ref32, lr=lr, weight_decay=wd)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        for a, b, c in zip(ours, ref, ref32):
            gr = torch.randn(a.shape, generator=g).cuda() * (10.0 ** (step - 2))
            a.grad, b.grad, c.grad = gr.clone(), gr.double(), gr.clone()
        o1.step(); o2.step(); o3.step()
        for a, b, c in zip(ours, ref, ref32):
            e_ours = (a.detach().double() - b.detach()).abs().max().item()
            e_t32 = (c.detach().double() - b.detach()).abs().max().item()
            assert e_ours <= max(4 * e_t32, 1e-7 * b.detach().abs().max().item()), (step, a.shape, e_ours, e_t32)
    for a, b in zip(ours, ref):
        sa, sb = o1.state[a], o2.state[b]
        assert float(sa["step"]) == float(sb["step"]) == 4.0
        assert rel(sa["exp_avg"], sb["exp_avg"]) <= 1e-5 and rel(sa["exp_avg_sq"], sb["exp_avg_sq"]) <= 1e-4
    sd = o1.state_dict()                                                       # base_trainer.py:99,112 checkpoints it
    o4 = FusedAdamW([torch.nn.Parameter(t.clone().cuda()) for t in init], lr=lr, weight_decay=wd)
    o4.load_state_dict(sd)
    assert float(o4.state[o4.param_groups[0]["params"][0]]["step"]) == 4.0


def test_fused_adamw_under_gradscaler_unscales_and_skips_on_inf():
    from vit_core.optim import FusedAdamW
    init = _param_sets(3)
    ours = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    ref = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    o1, o2 = FusedAdamW(ours, lr=1e-3, weight_decay=1e-2), torch.optim.AdamW(ref, lr=1e-3, weight_decay=1e-2)
    s1, s2 = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    g = torch.Generator().manual_seed(2)
    for step in range(4):
        for sc in (s1, s2):
            sc.scale(torch.zeros((), device="cuda"))                            # what scaler.scale(loss) does first: lazy init
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g).cuda()
            if step == 1:
                gr.view(-1)[0] = float("inf")                                   # this step must be skipped by both
            scale = float(s1.get_scale())
            a.grad, b.grad = gr * scale, (gr * scale).clone()
        for sc, opt in ((s1, o1), (s2, o2)):
            sc.step(opt)
            sc.update()
        assert float(s1.get_scale()) == float(s2.get_scale())
        for a, b in zip(ours, ref):
            assert torch.isfinite(a).all()
            assert (a.detach() - b.detach()).abs().max().item() <= 2e-6 * max(1.0, b.detach().abs().max().item()), (step, a.shape)
    assert float(o1.state[ours[0]]["step"]) == 3.0                              # the inf step did not count
    assert float(s1.get_scale()) == 512.0


def test_fused_adamw_and_ema_keep_the_bf16_weight_shadows_fresh(monkeypatch):
    """The optimizer / EMA kernels write the GEMM-operand shadows in their own pass: the next forward
    must not launch a cast, and must see exactly bf16(updated fp32 weights)."""
    from vit_core import EncoderBlock
    from vit_core._backend import functional as Fb, ops
    from vit_core.optim import FusedAdamW
    torch.manual_seed(0)
    blocks = torch.nn.ModuleList([EncoderBlock(128, 2, 256, 0.0) for _ in range(2)]).cuda().train()
    opt = FusedAdamW(blocks.parameters(), lr=1e-2, weight_decay=1e-2)
    x = torch.randn(3, 20, 128, device="cuda")
    casts = []
    real = ops.multi_cast_bf16
    monkeypatch.setattr(ops, "multi_cast_bf16", lambda s, d: (casts.append(len(s)), real(s, d))[1])
    out0, _ = Fb.encoder_stack(blocks, x)
    assert casts == [12]                                                        # first use: 6 weight tensors per block
    out0.square().mean().backward()
    opt.step()
    out1, _ = Fb.encoder_stack(blocks, x)
    assert casts == [12], casts                                                 # no cast after the fused step
    for blk in blocks:
        a = blk.self_attention
        sh = a.__dict__["_vitssl_cache"]["wqkv"][0]
        want = torch.cat([a.w_query.weight, a.w_key.weight, a.w_value.weight]).detach().bfloat16()
        assert torch.equal(sh, want)
        assert torch.equal(blk.feed_forward.__dict__["_vitssl_cache"]["w2"][0], blk.feed_forward.linear_out.weight.detach().bfloat16())
    assert not torch.equal(out0, out1)
    for blk in blocks:                                                          # a plain torch update still invalidates
        blk.self_attention.final_linear.weight.data.mul_(1.0)
        with torch.no_grad():
            blk.self_attention.final_linear.weight.mul_(0.5)
    Fb.encoder_stack(blocks, x)
    assert casts == [12, 2], casts


def test_dino_ema_writes_teacher_shadows():
    from oracle.cases import build_dino_case
    from vit_core._backend import ops
    from vit_core.ssl.dino import DINOViT
    cfg, m, views, B = build_dino_case(DINOViT)
    m.cuda().train()
    m([v.cuda() for v in views], 2)
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.momentum_update_teacher(0.9)
    blk = m.teacher_backbone.encoder_blocks[1]
    want = (0.9 * before["teacher_backbone.encoder_blocks.1.feed_forward.linear_in.weight"]
            + 0.1 * before["student_backbone.encoder_blocks.1.feed_forward.linear_in.weight"])
    assert (blk.feed_forward.linear_in.weight - want).abs().max().item() <= 1e-6
    assert torch.equal(blk.feed_forward.__dict__["_vitssl_cache"]["w1"][0], blk.feed_forward.linear_in.weight.detach().bfloat16())
    calls = []
    real = ops.multi_cast_bf16
    ops.multi_cast_bf16 = lambda s, d: (calls.append(len(s)), real(s, d))[1]
    try:
        with torch.no_grad():
            m.teacher_backbone(views[0].cuda())
    finally:
        ops.multi_cast_bf16 = real
    assert calls == [], calls


# ------------------------------------------------------------------------------------------
# uint8 input: value = byte / 255 as torchvision's ToTensor computes it
# ------------------------------------------------------------------------------------------
def test_uint8_images_equal_totensor_floats():
    from vit_core._backend import ops
    from vit_core.ssl.simmim import SimMIMViT
    g = torch.Generator().manual_seed(0)
    xb = torch.randint(0, 256, (5, 3, 32, 48), dtype=torch.uint8, generator=g).cuda()
    xf = xb.cpu().float().div(255).cuda()     # ToTensor arithmetic: true division, on the host (CUDA div-by-scalar multiplies by 1/255)
    assert torch.equal(ops.im2col_bf16(xb, 8), ops.im2col_bf16(xf, 8))
    assert torch.equal(ops.im2col_bf16(xb[:, :, :30, :45].contiguous(), 3), ops.im2col_bf16(xf[:, :, :30, :45].contiguous(), 3))
    rows = torch.tensor([0, 7, 23, 5 * 24 - 1], dtype=torch.int32, device="cuda")
    assert torch.equal(ops.gather_patches_f32(xb, rows, 8), ops.gather_patches_f32(xf, rows, 8))
    torch.manual_seed(1)
    m = SimMIMViT(num_blocks=2, input_shape=(3, 32, 32), embed_dim=128, patch_size=8, num_heads=2, mlp_dim=256,
                  dropout=0.0, mask_ratio=0.6).cuda()
    x8 = torch.randint(0, 256, (4, 3, 32, 32), dtype=torch.uint8, generator=g).cuda()
    st = torch.cuda.get_rng_state()
    p1, t1 = m(x8)
    torch.cuda.set_rng_state(st)
    p2, t2 = m(x8.cpu().float().div(255).cuda())
    assert torch.equal(p1, p2) and torch.equal(t1, t2)


# ------------------------------------------------------------------------------------------
# bicubic positional-embedding resize (patch_embedding.py:26-48) vs F.interpolate, fwd + bwd
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("gin,gout,D", [((14, 14), (6, 6), 384), ((4, 4), (2, 2), 64), ((6, 6), (14, 14), 192), ((14, 14), (7, 5), 128)])
def test_bicubic_pos_embedding_kernel_matches_f_interpolate(gin, gout, D):
    from vit_core._backend import functional as Fb
    torch.manual_seed(0)
    pos = torch.rand(1, 1 + gin[0] * gin[1], D, device="cuda", requires_grad=True)
    pos2 = pos.detach().clone().requires_grad_(True)
    out = Fb.interpolate_pos_embedding(pos, gin, gout)
    ref = vit_ref.interpolate_pos_encoding(pos2, gin, gout[0] * gout[1], gout[0], gout[1])
    assert out.shape == ref.shape == (1, 1 + gout[0] * gout[1], D)
    assert (out - ref).abs().max().item() <= 2e-6
    dy = torch.randn_like(out)
    (out * dy).sum().backward()
    (ref * dy).sum().backward()
    assert (pos.grad - pos2.grad).abs().max().item() <= 1e-5 * max(1.0, pos2.grad.abs().max().item())


def test_dynamic_patch_embedding_local_crops_use_the_kernel_and_match():
    from vit_core import DynamicPatchEmbedding
    torch.manual_seed(0)
    m = DynamicPatchEmbedding((3, 224, 224), 384, 16).cuda()
    x = torch.rand(3, 3, 96, 96, device="cuda")
    y = m(x)
    w = {"pe." + k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    ref = vit_ref.dynamic_patch_embedding(w, "pe.", x.double().cpu(), 16, (14, 14))
    assert y.shape == (3, 37, 384) and rel(y, ref) <= 1e-2
    y.sum().backward()
    assert m.positional_embedding.grad is not None and m.positional_embedding.grad.shape == (1, 197, 384)
    # d(sum)/d(pos) = B * (column sums of the interpolation matrix), CLS row = B
    assert abs(m.positional_embedding.grad[0, 0, 0].item() - 3.0) <= 1e-4


# ------------------------------------------------------------------------------------------
# several token batches through the same blocks as ONE autograd node (DINO global + local passes)
# ------------------------------------------------------------------------------------------
def test_multi_pass_encoder_stack_matches_separate_nodes(monkeypatch):
    from vit_core import EncoderBlock
    from vit_core._backend import functional as Fb
    torch.manual_seed(5)
    blocks = torch.nn.ModuleList([EncoderBlock(128, 2, 256, 0.1) for _ in range(3)]).cuda().train()
    xa, xb = torch.randn(4, 50, 128, device="cuda"), torch.randn(7, 17, 128, device="cuda")
    seeds = iter([11, 22, 11, 22])
    monkeypatch.setattr(Fb, "_new_seed", lambda: next(seeds))

    def run(multi):
        blocks.zero_grad(set_to_none=True)
        a, b = xa.clone().requires_grad_(True), xb.clone().requires_grad_(True)
        if multi:
            oa, ob = Fb.encoder_stack_multi(blocks, [a, b])
        else:
            oa, ob = Fb.encoder_stack(blocks, a)[0], Fb.encoder_stack(blocks, b)[0]
        (oa.square().mean() + ob.sum() * 1e-3).backward()
        return oa.detach(), ob.detach(), [p.grad.clone() for p in blocks.parameters()], a.grad, b.grad

    o1 = run(True)
    o2 = run(False)
    assert torch.equal(o1[0], o2[0]) and torch.equal(o1[1], o2[1])
    for g1, g2 in zip(o1[2], o2[2]):
        assert (g1 - g2).abs().max().item() <= 1e-5 * max(g2.abs().max().item(), 1e-6)
    assert torch.allclose(o1[3], o2[3], atol=1e-6) and torch.allclose(o1[4], o2[4], atol=1e-6)


# ------------------------------------------------------------------------------------------
# nn.L1Loss on the model's prediction runs the fused kernel (simmim_trainer.py:66-67)
# ------------------------------------------------------------------------------------------
def test_l1loss_on_model_output_dispatches_to_the_fused_kernel_and_matches_torch():
    from vit_core.ssl.simmim import SimMIMViT
    from vit_core.ssl.simmim.model import MaskedPrediction
    torch.manual_seed(2)
    m = SimMIMViT(num_blocks=2, input_shape=(3, 32, 32), embed_dim=128, patch_size=8, num_heads=2, mlp_dim=256,
                  dropout=0.0, mask_ratio=0.6).cuda().train()
    x = torch.rand(6, 3, 32, 32, device="cuda")
    st = torch.cuda.get_rng_state()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pred, tg = m(x)
        assert isinstance(pred, MaskedPrediction) and pred.dtype == torch.bfloat16
        loss = torch.nn.L1Loss()(pred, tg)
    # (the loss is a PrefetchedScalar: an alias node sits between it and the kernel's autograd node)
    chain = [type(loss.grad_fn).__name__] + [type(f).__name__ for f, _ in loss.grad_fn.next_functions if f is not None]
    assert any("_L1LossFn" in n for n in chain) and loss.dtype == torch.float32
    scaler = torch.amp.GradScaler("cuda", init_scale=4096.0)
    scaler.scale(loss).backward()
    g1 = m.simmim_head.weight.grad.clone() / 4096.0
    plain = pred.detach().as_subclass(torch.Tensor)
    ref_loss = (plain.float() - tg).abs().mean()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * ref_loss.item()
    # torch's own L1 on a plain tensor of the same forward: same gradients
    m.zero_grad(set_to_none=True)
    torch.cuda.set_rng_state(st)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pred2, tg2 = m(x)
        loss2 = F.l1_loss(pred2.as_subclass(torch.Tensor), tg2)
    assert "_L1LossFn" not in type(loss2.grad_fn).__name__
    loss2.backward()
    assert rel_l2(g1, m.simmim_head.weight.grad) <= 1e-2
    # everything else behaves like a plain tensor and drops the subclass
    r = torch.clamp(pred.reshape(-1, 3, 8, 8), 0, 1)
    assert type(r) is torch.Tensor and r.shape[1:] == (3, 8, 8)
    assert type(torch.nn.L1Loss(reduction="sum")(pred, tg)) is torch.Tensor
    assert abs(torch.nn.L1Loss(reduction="sum")(pred, tg).item() - (plain.float() - tg).abs().sum().item()) <= 1e-2 * ref_loss.item() * tg.numel()
    assert type(torch.nn.MSELoss()(pred.float(), tg)) is torch.Tensor


# ------------------------------------------------------------------------------------------
# per-launch profile of the C-sequenced stack (what bench.py's roofline numbers are made of)
# ------------------------------------------------------------------------------------------
def test_c_side_profile_records_every_kernel_family_of_the_stack():
    from vit_core import EncoderBlock
    from vit_core._backend import functional as Fb, lib
    torch.manual_seed(0)
    L = 3
    blocks = torch.nn.ModuleList([EncoderBlock(128, 2, 256, 0.1) for _ in range(L)]).cuda().train()
    x = torch.randn(4, 64, 128, device="cuda", requires_grad=True)
    Fb.encoder_stack(blocks, x)[0].sum().backward()                            # warm-up, not profiled
    lib.profile_begin()
    out, _ = Fb.encoder_stack(blocks, x)
    out.sum().backward()
    recs = lib.profile_collect()
    kinds = [k for k, _, _ in recs]
    assert kinds.count("attn_fwd") == L and kinds.count("attn_bwd") == L
    assert kinds.count("add_layernorm") == (2 * L + 1) * 2
    gemms = [k for k in kinds if k.startswith("gemm|")]
    assert len(gemms) == 4 * L + 8 * L                                          # 4 forward + 4 dgrad + 4 wgrad per block
    assert all(ms > 0 and work > 0 for _, ms, work in recs)
    M = 4 * 64
    assert f"gemm|{M}x256x128|a_mn=0 b_mn=0 epi=4" in kinds and f"gemm|{M}x256x128|a_mn=0 b_mn=1 epi=5" in kinds
    assert [k for k, _, _ in lib.profile_collect()] == kinds                   # closed profile: reading again is harmless
    Fb.encoder_stack(blocks, x)                                                 # and nothing is recorded when closed
    lib.profile_begin()
    assert lib.profile_collect() == []


# ------------------------------------------------------------------------------------------
# shape validation (advisor, round 1): a wrong-size positional embedding must raise, not read OOB
# ------------------------------------------------------------------------------------------
def test_patch_embedding_rejects_images_that_do_not_match_the_positional_embedding():
    from vit_core import ConvolutionalPatchEmbedding, ManualPatchEmbedding
    from vit_core.ssl.simmim import SimMIMViT
    for cls in (ConvolutionalPatchEmbedding, ManualPatchEmbedding):
        m = cls((3, 32, 32), 64, 8).cuda()
        with pytest.raises(ValueError):
            m(torch.rand(2, 3, 64, 64, device="cuda"))                          # larger grid than the embedding
        with pytest.raises(ValueError):
            m(torch.rand(2, 3, 16, 16, device="cuda"))                          # smaller grid
        with pytest.raises(ValueError):
            m(torch.rand(2, 1, 32, 32, device="cuda"))                          # wrong channel count
    s = SimMIMViT(num_blocks=1, input_shape=(3, 32, 32), embed_dim=128, patch_size=8, num_heads=2, mlp_dim=256, dropout=0.0).cuda()
    with pytest.raises(ValueError):
        s(torch.rand(2, 3, 64, 64, device="cuda"))
    with pytest.raises(ValueError):
        s.inference_forward(torch.rand(2, 3, 16, 16, device="cuda"))


# ------------------------------------------------------------------------------------------
# DINO loss beyond the trainer's shapes (advisor, round 1): > 12 views, 3-4 teacher views, 2-D inputs
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tshape,sshape", [((2, 3, 1024), (14, 3, 1024)), ((3, 5, 512), (7, 5, 512)), ((4, 2, 256), (25, 2, 256)),
                                           ((2, 2048), (6, 2048)), ((2, 2, 3, 256), (5, 2, 3, 256))])
def test_dino_loss_general_shapes(tshape, sshape):
    from vit_core.ssl.dino.loss import DINOLoss
    g = torch.Generator().manual_seed(sum(tshape) + sum(sshape))
    t = (torch.randn(*tshape, generator=g) * 2).bfloat16()
    s0 = (torch.randn(*sshape, generator=g) * 1.5).bfloat16()
    c = torch.randn(1, tshape[-1], generator=g) * 0.1
    s = s0.cuda().requires_grad_(True)
    loss = DINOLoss(0.05, 0.1)(t.cuda(), s, c.cuda())
    sb = s0.double().requires_grad_(True)
    ref = vit_ref.dino_loss(t.double(), sb, c.double(), 0.05, 0.1)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-4 * abs(ref.item()), (loss.item(), ref.item())
    loss.backward()
    assert rel_l2(s.grad, sb.grad) <= 1e-2
    with pytest.raises(ValueError):
        DINOLoss(0.05, 0.1)(t.cuda()[..., :8], s, c.cuda())


# ------------------------------------------------------------------------------------------
# view packing: a list of equally-shaped crops == their concatenation, bit for bit
# ------------------------------------------------------------------------------------------
def test_view_lists_are_embedded_like_their_concatenation():
    from oracle.cases import build_dino_case
    from vit_core import DynamicPatchEmbedding
    from vit_core.ssl.dino import DINOViT
    torch.manual_seed(0)
    m = DynamicPatchEmbedding((3, 32, 32), 64, 8).cuda()
    views = [torch.rand(3, 3, 32, 32, device="cuda") for _ in range(3)]
    a = m(views)
    b = m(torch.cat(views))
    assert a.shape == (9, 17, 64) and torch.equal(a, b)
    a.sum().backward()
    g1 = m.proj.weight.grad.clone()
    m.zero_grad()
    m(torch.cat(views)).sum().backward()
    assert torch.equal(g1, m.proj.weight.grad)
    u8 = [torch.randint(0, 256, (2, 3, 16, 16), dtype=torch.uint8, device="cuda") for _ in range(2)]
    assert torch.equal(m(u8), m(torch.cat(u8)))
    with pytest.raises(ValueError):
        m([views[0], views[1][:, :, :16, :16]])
    # the whole DINO forward: list-of-views path (no torch.cat of images) == reference-shaped call
    cfg, dm, dviews, B = build_dino_case(DINOViT)
    dm.cuda().eval()
    with torch.no_grad():
        t1, s1 = dm([v.cuda() for v in dviews], 2)
        c1 = dm.center.clone()
        g = torch.cat([v.cuda() for v in dviews[:2]])
        t2 = dm.teacher_head(dm.teacher_backbone(g))
    assert torch.equal(t1, t2) and t1.shape == (2 * B, 512)


# ------------------------------------------------------------------------------------------
# evaluation path: token mean-pool kernel and GPU cosine k-NN vs scikit-learn
# ------------------------------------------------------------------------------------------
def test_mean_tokens_kernel_and_simmim_inference():
    from vit_core._backend import ops
    x = torch.randn(7, 37, 192, device="cuda")
    assert (ops.mean_tokens(x) - x.mean(dim=1)).abs().max().item() <= 1e-6


@pytest.mark.parametrize("Nt,Nv,D,classes", [(500, 300, 64, 10), (1203, 777, 384, 10), (64, 33, 20, 3)])
def test_gpu_knn_matches_sklearn_cosine_knn(Nt, Nv, D, classes):
    sk = pytest.importorskip("sklearn.neighbors")
    from vit_core.evaluation import knn_predict, run_knn_evaluation
    g = torch.Generator().manual_seed(Nt + Nv)
    centers = torch.randn(classes, D, generator=g)
    yt = torch.randint(0, classes, (Nt,), generator=g)
    yv = torch.randint(0, classes, (Nv,), generator=g)
    xt = centers[yt] + 1.5 * torch.randn(Nt, D, generator=g)
    xv = centers[yv] + 1.5 * torch.randn(Nv, D, generator=g)
    knn = sk.KNeighborsClassifier(n_neighbors=classes, metric="cosine")     # evaluators/unsupervised_evaluator.py:52
    knn.fit(xt.numpy(), yt.numpy())
    want = torch.from_numpy(knn.predict(xv.numpy()))
    got = knn_predict(xt.cuda(), yt.cuda(), xv.cuda(), classes)
    agree = (got.cpu() == want).float().mean().item()
    assert agree >= 0.995, agree          # identical up to float-rounding of near-tied distances
    res = run_knn_evaluation(xt, yt, xv, yv, classes)
    assert abs(res["accuracy"] - (want == yv).float().mean().item()) <= 0.01 and res["num_neighbors"] == classes


@pytest.mark.gpu
def test_prefetched_loss_scalar_value_autograd_and_ring():
    """The loss the SimMIM / DINO paths return: `.item()` reads the copy that was enqueued behind the
    loss kernel (not behind the whole step), with the plain tensor's value, autograd and arithmetic."""
    from vit_core._backend.scalar import PrefetchedScalar, prefetch_scalar
    w = torch.randn(64, device="cuda", requires_grad=True)
    raw = (w * w).sum()
    loss = prefetch_scalar(raw)
    assert isinstance(loss, PrefetchedScalar) and loss.data_ptr() == raw.data_ptr()
    scaled = loss * 3.0
    assert type(scaled) is torch.Tensor                      # e.g. scaler.scale(loss)
    scaled.backward()
    assert torch.allclose(w.grad, 6.0 * w.detach())
    assert loss.item() == torch.Tensor.item(raw) == float(loss)
    assert loss.item() == loss.item()                        # second read: remembered float
    # more outstanding losses than ring slots: the oldest are read before their slot is reused
    vals = [prefetch_scalar(torch.full((), float(i), device="cuda")) for i in range(80)]
    assert [v.item() for v in vals] == [float(i) for i in range(80)]
    # not a scalar / not CUDA: handed back untouched
    assert type(prefetch_scalar(torch.ones(2, device="cuda"))) is torch.Tensor
    assert type(prefetch_scalar(torch.ones(()))) is torch.Tensor


@pytest.mark.gpu
def test_simmim_and_dino_losses_are_prefetched():
    from vit_core._backend.scalar import PrefetchedScalar
    from vit_core.ssl.dino.loss import DINOLoss
    from vit_core.ssl.simmim import SimMIMViT
    torch.manual_seed(0)
    m = SimMIMViT(num_blocks=1, input_shape=(3, 32, 32), embed_dim=128, patch_size=8, num_heads=2, mlp_dim=256,
                  dropout=0.0, mask_ratio=0.5).cuda().train()
    x = torch.rand(4, 3, 32, 32, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pred, tgt = m(x)
        loss = torch.nn.L1Loss()(pred, tgt)
    assert isinstance(loss, PrefetchedScalar)
    ref = (pred.detach().float().as_subclass(torch.Tensor) - tgt.float()).abs().mean().item()
    loss.backward()
    assert abs(loss.item() - ref) <= 2e-3 * abs(ref)
    t = torch.randn(2, 4, 512, device="cuda")
    s = torch.randn(4, 4, 512, device="cuda", requires_grad=True)
    dl = DINOLoss(0.04, 0.1)(t, s, torch.zeros(512, device="cuda"))
    assert isinstance(dl, PrefetchedScalar) and dl.item() == torch.Tensor.item(dl)


# ------------------------------------------------------------------------------------------
# CUDA-graph replay of the C-sequenced stack (csrc/encoder.cu): same bits as the direct path
# ------------------------------------------------------------------------------------------
_GRAPH_SCRIPT = r"""
import json, os, sys
sys.path.insert(0, os.path.join(os.environ["VITSSL_ROOT"], "vit-ssl_b200"))
import torch
from torch import nn
from vit_core import EncoderBlock
from vit_core._backend import functional as Fb, lib
torch.manual_seed(0)
blocks = nn.ModuleList([EncoderBlock(128, 2, 256, 0.1) for _ in range(3)]).cuda().train()
params = [p for p in blocks.parameters()]
x = torch.randn(64, 64, 128, device="cuda", requires_grad=True)      # 4096 token rows
w = torch.randn(64, 64, 128, device="cuda")
out = []
for step in range(8):
    torch.manual_seed(100 + step)               # the dropout seed of the step is drawn from this stream
    for p in params + [x]:
        p.grad = None
    y, _ = Fb.encoder_stack(blocks, x)
    (y * w).sum().backward()
    g = [(float(p.grad.double().sum()), float(p.grad.double().abs().sum())) for p in params]
    out.append((float(y.double().sum()).hex(), float(y.double().abs().sum()).hex(), float(x.grad.double().abs().sum()).hex(), g))
    with torch.no_grad():                      # weights move (by the same amounts in both runs): shadows are re-cast
        gen = torch.Generator(device="cuda").manual_seed(7 + step)
        for p in params:
            p.add_(torch.randn(p.shape, device="cuda", generator=gen), alpha=1e-3)
print(json.dumps({"steps": out, "graphs": lib.graph_stats()}))
"""


def _run_graph_script(graph):
    import json
    import subprocess
    import sys
    env = dict(os.environ, VITSSL_GRAPH=str(graph), VITSSL_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-c", _GRAPH_SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_stack_graph_replay_is_bit_identical_to_direct_launches():
    """Eight forward + backward passes of a 3-block stack with dropout 0.1 (fresh seed every step, weights
    moving): with the graph cache the later steps are replays of a captured launch sequence whose only
    per-step input, the dropout seed, travels through a device word. Outputs and every gradient must
    equal the direct path's bit for bit."""
    direct = _run_graph_script(0)
    replay = _run_graph_script(1)
    assert direct["graphs"] == [0, 0]
    captured, replayed = replay["graphs"]
    assert captured >= 2 and replayed >= 6, replay["graphs"]       # forward and backward graphs, reused
    for a, b in zip(replay["steps"], direct["steps"]):
        # outputs and the input gradient: bit for bit. Parameter gradients are accumulated with split-K
        # reduce-adds / atomics whose order differs from run to run (with or without graphs)
        assert a[:3] == b[:3]
        for (ga, na), (gb, nb) in zip(a[3], b[3]):
            assert abs(ga - gb) <= 1e-5 * nb and abs(na - nb) <= 1e-5 * nb, (ga, gb, na, nb)
    assert len({s[0] for s in direct["steps"]}) == 8              # different seeds / weights every step
