"""GPU parity of the drop-in `vit_core` modules against golden vectors produced by the real
reference (float64, same weights, same inputs, dropout 0). Tolerances are the north-star ones:
max relative error <= 1e-2 on activations and gradients (bf16 compute, fp32 accumulation),
<= 1e-3 on losses; masks bit-exact."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import vit_ref
from oracle.cases import build_dino_case, digest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ACT_TOL, GRAD_TOL, LOSS_TOL = 1e-2, 1e-2, 1e-3


def load(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def rel(a, b):
    b = b.double().cpu()
    return ((a.detach().double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    b = b.double().cpu()
    return ((a.detach().double().cpu() - b).norm() / b.norm().clamp_min(1e-30)).item()


def check_grads(module, golden_grads, yard=None, tol=GRAD_TOL):
    """Every parameter gradient: relative L2 error <= tol, and max-abs relative error <= tol — or,
    for the few tensors where bf16 rounding alone exceeds that (e.g. last-block query weights fed
    by a handful of CLS rows), <= 2x the error the reference algorithm itself shows under
    autocast(bf16) on the same GPU (`yard`, computed with the oracle)."""
    worst = ("", 0.0, 0.0)
    for k, p in module.named_parameters():
        if k not in golden_grads:
            continue
        assert p.grad is not None, k
        e2, e = rel_l2(p.grad, golden_grads[k]), rel(p.grad, golden_grads[k])
        allowed2 = allowed = tol
        if yard is not None and k in yard:
            allowed2 = max(tol, 2.0 * rel_l2(yard[k], golden_grads[k]))
            allowed = max(tol, 2.0 * rel(yard[k], golden_grads[k]))
        assert e2 <= allowed2, (k, "rel-L2", e2, "allowed", allowed2)
        assert e <= allowed, (k, "max-rel", e, "allowed", allowed)
        if e > worst[1]:
            worst = (k, e, allowed)
    return worst


def autocast_yardstick(weights, run):
    """Gradients of the reference algorithm (oracle) under autocast(bf16) on this GPU."""
    w = {k: v.cuda().clone().requires_grad_(v.is_floating_point()) for k, v in weights.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = run(w)
    loss.backward()
    return {k: v.grad for k, v in w.items() if v.grad is not None}


def test_encoder_block():
    from vit_core import EncoderBlock
    g = load("encoder_block")
    m = EncoderBlock(d_model=128, num_heads=2, mlp_dim=256, dropout=0.0)
    m.load_state_dict(g["weights"])
    m.cuda()
    x = g["x"].cuda().requires_grad_(True)
    y, probs = m(x, return_attn=True)
    assert y.dtype == torch.float32 and y.shape == g["y"].shape
    assert rel(y, g["y"]) <= ACT_TOL
    assert rel(probs, g["probs"]) <= ACT_TOL
    (y * g["dy"].cuda()).sum().backward()
    assert rel(x.grad, g["dx"]) <= GRAD_TOL
    check_grads(m, g["grads"])
    y2, none = m(g["x"].cuda())
    assert none is None and torch.equal(y2, y)


def test_mha_cross_attention():
    from vit_core import MultiHeadedAttention
    g = load("mha_cross")
    m = MultiHeadedAttention(128, 2)
    m.load_state_dict(g["weights"])
    m.cuda()
    q, k, v = (g[n].cuda().requires_grad_(True) for n in ("q", "k", "v"))
    out, probs = m(q, k, v, return_attn=True)
    assert out.shape == (2, 10, 128) and probs.shape == (2, 2, 10, 12)
    assert rel(out, g["out"]) <= ACT_TOL and rel(probs, g["probs"]) <= ACT_TOL
    (out * g["dy"].cuda()).sum().backward()
    assert rel(q.grad, g["dq"]) <= GRAD_TOL and rel(k.grad, g["dk"]) <= GRAD_TOL and rel(v.grad, g["dv"]) <= GRAD_TOL
    check_grads(m, g["grads"])


def test_feed_forward():
    from vit_core import FeedForwardBlock
    g = load("ffn")
    m = FeedForwardBlock(64, 128, dropout=0.0)
    m.load_state_dict(g["weights"])
    m.cuda()
    x = g["x"].cuda().requires_grad_(True)
    y = m(x)
    assert rel(y, g["y"]) <= ACT_TOL
    (y * g["dy"].cuda()).sum().backward()
    assert rel(x.grad, g["dx"]) <= GRAD_TOL
    check_grads(m, g["grads"])


def test_patch_embeddings():
    from vit_core import ConvolutionalPatchEmbedding, DynamicPatchEmbedding, ManualPatchEmbedding
    g = load("patch_embeddings")
    x = g["x"].cuda()
    for name, cls in (("conv", ConvolutionalPatchEmbedding), ("manual", ManualPatchEmbedding)):
        m = cls((3, 32, 32), 64, 8)
        m.load_state_dict(g[name]["weights"])
        y = m.cuda()(x)
        assert y.shape == (3, 17, 64) and y.dtype == torch.float32
        assert rel(y, g[name]["y"]) <= ACT_TOL, name
    m = DynamicPatchEmbedding((3, 32, 32), 64, 8)
    m.load_state_dict(g["dynamic"]["weights"])
    m.cuda()
    assert rel(m(x), g["dynamic"]["y"]) <= ACT_TOL
    assert rel(m(g["dynamic"]["x_local"].cuda()), g["dynamic"]["y_local"]) <= ACT_TOL  # bicubic pos interpolation
    with pytest.raises(ValueError):
        m(torch.rand(1, 3, 30, 32, device="cuda"))


def test_vit_supervised_step():
    from vit_core import ViT
    g = load("vit")
    m = ViT(**g["cfg"])
    m.load_state_dict(g["weights"])
    m.cuda()
    logits, probs = m(g["x"].cuda(), return_attn=True)
    assert logits.shape == (4, 10) and logits.dtype == torch.float32
    assert rel(logits, g["logits"]) <= ACT_TOL and rel(probs, g["probs"]) <= ACT_TOL
    loss = F.cross_entropy(logits, g["labels"].cuda())
    assert abs(loss.item() - g["loss"].item()) <= LOSS_TOL * abs(g["loss"].item())
    loss.backward()
    yard = autocast_yardstick(g["weights"], lambda w: F.cross_entropy(
        vit_ref.vit_forward(w, g["x"].cuda(), patch_size=8, num_blocks=2, num_heads=2)[0].float(), g["labels"].cuda()))
    check_grads(m, g["grads"], yard)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert m(g["x"].cuda()).dtype == torch.bfloat16  # nn.Linear output dtype under autocast


def _simmim(g):
    from vit_core.ssl.simmim import SimMIMViT
    import vit_core.ssl.simmim.model as mod
    m = SimMIMViT(**g["cfg"])
    m.load_state_dict(g["weights"])
    m.cuda()
    idx = g["perms"][:, : int(16 * 0.6)].cuda()
    return m, mod, idx


def test_simmim_forward_backward(monkeypatch):
    g = load("simmim")
    m, mod, idx = _simmim(g)
    from vit_core.ssl.simmim.masking import mask_tables
    monkeypatch.setattr(mod, "draw_mask", lambda B, N, r, dev, want_indices=True: (idx, *mask_tables(idx, N)))
    pred, targets, bool_mask = m(g["x"].cuda(), return_bool_mask=True)
    assert torch.equal(bool_mask.squeeze(-1).cpu(), g["bool_mask"])          # mask bit-exact
    assert torch.equal(targets.cpu().double(), g["targets"])                 # raw pixels bit-exact
    assert rel(pred, g["pred"]) <= ACT_TOL
    loss = torch.nn.L1Loss()(pred, targets)                                  # the reference trainer's criterion
    assert abs(loss.item() - g["loss"].item()) <= LOSS_TOL * g["loss"].item()
    loss.backward()
    mask_cuda = g["bool_mask"].cuda()

    def run(w):
        pred_, tg_ = vit_ref.simmim_forward(w, g["x"].cuda(), mask_cuda, patch_size=8, num_blocks=2, num_heads=2)
        return vit_ref.l1_loss(pred_.float(), tg_)
    yard = autocast_yardstick(g["weights"], run)
    check_grads(m, g["grads"], yard)
    # fused objective path: same loss, same gradients
    m.zero_grad()
    loss2 = m.reconstruction_loss(g["x"].cuda())
    assert abs(loss2.item() - g["loss"].item()) <= LOSS_TOL * g["loss"].item()
    loss2.backward()
    check_grads(m, g["grads"], yard)
    feats = m.inference_forward(g["x"].cuda())
    assert not m.training and rel(feats, g["inference"]) <= ACT_TOL


def test_simmim_mask_is_the_reference_rng_sequence():
    """masking.py:22-25 draws B sequential torch.randperm(N, device)[:n_m]; our one-launch replay of
    that integer algorithm must reproduce torch's own draws bit for bit from the same generator
    state, leave the generator where the B calls would, and emit the oracle's mask tables."""
    from vit_core.ssl.simmim.masking import draw_mask, draw_mask_indices, mask_tables, simple_masking
    dev = torch.device("cuda")
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    # (B, N, ratio): BASELINE shapes (196, 36), the shipped 144, edge sizes, > 256 (second key block)
    cases = [(7, 196, 0.6), (256, 196, 0.6), (64, 36, 0.6), (33, 144, 0.75), (5, 1, 0.6), (9, 2, 0.5),
             (4, 257, 0.6), (3, 576, 0.6), (2, 1024, 0.9), (300, 64, 1.0), (6, 49, 0.0)]
    for seed, (B, N, r) in enumerate(cases):
        torch.manual_seed(1234 + seed)
        torch.rand(3, device="cuda")                      # move the offset off zero
        state = torch.cuda.get_rng_state()
        n_m = int(N * r)
        ref = torch.stack([torch.randperm(N, device="cuda")[:n_m] for _ in range(B)])
        off_ref = gen.get_offset()
        after_ref = torch.rand(4, device="cuda")
        torch.cuda.set_rng_state(state)
        ours, bool_mask, rows, inv = draw_mask(B, N, r, dev)
        assert gen.get_offset() == off_ref, (B, N, gen.get_offset(), off_ref)
        assert torch.equal(torch.rand(4, device="cuda"), after_ref)   # later draws continue the same stream
        assert ours.dtype == torch.int64 and torch.equal(ours, ref), (B, N, r)
        bm2, rows2, inv2 = mask_tables(ref, N)
        assert torch.equal(bool_mask, bm2) and torch.equal(rows, rows2) and torch.equal(inv, inv2), (B, N, r)
        perms = torch.cat([ours.cpu(), torch.zeros(B, N - n_m, dtype=torch.long)], dim=1)
        assert torch.equal(bool_mask.cpu(), vit_ref.mask_from_perms(perms, N, r))
        assert torch.equal(rows.cpu().long(), bool_mask.reshape(-1).nonzero().squeeze(1).cpu())
    # duplicate 18-bit keys (islands) do occur at these sizes: 4096 samples of N=196 hold ~300 of them
    torch.manual_seed(99)
    state = torch.cuda.get_rng_state()
    ref = torch.stack([torch.randperm(196, device="cuda") for _ in range(4096)])
    torch.cuda.set_rng_state(state)
    ours = draw_mask_indices(4096, 196, 1.0, dev)
    assert torch.equal(ours, ref)
    # the reference-shaped helper
    B, N, r = 7, 196, 0.6
    patches = torch.rand(B, N, 48, device="cuda")
    state = torch.cuda.get_rng_state()
    ref = torch.stack([torch.randperm(N, device="cuda")[: int(N * r)] for _ in range(B)])
    torch.cuda.set_rng_state(state)
    _, bm, tg = simple_masking(patches, r)
    assert torch.equal(bm, mask_tables(ref, N)[0]) and torch.equal(tg, patches[bm])


def test_dino_forward_loss_backward_ema():
    from vit_core.ssl.dino import DINOViT
    from vit_core.ssl.dino.loss import DINOLoss
    g = load("dino")
    cfg, m, views, B = build_dino_case(DINOViT)
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.cuda()
    m.train()
    teacher, student = m([v.cuda() for v in views], 2)
    assert teacher.shape == (2 * B, 512) and student.shape == (4 * B, 512)
    assert not teacher.requires_grad and student.requires_grad
    assert rel(teacher, g["teacher"]) <= ACT_TOL and rel(student, g["student"]) <= ACT_TOL
    assert m.center.shape == (1, 512) and rel(m.center, g["center_after"]) <= ACT_TOL
    crit = DINOLoss(*g["temps"])
    loss = crit(teacher.view(2, B, 512), student.view(4, B, 512), m.center)
    assert abs(loss.item() - g["loss"].item()) <= LOSS_TOL * abs(g["loss"].item())
    loss.backward()

    # yardstick: the reference algorithm (oracle) under autocast(bf16) on this GPU, digested the same way
    kw = dict(patch_size=8, num_blocks=2, num_heads=2, grid=(4, 4), center_momentum=0.9)
    wy = {k: v.cuda().clone().requires_grad_(k.startswith("student_")) for k, v in state.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        t_, s_, c_ = vit_ref.dino_forward(wy, [v.cuda() for v in views], 2, wy["center"], **kw)
        ly = vit_ref.dino_loss(t_.view(2, B, 512).detach(), s_.view(4, B, 512), c_.detach().float(), *g["temps"])
    ly.backward()

    def digest_err(t, dg):
        mine = digest(t.cpu())
        e = ((mine["sample"] - dg["sample"]).abs().max() / dg["sample"].abs().max().clamp_min(1e-30)).item()
        return max(e, abs(mine["norm"] - dg["norm"]) / max(dg["norm"], 1e-30))

    worst, failures = ("", 0.0, 0.0), []
    for k, p in m.named_parameters():
        if k in g["grad_digests"]:
            e = digest_err(p.grad, g["grad_digests"][k])
            ey = digest_err(wy[k].grad, g["grad_digests"][k])
            # digests are 256-element samples of a 3-image batch: 3x the reference-under-autocast error
            # here; the full-tensor check at the BASELINE shape (test_baseline_shapes.py) uses 2x
            allowed = max(GRAD_TOL, 3.0 * ey)
            if e > allowed:
                failures.append((k, round(e, 5), "autocast reference", round(ey, 5)))
            if e > worst[1]:
                worst = (k, e, allowed)
        else:
            assert p.grad is None, k  # teacher is frozen
    assert not failures, failures
    m.momentum_update_teacher(g["momentum"])
    sd = m.state_dict()
    for k, dg in g["teacher_after_digests"].items():
        e = (digest(sd[k].cpu())["sample"] - dg["sample"]).abs().max().item()
        assert e <= 1e-6 * max(1.0, dg["sample"].abs().max().item()), k
    feats = m.inference_forward(views[0].cuda())
    assert not m.training and feats.shape == (B, 512)
    assert rel(feats, g["inference"]) <= ACT_TOL                    # a22: VALUES of the evaluation entry
    backbone_feats = m.inference_forward(views[0].cuda(), return_features=True)
    assert backbone_feats.shape == (B, 128)


def test_dino_loss_kernel_matches_reference_and_closed_form():
    from vit_core.ssl.dino.loss import DINOLoss
    g = load("dino_loss")
    s = g["student"].cuda().requires_grad_(True)
    loss = DINOLoss(*g["temps"])(g["teacher"].cuda(), s, g["center"].cuda())
    # the kernels read bf16 logits: the oracle on the same bf16-rounded inputs is the tight check,
    # the golden value (fp32 inputs through the real reference) the loose one
    tb = g["teacher"].bfloat16().double()
    sb = g["student"].bfloat16().double().requires_grad_(True)
    ref_b = vit_ref.dino_loss(tb, sb, g["center"].double(), *g["temps"])
    ref_b.backward()
    assert abs(loss.item() - ref_b.item()) <= 1e-4 * abs(ref_b.item())
    assert abs(loss.item() - g["loss"].item()) <= 5e-3 * abs(g["loss"].item())
    (loss * 65536.0).backward()  # GradScaler-style scaled backward
    assert rel(s.grad / 65536.0, sb.grad) <= 1e-2
    assert rel_l2(s.grad / 65536.0, sb.grad) <= 1e-2


@pytest.mark.parametrize("G,V,B,K", [(2, 8, 6, 16384), (2, 6, 3, 65536), (1, 2, 160, 8192), (2, 3, 5, 4104)])
def test_dino_loss_long_rows_match_the_oracle(G, V, B, K):
    """DINO loss / gradient at head widths up to the BASELINE K = 65536 (online max/sum rescaling over
    thousands of 8-logit groups per thread, ragged K, one global view) against the oracle on the
    same bf16 logits."""
    from vit_core.ssl.dino.loss import DINOLoss
    gen = torch.Generator().manual_seed(G * 100 + V * 10 + B)
    t = (torch.randn(G, B, K, generator=gen) * 2.0).bfloat16()
    s0 = (torch.randn(V, B, K, generator=gen) * 1.5).bfloat16()
    c = torch.randn(1, K, generator=gen) * 0.1
    s = s0.cuda().requires_grad_(True)
    loss = DINOLoss(0.05, 0.1)(t.cuda(), s, c.cuda())
    sb = s0.double().requires_grad_(True)
    ref = vit_ref.dino_loss(t.double(), sb, c.double(), 0.05, 0.1)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-4 * abs(ref.item()), (loss.item(), ref.item())
    loss.backward()
    assert rel_l2(s.grad, sb.grad) <= 1e-2


def test_no_cpu_fallback():
    from vit_core import ViT
    from vit_core._backend.lib import VitsslError
    g = load("vit")
    m = ViT(**g["cfg"])
    with pytest.raises(VitsslError):
        m(g["x"])  # CPU tensors: must fail loudly, never fall back


def test_models_survive_the_reference_compile_wrapper(monkeypatch):
    """utils/model_builder.py:182-183 returns torch.compile(model): the wrapper must call straight
    through our eager regions, keep attribute access (`_orig_mod.` state_dict prefix) and give the
    same numbers as the bare module."""
    g = load("simmim")
    m, mod, idx = _simmim(g)
    from vit_core.ssl.simmim.masking import mask_tables
    monkeypatch.setattr(mod, "draw_mask", lambda B, N, r, dev, want_indices=True: (idx, *mask_tables(idx, N)))
    cm = torch.compile(m)
    assert all(k.startswith("_orig_mod.") for k in cm.state_dict())
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pred, targets = cm(g["x"].cuda())
        loss = torch.nn.L1Loss()(pred, targets)
    loss.backward()
    assert pred.dtype == torch.bfloat16 and rel(pred, g["pred"]) <= ACT_TOL
    assert abs(loss.item() - g["loss"].item()) <= LOSS_TOL * g["loss"].item()
    assert m.simmim_head.weight.grad is not None
    gv = load("vit")
    from vit_core import ViT
    v = ViT(**gv["cfg"])
    v.load_state_dict(gv["weights"])
    v = torch.compile(v.cuda().eval())
    assert rel(v(gv["x"].cuda()), gv["logits"]) <= ACT_TOL


def test_encoder_stack_c_sequencer_matches_per_op_path(monkeypatch):
    """csrc/encoder.cu issues the same kernels as the per-op Python sequencing: outputs must agree bit
    for bit and gradients to fp32 rounding (dropout off), with and without saved activations, and with
    dropout on the two paths must draw the same masks (same seed -> same result)."""
    from vit_core import EncoderBlock
    from vit_core._backend import functional as Fb, ops
    torch.manual_seed(7)
    blocks = torch.nn.ModuleList([EncoderBlock(128, 2, 256, 0.0) for _ in range(3)]).cuda()
    x = torch.randn(5, 37, 128, device="cuda")

    def run(c_path, train=True, p=0.0):
        monkeypatch.setattr(ops, "encoder_stack_supported", (lambda S, D, H: True) if c_path else (lambda S, D, H: False))
        monkeypatch.setattr(Fb, "_new_seed", lambda: 1234)
        for b in blocks:
            b.drop1.p = b.drop2.p = b.feed_forward.dropout.p = p
        blocks.train(train)
        blocks.zero_grad()
        xi = x.clone().requires_grad_(train)
        with torch.set_grad_enabled(train):
            out, probs = Fb.encoder_stack(blocks, xi, return_attn=True)
        if train:
            (out.square().mean() + out.sum() * 1e-3).backward()
        grads = [p_.grad.clone() for p_ in blocks.parameters()] + [xi.grad.clone()] if train else []
        return out.detach().clone(), probs.clone(), grads

    for train, p in ((True, 0.0), (False, 0.0), (True, 0.2)):
        o1, pr1, g1 = run(True, train, p)
        o2, pr2, g2 = run(False, train, p)
        assert torch.equal(o1, o2) and torch.equal(pr1, pr2), (train, p)
        for a, b in zip(g1, g2):   # split-K wgrad accumulates with fp32 atomics: order-dependent last bits
            assert (a - b).abs().max().item() <= 1e-5 * max(b.abs().max().item(), 1e-6), (train, p)


def test_encoder_stack_chunked_backward_matches_single_call(monkeypatch):
    """Under data parallelism the stack backward runs as layer chunks (csrc/encoder.cu layer ranges)
    and hands each chunk's contiguous gradient slice to GradSync.prereduce before the next chunk
    starts. With a stand-in sync object the chunked path must reproduce the single-call gradients,
    and every parameter of the stack must be covered by exactly one slice."""
    from vit_core import EncoderBlock
    from vit_core._backend import functional as Fb
    torch.manual_seed(11)
    L = 5
    blocks = torch.nn.ModuleList([EncoderBlock(128, 2, 256, 0.1) for _ in range(L)]).cuda().train()
    x = torch.randn(3, 50, 128, device="cuda")
    monkeypatch.setattr(Fb, "_new_seed", lambda: 99)

    def run():
        blocks.zero_grad()
        xi = x.clone().requires_grad_(True)
        out = Fb.encoder_stack(blocks, xi)
        out = out[0] if isinstance(out, tuple) else out
        (out.square().mean() + out.sum() * 1e-3).backward()
        return [p_.grad.clone() for p_ in blocks.parameters()] + [xi.grad.clone()]

    ref = run()
    seen = []

    class FakeSync:
        def prereduce(self, flat, params):
            assert flat.is_contiguous() and flat.dtype == torch.float32
            params = list(params)
            assert flat.numel() == sum(p_.numel() for p_ in params)
            seen.extend(id(p_) for p_ in params)

        def join(self):
            pass

    monkeypatch.setattr(Fb.dp, "sync_for", lambda params: FakeSync())
    got = run()
    assert sorted(seen) == sorted(id(p_) for p_ in blocks.parameters())
    assert len(ref) == len(got)
    for a, b in zip(got, ref):
        assert (a - b).abs().max().item() <= 1e-5 * max(b.abs().max().item(), 1e-6)
