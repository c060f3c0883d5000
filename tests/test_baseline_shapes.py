"""Model-level GPU parity at the BASELINE.json shapes (north_star tolerances).

Each test builds OUR model with a fixed seed, runs the CUDA path through the drop-in API, and
compares activations, loss and EVERY parameter gradient with the float64 oracle (oracle/vit_ref.py)
evaluated on the CPU with the same weights, inputs and mask. Where the unmodified reference is
installed (baseline/_ref, shipped with the repo snapshot) it is also run — float64, CPU, its own
process — and must agree with the oracle to 1e-6, which pins the oracle at these shapes too.

  cfg 1  ViT-Ti/16 supervised, 96x96, B=8, cross-entropy             (vit.py:33-45)
  cfg 2  ViT-S/16 SimMIM, 224x224, mask 0.6, L=12                    (ssl/simmim/model.py:42-62)
  cfg 3  ViT-S/16 DINO, 2x224 + 6x96 crops, K=65536, EMA teacher     (ssl/dino/model.py:110-139)
  cfg 4  ViT-B/16 SimMIM at a token count that runs the GEMMs as CTA pairs (M >= 4096)
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import vit_ref
from parity_utils import (ACT_TOL, LOSS_TOL, autocast_yardstick, check_grads, note, reference_available, rel, rel_l2,
                          run_real_reference)

pytestmark = pytest.mark.gpu

VIT_TI = dict(num_blocks=12, embed_dim=192, num_heads=3, mlp_dim=768, patch_size=16)
VIT_S = dict(num_blocks=12, embed_dim=384, num_heads=6, mlp_dim=1536, patch_size=16)
VIT_B = dict(num_blocks=12, embed_dim=768, num_heads=12, mlp_dim=3072, patch_size=16)


def _perms_from_indices(idx, N):
    """Full-length permutations whose first n_m entries are the drawn indices (what the reference's
    torch.randperm calls returned on the GPU; only the first n_m entries are ever used)."""
    out = []
    for row in idx.cpu():
        rest = torch.tensor(sorted(set(range(N)) - set(row.tolist())), dtype=torch.long)
        out.append(torch.cat([row.long(), rest]))
    return out


def _simmim_case(arch, B, seed, monkeypatch=None, check_reference=True, tag=""):
    from vit_core._backend import functional as Fb
    from vit_core.ssl.simmim import SimMIMViT
    torch.manual_seed(seed)
    cfg = dict(input_shape=(3, 224, 224), dropout=0.0, mask_ratio=0.6, **arch)
    m = SimMIMViT(**cfg)
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.rand(B, 3, 224, 224)
    m.cuda().train()
    N, n_m, P = 196, 117, 768
    kw = dict(patch_size=16, num_blocks=arch["num_blocks"], num_heads=arch["num_heads"])

    # ---- the trainer's call sequence (simmim_trainer.py:65-69): autocast, model(x), nn.L1Loss
    torch.cuda.manual_seed(1000 + seed)
    torch.rand(5, device="cuda")
    rng = torch.cuda.get_rng_state()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pred, targets, bool_mask = m(x.cuda(), return_bool_mask=True)
        loss = torch.nn.L1Loss(reduction="mean")(pred, targets)
    assert pred.shape == (B * n_m, P) and pred.dtype == torch.bfloat16 and targets.dtype == torch.float32
    chain = [type(loss.grad_fn).__name__] + [type(f).__name__ for f, _ in loss.grad_fn.next_functions if f is not None]
    assert any("_L1LossFn" in n for n in chain), chain  # fused kernel behind the PrefetchedScalar's alias node
    loss.backward()
    # mask: bit-exact with the reference's B sequential torch.randperm(N, device=cuda)[:n_m] draws
    torch.cuda.set_rng_state(rng)
    idx = torch.stack([torch.randperm(N, device="cuda")[:n_m] for _ in range(B)])
    ref_mask = torch.zeros(B, N, dtype=torch.bool, device="cuda").scatter_(1, idx, True)
    assert torch.equal(bool_mask.squeeze(-1), ref_mask)
    mask = ref_mask.cpu()

    # ---- oracle, float64 on the CPU
    w = {k: v.double().requires_grad_(True) for k, v in state.items()}
    pred_ref, tg_ref = vit_ref.simmim_forward(w, x.double(), mask, **kw)
    loss_ref = vit_ref.l1_loss(pred_ref, tg_ref)
    loss_ref.backward()
    grads = {k: v.grad for k, v in w.items()}
    assert torch.equal(targets.cpu().double(), tg_ref)                      # raw pixels: bit-exact
    e_pred = rel(pred, pred_ref)
    e_loss = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
    assert e_pred <= ACT_TOL, e_pred
    assert e_loss <= LOSS_TOL, (loss.item(), loss_ref.item())

    def run(wg):
        p_, t_ = vit_ref.simmim_forward(wg, x.cuda(), ref_mask, **kw)
        return vit_ref.l1_loss(p_.float(), t_)
    yard = autocast_yardstick(state, run)
    rep = {}
    worst = check_grads(m.named_parameters(), grads, yard, report=rep)
    over = sum(1 for v in rep.values() if v[1] > 1e-2)
    note(f"simmim{tag} trainer path", pred_max_rel=e_pred, loss_rel=e_loss, worst_grad=worst[0], worst_grad_max_rel=worst[1],
         allowed=worst[2], worst_grad_rel_l2=max(v[0] for v in rep.values()), tensors_over_1e2_maxrel=over, tensors=len(rep))

    # ---- fused objective entry, no autocast: same mask (same generator state), same gradients
    m.zero_grad(set_to_none=True)
    torch.cuda.set_rng_state(rng)
    loss2 = m.reconstruction_loss(x.cuda())
    assert abs(loss2.item() - loss_ref.item()) <= LOSS_TOL * abs(loss_ref.item())
    loss2.backward()
    worst2 = check_grads(m.named_parameters(), grads, yard)

    # ---- data-parallel form of the stack backward (layer chunks + per-chunk gradient slices)
    if monkeypatch is not None:
        seen = []

        class FakeSync:
            def prereduce(self, flat, params):
                seen.extend(id(p_) for p_ in params)

            def join(self):
                pass

        monkeypatch.setattr(Fb.dp, "sync_for", lambda params: FakeSync())
        m.zero_grad(set_to_none=True)
        torch.cuda.set_rng_state(rng)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            p3, t3 = m(x.cuda())
            l3 = torch.nn.L1Loss()(p3, t3)
        l3.backward()
        assert sorted(seen) == sorted(id(p_) for p_ in m.encoder_blocks.parameters())
        check_grads(m.named_parameters(), grads, yard)
        monkeypatch.undo()

    feats = m.inference_forward(x.cuda())
    inf_ref = vit_ref.simmim_inference({k: v.detach() for k, v in w.items()}, x.double(), **kw)
    assert not m.training and rel(feats, inf_ref) <= ACT_TOL
    note(f"simmim{tag} fused-objective path", worst_grad=worst2[0], worst_grad_max_rel=worst2[1], inference_max_rel=rel(feats, inf_ref))

    # ---- the unmodified reference itself (float64, CPU) must agree with the oracle
    if check_reference and reference_available():
        out = run_real_reference(dict(kind="simmim", cfg=cfg, state=state, x=x, perms=_perms_from_indices(idx, N)))
        assert torch.equal(out["bool_mask"].squeeze(-1), mask) and torch.equal(out["targets"], tg_ref)
        assert rel(out["pred"], pred_ref) <= 1e-9 and abs(out["loss"].item() - loss_ref.item()) <= 1e-12
        e = max(rel_l2(out["grads"][k], grads[k]) for k in grads)
        assert e <= 1e-6, e
        assert rel(out["inference"], inf_ref) <= 1e-9
        note(f"simmim{tag} oracle vs REAL reference (fp64)", worst_grad_rel_l2=e)


def test_simmim_vit_s16_224(monkeypatch):
    _simmim_case(VIT_S, B=4, seed=11, monkeypatch=monkeypatch, tag=" ViT-S/16 224 B=4")


def test_simmim_vit_b16_cta_pair_gemms():
    # 21 images x 196 tokens = 4116 rows >= 4096: the QKV / FFN GEMMs take the cta_group::2 path
    _simmim_case(VIT_B, B=21, seed=13, check_reference=False, tag=" ViT-B/16 224 B=21 (CTA pairs)")


def test_vit_tiny_supervised_96():
    from vit_core import ViT
    torch.manual_seed(17)
    cfg = dict(num_classes=10, input_shape=(3, 96, 96), dropout=0.0, **VIT_TI)
    m = ViT(**cfg)
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x, labels = torch.rand(8, 3, 96, 96), torch.randint(0, 10, (8,))
    m.cuda().train()
    kw = dict(patch_size=16, num_blocks=12, num_heads=3)
    with torch.autocast("cuda", dtype=torch.bfloat16):                       # supervised_trainer.py:34-36
        logits = m(x.cuda())
        loss = F.cross_entropy(logits, labels.cuda())
    loss.backward()
    w = {k: v.double().requires_grad_(True) for k, v in state.items()}
    logits_ref, _ = vit_ref.vit_forward(w, x.double(), **kw)
    loss_ref = F.cross_entropy(logits_ref, labels)
    loss_ref.backward()
    grads = {k: v.grad for k, v in w.items()}
    e_logits = rel(logits, logits_ref)
    e_loss = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
    yard = autocast_yardstick(state, lambda wg: F.cross_entropy(vit_ref.vit_forward(wg, x.cuda(), **kw)[0].float(), labels.cuda()))
    ly = None
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ly = vit_ref.vit_forward({k: v.cuda() for k, v in state.items()}, x.cuda(), **kw)[0]
    # logits of 10 classes after 12 bf16 layers: allow what the reference under autocast shows
    assert e_logits <= max(ACT_TOL, 2 * rel(ly, logits_ref)), (e_logits, rel(ly, logits_ref))
    assert e_loss <= max(LOSS_TOL, 2 * abs(F.cross_entropy(ly.float(), labels.cuda()).item() - loss_ref.item()) / abs(loss_ref.item())), e_loss
    rep = {}
    worst = check_grads(m.named_parameters(), grads, yard, report=rep)
    note("vit-ti/16 96 B=8 supervised", logits_max_rel=e_logits, autocast_ref_logits_max_rel=rel(ly, logits_ref), loss_rel=e_loss,
         worst_grad=worst[0], worst_grad_max_rel=worst[1], allowed=worst[2], worst_grad_rel_l2=max(v[0] for v in rep.values()))
    if reference_available():
        out = run_real_reference(dict(kind="vit", cfg=cfg, state=state, x=x, labels=labels))
        assert rel(out["logits"], logits_ref) <= 1e-9 and abs(out["loss"].item() - loss_ref.item()) <= 1e-12
        assert max(rel_l2(out["grads"][k], grads[k]) for k in grads) <= 1e-6


def test_dino_vit_s16_multicrop_k65536():
    from vit_core.ssl.dino import DINOViT
    from vit_core.ssl.dino.loss import DINOLoss
    torch.manual_seed(19)
    K, B, G, V = 65536, 2, 2, 8
    cfg = dict(input_shape=(3, 224, 224), dropout=0.0, output_dim=K, center_momentum=0.9, **VIT_S)
    m = DINOViT(**cfg)
    with torch.no_grad():  # teacher != student, center != 0: not the degenerate initial state
        for p in list(m.teacher_backbone.parameters()) + list(m.teacher_head.parameters()):
            p.add_(0.01 * torch.randn_like(p))
        m.center.copy_(0.05 * torch.randn_like(m.center))
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    views = [torch.rand(B, 3, 224, 224) for _ in range(G)] + [torch.rand(B, 3, 96, 96) for _ in range(V - G)]
    temps, momentum = (0.04, 0.1), 0.996
    m.cuda().train()
    crit = DINOLoss(*temps)
    with torch.autocast("cuda", dtype=torch.bfloat16):                       # dino_trainer.py:86-99
        teacher, student = m([v.cuda() for v in views], G)
        loss = crit(teacher.view(G, B, K), student.view(V, B, K), m.center)
    assert teacher.shape == (G * B, K) and student.shape == (V * B, K) and not teacher.requires_grad
    loss.backward()

    kw = dict(patch_size=16, num_blocks=12, num_heads=6, grid=(14, 14))
    w = {k: v.double() for k, v in state.items()}
    for k in w:
        if k.startswith("student_"):
            w[k].requires_grad_(True)
    t_ref, s_ref, c_ref = vit_ref.dino_forward(w, [v.double() for v in views], G, w["center"], center_momentum=0.9, **kw)
    loss_ref = vit_ref.dino_loss(t_ref.view(G, B, K).detach(), s_ref.view(V, B, K), c_ref.detach(), *temps)
    loss_ref.backward()
    grads = {k: v.grad for k, v in w.items() if v.grad is not None}

    def run(wg):
        t_, s_, c_ = vit_ref.dino_forward(wg, [v.cuda() for v in views], G, wg["center"], center_momentum=0.9, **kw)
        return vit_ref.dino_loss(t_.view(G, B, K).detach(), s_.view(V, B, K), c_.detach().float(), *temps)
    wy = {k: v.cuda().float().clone().requires_grad_(k.startswith("student_")) for k, v in state.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ly = run(wy)
    ly.backward()
    yard = {k: v.grad for k, v in wy.items() if v.grad is not None}
    e_t, e_s, e_c = rel(teacher, t_ref), rel(student, s_ref), rel(m.center, c_ref)
    e_loss = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
    y_loss = abs(ly.item() - loss_ref.item()) / abs(loss_ref.item())
    assert e_t <= ACT_TOL and e_s <= ACT_TOL and e_c <= ACT_TOL, (e_t, e_s, e_c)
    assert m.center.shape == (1, K)
    assert e_loss <= max(LOSS_TOL, 2 * y_loss), (e_loss, y_loss)
    rep = {}
    worst = check_grads(m.named_parameters(), grads, yard, report=rep)      # every student gradient, same bar as SimMIM
    for k, p in m.named_parameters():
        if k.startswith("teacher_"):
            assert p.grad is None, k
    over = sum(1 for v in rep.values() if v[1] > 1e-2)
    note("dino ViT-S/16 2x224+6x96 K=65536 B=2", teacher_max_rel=e_t, student_max_rel=e_s, center_max_rel=e_c, loss_rel=e_loss,
         autocast_ref_loss_rel=y_loss, worst_grad=worst[0], worst_grad_max_rel=worst[1], allowed=worst[2],
         worst_grad_rel_l2=max(v[0] for v in rep.values()), tensors_over_1e2_maxrel=over, tensors=len(rep))

    # EMA teacher update (dino_trainer.py:105) and the evaluation entry (model.py:141-155): VALUES
    m.momentum_update_teacher(momentum)
    sd = m.state_dict()
    w_after = {k: v.detach() for k, v in w.items()}
    worst_ema = 0.0
    for k in state:
        if k.startswith("teacher_"):
            ks = "student_" + k[len("teacher_"):]
            w_after[k] = momentum * w[k].detach() + (1 - momentum) * w[ks].detach()
            worst_ema = max(worst_ema, (sd[k].double().cpu() - w_after[k]).abs().max().item() / max(1.0, w_after[k].abs().max().item()))
    assert worst_ema <= 1e-6, worst_ema
    out = m.inference_forward(views[0].cuda())
    inf_ref = vit_ref.dino_head(w_after, "teacher_head.", vit_ref.dino_backbone(w_after, "teacher_backbone.", views[0].double(), **kw))
    assert not m.training and out.shape == (B, K)
    e_inf = rel(out, inf_ref)
    assert e_inf <= ACT_TOL, e_inf
    feats = m.inference_forward(views[0].cuda(), return_features=True)
    assert rel(feats, vit_ref.dino_backbone(w_after, "teacher_backbone.", views[0].double(), **kw)) <= ACT_TOL
    note("dino EMA + inference_forward", teacher_after_max_err=worst_ema, inference_max_rel=e_inf)

    if reference_available():
        r = run_real_reference(dict(kind="dino", cfg=cfg, state=state, views=views, num_global=G, temps=temps, momentum=momentum))
        assert rel(r["teacher"], t_ref) <= 1e-9 and rel(r["student"], s_ref) <= 1e-9 and rel(r["center"], c_ref) <= 1e-9
        assert abs(r["loss"].item() - loss_ref.item()) <= 1e-12 * max(1.0, abs(loss_ref.item())) + 1e-15
        e = max(rel_l2(r["grads"][k], grads[k]) for k in grads)
        assert e <= 1e-6, e
        assert rel(r["inference"], inf_ref) <= 1e-9
        assert max(rel(r["teacher_after"][k], w_after[k]) for k in r["teacher_after"] if k in w_after and k.startswith("teacher_")) <= 1e-6
        note("dino oracle vs REAL reference (fp64)", worst_grad_rel_l2=e)
