"""CPU: pin the oracle restatement (oracle/vit_ref.py) against golden vectors produced by the REAL
reference (oracle/make_golden.py). float64 throughout, so agreement is to rounding."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import vit_ref
from oracle.cases import build_dino_case, digest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def dbl(w):
    return {k: v.double() for k, v in w.items()}


def close(a, b, tol=1e-9):
    return (a.double() - b.double()).abs().max().item() <= tol * max(1.0, b.double().abs().max().item())


def test_encoder_block_matches_reference():
    g = load("encoder_block")
    w = {k: v.double().requires_grad_(True) for k, v in g["weights"].items()}
    x = g["x"].double().requires_grad_(True)
    y, probs = vit_ref.encoder_block(w, "", x, g["cfg"]["num_heads"])
    assert close(y, g["y"]) and close(probs, g["probs"])
    (y * g["dy"].double()).sum().backward()
    assert close(x.grad, g["dx"])
    for k, ref in g["grads"].items():
        assert close(w[k].grad, ref, 1e-6), k


def test_mha_cross_attention_matches_reference():
    g = load("mha_cross")
    out, probs = vit_ref.multi_head_attention(dbl(g["weights"]), "", g["q"].double(), g["k"].double(),
                                              g["v"].double(), 2)
    assert close(out, g["out"]) and close(probs, g["probs"])


def test_ffn_matches_reference():
    g = load("ffn")
    assert close(vit_ref.feed_forward(dbl(g["weights"]), "", g["x"].double()), g["y"])


def test_patch_embeddings_match_reference():
    g = load("patch_embeddings")
    x = g["x"].double()
    assert close(vit_ref.conv_patch_embedding(dbl(g["conv"]["weights"]), "", x, 8), g["conv"]["y"])
    assert close(vit_ref.manual_patch_embedding(dbl(g["manual"]["weights"]), "", x, 8), g["manual"]["y"])
    wd = dbl(g["dynamic"]["weights"])
    assert close(vit_ref.dynamic_patch_embedding(wd, "", x, 8, (4, 4)), g["dynamic"]["y"])
    assert close(vit_ref.dynamic_patch_embedding(wd, "", g["dynamic"]["x_local"].double(), 8, (4, 4)),
                 g["dynamic"]["y_local"])
    with pytest.raises(ValueError):
        vit_ref.dynamic_patch_embedding(wd, "", torch.rand(1, 3, 30, 32).double(), 8, (4, 4))


def test_vit_matches_reference():
    g = load("vit")
    w = {k: v.double().requires_grad_(True) for k, v in g["weights"].items()}
    logits, probs = vit_ref.vit_forward(w, g["x"].double(), patch_size=8, num_blocks=2, num_heads=2)
    assert close(logits, g["logits"]) and close(probs, g["probs"])
    loss = F.cross_entropy(logits, g["labels"])
    assert close(loss, g["loss"])
    loss.backward()
    for k, ref in g["grads"].items():
        assert close(w[k].grad, ref, 1e-6), k


def test_simmim_matches_reference_including_mask():
    g = load("simmim")
    mask = vit_ref.mask_from_perms(g["perms"], 16, 0.6)
    assert torch.equal(mask, g["bool_mask"])  # bit-exact mask from the recorded permutations
    w = {k: v.double().requires_grad_(True) for k, v in g["weights"].items()}
    pred, targets = vit_ref.simmim_forward(w, g["x"].double(), mask, patch_size=8, num_blocks=2, num_heads=2)
    assert close(pred, g["pred"]) and torch.equal(targets, g["targets"])
    loss = vit_ref.l1_loss(pred, targets)
    assert close(loss, g["loss"])
    loss.backward()
    for k, ref in g["grads"].items():
        assert close(w[k].grad, ref, 1e-6), k
    feats = vit_ref.simmim_inference(dbl(g["weights"]), g["x"].double(), patch_size=8, num_blocks=2, num_heads=2)
    assert close(feats, g["inference"])


def test_dino_matches_reference():
    import vit_core.ssl.dino.model as ours  # construction only (CPU): same seed -> same weights
    g = load("dino")
    cfg, m, views, B = build_dino_case(ours.DINOViT)
    sd = m.state_dict()
    for k, dg in g["weight_digests"].items():
        mine = digest(sd[k])
        assert mine["shape"] == dg["shape"] and abs(mine["norm"] - dg["norm"]) <= 1e-9 * max(1.0, dg["norm"]), k
        assert torch.equal(mine["sample"], dg["sample"]), k
    w = {k: v.double() for k, v in sd.items()}
    for k in w:
        if k.startswith("student_"):
            w[k].requires_grad_(True)
    teacher, student, center = vit_ref.dino_forward(
        w, [v.double() for v in views], 2, w["center"], patch_size=8, num_blocks=2, num_heads=2, grid=(4, 4),
        center_momentum=cfg["center_momentum"])
    assert close(teacher, g["teacher"]) and close(student, g["student"]) and close(center, g["center_after"])
    K = teacher.shape[1]
    tt, ts = g["temps"]
    loss = vit_ref.dino_loss(teacher.detach().view(2, B, K), student.view(4, B, K), center, tt, ts)
    assert close(loss, g["loss"])
    assert close(vit_ref.dino_loss_factorised(teacher.detach().view(2, B, K), student.view(4, B, K), center, tt, ts),
                 g["loss"])
    loss.backward()
    for k, dg in g["grad_digests"].items():
        mine = digest(w[k].grad)
        assert abs(mine["norm"] - dg["norm"]) <= 1e-8 * max(1e-12, dg["norm"]), k
        assert close(mine["sample"], dg["sample"], 1e-8), k
    # EMA (model.py:126-139)
    t_keys = [k for k in sd if k.startswith("teacher_") and not k.endswith("center")]
    new_t = vit_ref.ema_update([w[k].detach() for k in t_keys],
                               [w[k.replace("teacher_", "student_", 1)].detach() for k in t_keys], g["momentum"])
    for k, t in zip(t_keys, new_t):
        dg = g["teacher_after_digests"][k]
        assert close(digest(t)["sample"], dg["sample"], 1e-12), k


def test_dino_loss_and_closed_form_gradient():
    g = load("dino_loss")
    s = g["student"].double().requires_grad_(True)
    tt, ts = g["temps"]
    loss = vit_ref.dino_loss(g["teacher"].double(), s, g["center"].double(), tt, ts)
    assert close(loss, g["loss"])
    loss.backward()
    assert close(s.grad, g["dstudent"])
    # closed form
    G, B, K = g["teacher"].shape
    pbar = F.softmax((g["teacher"].double() - g["center"].double()) / tt, dim=-1).sum(0)
    closed = -(pbar.unsqueeze(0) - G * F.softmax(s.detach() / ts, dim=-1)) / (ts * G * B * K)
    assert close(closed, g["dstudent"], 1e-9)


def test_schedulers_match_reference_formulae():
    from vit_core.ssl.dino.dino_utils import DINOMomentumScheduler, DINOTeacherTempScheduler
    ms = DINOMomentumScheduler(0.996, 1.0, 50)
    ts_c = DINOTeacherTempScheduler(0.04, 0.07, 30, "cosine")
    ts_l = DINOTeacherTempScheduler(0.04, 0.07, 30, "linear")
    for step in (0, 1, 7, 29, 30, 49, 50, 80):
        assert ms.get_momentum(step) == pytest.approx(vit_ref.momentum_schedule(0.996, 1.0, 50, step), abs=1e-15)
        assert ts_c.get_temp(step) == pytest.approx(vit_ref.teacher_temp_schedule(0.04, 0.07, 30, step), abs=1e-15)
        assert ts_l.get_temp(step) == pytest.approx(vit_ref.teacher_temp_schedule(0.04, 0.07, 30, step, "linear"), abs=1e-15)


def test_randperm_oracle_matches_torch_cuda_draws():
    """oracle/randperm_ref.py (numpy restatement of torch.randperm on CUDA, which defines the SimMIM
    mask: masking.py:22-25) against permutations torch itself drew on a B200
    (tests/golden/randperm_cuda.pt, made by oracle/make_randperm_golden.py)."""
    import numpy as np
    from oracle import randperm_ref as R
    g = torch.load(os.path.join(GOLD, "randperm_cuda.pt"), weights_only=False)
    assert len(g["cases"]) >= 9
    for c in g["cases"]:
        off = c["offset"]
        for r in range(min(c["reps"], 48)):
            perm, off = R.cuda_randperm(c["n"], c["seed"], off)
            assert np.array_equal(perm, c["perms"][r].numpy().astype(np.int64)), (c["n"], c["seed"], r)
        if c["reps"] <= 48:
            assert off == c["offset_after"]
        assert c["offset"] + c["reps"] * R.offset_per_call(c["n"]) == c["offset_after"]
    idx, off = R.simmim_mask_indices(3, 196, 117, 42, 0)
    assert np.array_equal(idx, g["cases"][0]["perms"][:3, :117].numpy()) and off == 600
