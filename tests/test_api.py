"""CPU: the C-ABI library loads and exports what include/vitssl_b200.h declares; the drop-in
package keeps the reference's public names, constructor validation and state_dict schema."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from vit_core._backend import lib
    l = lib.lib()
    hdr = open(os.path.join(ROOT, "include", "vitssl_b200.h")).read()
    names = set(re.findall(r"\b(vitssl_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in names:
        assert hasattr(l, n), f"{n} declared in the header but not exported"
    assert l.vitssl_version() >= 100
    assert isinstance(l.vitssl_last_error(), bytes)


def test_ctypes_signatures_match_header_arity():
    from vit_core._backend import lib
    hdr = open(os.path.join(ROOT, "include", "vitssl_b200.h")).read()
    for name, sig in lib.SIGNATURES.items():
        m = re.search(r"int\s+" + name + r"\s*\((.*?)\);", hdr, re.S)
        assert m, name
        nargs = len([a for a in m.group(1).split(",") if a.strip()])
        assert nargs == len(sig), (name, nargs, len(sig))


def test_public_names_and_constructor_validation():
    import vit_core
    from vit_core import (ConvolutionalPatchEmbedding, DynamicPatchEmbedding, EncoderBlock, FeedForwardBlock,  # noqa: F401
                          ManualPatchEmbedding, MultiHeadedAttention, ScaledDotProductAttention, ViT)
    from vit_core.mlp_head import MLPHead  # noqa: F401
    from vit_core.ssl import DINOViT  # noqa: F401
    from vit_core.ssl.dino import DINOHead, DINOMomentumScheduler, DINOTeacherTempScheduler  # noqa: F401
    from vit_core.ssl.dino.loss import DINOLoss  # noqa: F401
    from vit_core.ssl.dino.model import ViTBackbone  # noqa: F401
    from vit_core.ssl.simmim import SimMIMViT, simple_masking  # noqa: F401
    with pytest.raises(ValueError):
        ConvolutionalPatchEmbedding((3, 30, 32), 64, 8)
    with pytest.raises(ValueError):
        ManualPatchEmbedding((3, 32, 30), 64, 8)
    with pytest.raises(AssertionError):
        MultiHeadedAttention(65, 8)
    assert vit_core.__all__


def test_state_dict_schema_matches_reference():
    from vit_core import ViT
    from vit_core.ssl import DINOViT
    from vit_core.ssl.simmim import SimMIMViT
    g = torch.load(os.path.join(ROOT, "tests", "golden", "vit.pt"), weights_only=False)
    m = ViT(**g["cfg"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in g["weights"].items()}
    g = torch.load(os.path.join(ROOT, "tests", "golden", "simmim.pt"), weights_only=False)
    s = SimMIMViT(**g["cfg"])
    assert {k: tuple(v.shape) for k, v in s.state_dict().items()} == {k: tuple(v.shape) for k, v in g["weights"].items()}
    g = torch.load(os.path.join(ROOT, "tests", "golden", "dino.pt"), weights_only=False)
    d = DINOViT(**g["cfg"])
    assert {k: tuple(v.shape) for k, v in d.state_dict().items()} == {k: dg["shape"] for k, dg in g["weight_digests"].items()}
    assert all(not p.requires_grad for p in d.teacher_backbone.parameters())
    assert all(not p.requires_grad for p in d.teacher_head.parameters())
    assert all(p.requires_grad for p in d.student_head.parameters())


def test_same_seed_gives_reference_initialisation():
    """Construction order == RNG order: same seed -> the reference's weights."""
    from vit_core import ViT
    from vit_core.ssl.simmim import SimMIMViT
    g = torch.load(os.path.join(ROOT, "tests", "golden", "vit.pt"), weights_only=False)
    torch.manual_seed(105)
    m = ViT(**g["cfg"])
    for k, v in m.state_dict().items():
        assert torch.equal(v, g["weights"][k]), k
    g = torch.load(os.path.join(ROOT, "tests", "golden", "simmim.pt"), weights_only=False)
    torch.manual_seed(106)
    s = SimMIMViT(**g["cfg"])
    for k, v in s.state_dict().items():
        assert torch.equal(v, g["weights"][k]), k


def test_mask_tables_match_oracle_on_cpu():
    from oracle import vit_ref
    from vit_core.ssl.simmim.masking import mask_tables
    torch.manual_seed(3)
    B, N, r = 6, 49, 0.6
    perms = torch.stack([torch.randperm(N) for _ in range(B)])
    bool_mask, rows, inv = mask_tables(perms[:, : int(N * r)], N)
    assert torch.equal(bool_mask, vit_ref.mask_from_perms(perms, N, r))
    assert torch.equal(rows.long(), bool_mask.reshape(-1).nonzero().squeeze(1))
    assert torch.equal(inv[rows.long()], torch.arange(rows.numel(), dtype=torch.int32))
    # edge cases: ratio 0 -> nothing masked; ratio 1 -> everything
    bm0, rows0, inv0 = mask_tables(perms[:, :0], N)
    assert bm0.sum() == 0 and rows0.numel() == 0 and (inv0 == -1).all()
    bm1, rows1, _ = mask_tables(perms, N)
    assert bm1.all() and rows1.numel() == B * N


def test_modules_survive_deepcopy_and_state_roundtrip():
    import copy
    from vit_core.ssl.dino.model import ViTBackbone
    b = ViTBackbone(1, (3, 16, 16), 64, 8, 1, 128, 0.0)
    c = copy.deepcopy(b)
    c.load_state_dict(b.state_dict())
    assert all(torch.equal(x, y) for x, y in zip(b.state_dict().values(), c.state_dict().values()))


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(ROOT, "vit-ssl_b200")
    for dp_, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp_, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"


def test_committed_launch_list_reproduces_the_traffic_table(tmp_path):
    """profiles/traffic.json (read by bench.py for `roofline.traffic`) must be what
    scripts/summarize_launches.py derives from the committed ncu launch list."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "traffic.json"
    subprocess.run([sys.executable, os.path.join(root, "scripts", "summarize_launches.py"),
                    os.path.join(root, "profiles", "r2_launches.csv"), str(out)], check=True, capture_output=True)
    got, want = json.load(open(out)), json.load(open(os.path.join(root, "profiles", "traffic.json")))
    for fam in ("gemm_tcgen05_kernel", "ln_kernel", "attn_bwd_kernel", "attn_fwd_kernel"):
        assert got[fam] == want[fam], fam
        assert want[fam]["launches"] > 0 and want[fam]["dram_bytes_per_launch"] > 1e6


def test_prefetch_scalar_is_identity_off_gpu():
    from vit_core._backend.scalar import prefetch_scalar
    t = torch.ones(())
    assert prefetch_scalar(t) is t
    assert prefetch_scalar(3.0) == 3.0
