"""Runs the UNMODIFIED reference (`baseline/_ref`, installed by baseline/install_reference.py) on the
CPU in float64 for one parity case and writes its outputs to a file. Test infrastructure only.

It runs in its own process because the reference's package is also called `vit_core`:
    python tests/ref_runner.py <case.pt> <out.pt>
`case.pt` holds {"kind": "simmim" | "dino" | "vit", "cfg": ctor kwargs, "state": state_dict (fp32),
inputs ...}. The reference code is not patched; for SimMIM the permutations its `torch.randperm`
calls return are replayed from the case file (the GPU run's draws), because a CPU generator cannot
reproduce the CUDA generator's stream.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import refenv  # noqa: E402


def main(case_path, out_path):
    refenv.use_reference_vit_core()
    import vit_core
    assert vit_core.__file__.startswith(refenv.REF), vit_core.__file__
    case = torch.load(case_path, weights_only=False)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(os.cpu_count() or 1)
    kind, cfg = case["kind"], case["cfg"]
    out = {}
    if kind == "simmim":
        from vit_core.ssl.simmim.model import SimMIMViT
        m = SimMIMViT(**cfg)
        m.load_state_dict({k: v.double() for k, v in case["state"].items()})
        m.train()
        perms = [p for p in case["perms"]]
        real_randperm = torch.randperm

        def replay(n, *a, **k):
            p = perms.pop(0)
            assert p.numel() == n
            return p.clone()

        torch.randperm = replay
        try:
            pred, targets, mask = m(case["x"].double(), return_bool_mask=True)
        finally:
            torch.randperm = real_randperm
        loss = torch.nn.L1Loss()(pred, targets)
        loss.backward()
        out = dict(pred=pred.detach(), targets=targets.detach(), bool_mask=mask.detach(), loss=loss.detach(),
                   grads={k: p.grad for k, p in m.named_parameters()})
        out["inference"] = m.inference_forward(case["x"].double())
    elif kind == "vit":
        from vit_core.vit import ViT
        m = ViT(**cfg)
        m.load_state_dict({k: v.double() for k, v in case["state"].items()})
        m.train()
        logits = m(case["x"].double())
        loss = torch.nn.functional.cross_entropy(logits, case["labels"])
        loss.backward()
        out = dict(logits=logits.detach(), loss=loss.detach(), grads={k: p.grad for k, p in m.named_parameters()})
    elif kind == "dino":
        from vit_core.ssl.dino.loss import DINOLoss
        from vit_core.ssl.dino.model import DINOViT
        m = DINOViT(**cfg)
        m.load_state_dict({k: v.double() for k, v in case["state"].items()})
        m.train()
        views = [v.double() for v in case["views"]]
        G = case["num_global"]
        teacher, student = m(views, G)
        B = views[0].shape[0]
        loss = DINOLoss(*case["temps"])(teacher.view(G, B, -1), student.view(len(views), B, -1), m.center)
        loss.backward()
        out = dict(teacher=teacher.detach(), student=student.detach(), center=m.center.detach().clone(),
                   loss=loss.detach(),
                   grads={k: p.grad for k, p in m.named_parameters() if p.grad is not None})
        m.momentum_update_teacher(case["momentum"])
        out["teacher_after"] = {k: v.detach().clone() for k, v in m.state_dict().items() if k.startswith("teacher_")}
        out["inference"] = m.inference_forward(views[0])
    else:
        raise SystemExit(f"unknown case kind {kind}")
    # large dictionaries travel as float32 (the comparison against the float64 oracle is at 1e-6)
    for k in ("grads", "teacher_after"):
        if k in out:
            out[k] = {n: t.detach().float() for n, t in out[k].items()}
    torch.save(out, out_path)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
