"""Shared helpers of the GPU parity tests (north-star tolerances: max relative error <= 1e-2 on
activations and gradients with bf16 compute / fp32 accumulation, <= 1e-3 on losses)."""
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACT_TOL, GRAD_TOL, LOSS_TOL = 1e-2, 1e-2, 1e-3


def rel(a, b):
    b = b.double().cpu()
    return ((a.detach().double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    b = b.double().cpu()
    return ((a.detach().double().cpu() - b).norm() / b.norm().clamp_min(1e-30)).item()


def check_grads(named_params, ref_grads, yard=None, tol=GRAD_TOL, report=None):
    """Every parameter gradient: relative L2 error <= tol and max-abs relative error <= tol — or, for
    tensors where bf16 rounding alone exceeds that, <= 2x the error the reference algorithm itself
    shows under autocast(bf16) on the same GPU (`yard`, computed with the oracle). Returns the worst
    (name, max-rel, allowed); `report` (dict) collects every tensor's errors."""
    worst = ("", 0.0, 0.0)
    n = 0
    failures = []
    for k, p in named_params:
        if k not in ref_grads:
            continue
        n += 1
        assert p.grad is not None, k
        e2, e = rel_l2(p.grad, ref_grads[k]), rel(p.grad, ref_grads[k])
        allowed2 = allowed = tol
        if yard is not None and k in yard:
            allowed2 = max(tol, 2.0 * rel_l2(yard[k], ref_grads[k]))
            allowed = max(tol, 2.0 * rel(yard[k], ref_grads[k]))
        if report is not None:
            report[k] = (e2, e, allowed2, allowed)
        if e2 > allowed2:
            failures.append((k, "rel-L2", round(e2, 5), "allowed", round(allowed2, 5)))
        if e > allowed:
            failures.append((k, "max-rel", round(e, 5), "allowed", round(allowed, 5)))
        if e > worst[1]:
            worst = (k, e, allowed)
    assert n > 0, "no gradient was compared"
    if failures:
        d = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(d):
            with open(os.path.join(d, "parity_r2.log"), "a") as f:
                f.write("GRADIENT FAILURES: " + repr(failures) + "\n")
    assert not failures, failures[:12]
    return worst


def autocast_yardstick(weights, run):
    """Gradients of the reference algorithm (oracle) under autocast(bf16) on this GPU."""
    w = {k: v.cuda().float().clone().requires_grad_(v.is_floating_point()) for k, v in weights.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = run(w)
    loss.backward()
    return {k: v.grad for k, v in w.items() if v.grad is not None}


def reference_available():
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "vit_core"))


def run_real_reference(case):
    """The UNMODIFIED reference (baseline/_ref) on this case, CPU float64, in its own process
    (tests/ref_runner.py). Returns its output dict."""
    with tempfile.TemporaryDirectory() as tmp:
        cin, cout = os.path.join(tmp, "case.pt"), os.path.join(tmp, "out.pt")
        torch.save(case, cin)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_runner.py"), cin, cout],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        return torch.load(cout, weights_only=False)


def note(name, **vals):
    """Measured worst-case errors, printed with -s and appended to gpurun_out/parity_r2.log when that
    directory exists (DESIGN.md §4 quotes them)."""
    line = name + ": " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items())
    print(line)
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_r2.log"), "a") as f:
            f.write(line + "\n")
