"""Import environment for the UNMODIFIED reference in `baseline/_ref` (test / bench infrastructure).

The reference's `utils` package imports three observability dependencies that are not in this image
(`ignite`, `torcheval`, `matplotlib`). They are only used for once-per-epoch PSNR/SSIM
and plotting, so stand-in modules are inserted into `sys.modules` before `utils` is imported. With
them the reference's real `SimMIMTrainer` / `DINOTrainer` / `SupervisedTrainer`, `make_optimizer`,
`make_criterion`, `build_model` import unmodified.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "vit_core")) and os.path.isdir(os.path.join(REF, "utils"))


def stub_observability() -> None:
    if "ignite.metrics" not in sys.modules:
        ignite = types.ModuleType("ignite")
        metrics = types.ModuleType("ignite.metrics")

        class SSIM:
            """Stand-in for ignite.metrics.SSIM (once-per-epoch observability, out of scope): global-
            statistics SSIM per image, averaged — same call protocol (update((pred, target)), compute())."""

            def __init__(self, data_range=1.0, **k):
                self.c1, self.c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
                self.vals = []

            def reset(self):
                self.vals = []

            def update(self, pair):
                x, y = (t.detach().float().flatten(1) for t in pair)
                mx, my = x.mean(1), y.mean(1)
                vx, vy = x.var(1, unbiased=False), y.var(1, unbiased=False)
                cov = ((x - mx[:, None]) * (y - my[:, None])).mean(1)
                s = ((2 * mx * my + self.c1) * (2 * cov + self.c2)) / ((mx * mx + my * my + self.c1) * (vx + vy + self.c2))
                self.vals.append(s.mean().item())

            def compute(self):
                return sum(self.vals) / max(len(self.vals), 1)

        metrics.SSIM = SSIM
        ignite.metrics = metrics
        sys.modules["ignite"], sys.modules["ignite.metrics"] = ignite, metrics
    if "torcheval.metrics" not in sys.modules:
        te = types.ModuleType("torcheval")
        tm = types.ModuleType("torcheval.metrics")

        class PeakSignalNoiseRatio:
            """Stand-in for torcheval.metrics.PeakSignalNoiseRatio: 10 log10(range^2 / MSE)."""

            def __init__(self, data_range=1.0, **k):
                self.r2, self.se, self.n = float(data_range) ** 2, 0.0, 0

            def reset(self):
                self.se, self.n = 0.0, 0

            def update(self, pred, target):
                d = pred.detach().float() - target.detach().float()
                self.se += d.square().sum().item()
                self.n += d.numel()

            def compute(self):
                import math
                return 10.0 * math.log10(self.r2 / max(self.se / max(self.n, 1), 1e-20))

        tm.PeakSignalNoiseRatio = PeakSignalNoiseRatio
        te.metrics = tm
        sys.modules["torcheval"], sys.modules["torcheval.metrics"] = te, tm
    if "matplotlib.pyplot" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt


def use_reference_vit_core() -> None:
    """Make `import vit_core` / `import utils` resolve to the reference (for the reference arm).
    Must run in a process that has not imported our `vit_core`."""
    if "vit_core" in sys.modules and not sys.modules["vit_core"].__file__.startswith(REF):
        raise RuntimeError("our vit_core is already imported in this process")
    stub_observability()
    pkg = os.path.join(ROOT, "vit-ssl_b200")
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != pkg]
    sys.path.insert(0, REF)


def use_reference_callers_over_our_vit_core() -> None:
    """`utils`, `data`, `evaluators` from the reference; `vit_core` from this repo (drop-in test)."""
    stub_observability()
    pkg = os.path.join(ROOT, "vit-ssl_b200")
    for p in (REF, pkg):  # pkg ends up first
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


class AttrDict(dict):
    """Stand-in for the OmegaConf DictConfig the trainers receive: item, attribute and .get access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    @staticmethod
    def wrap(o):
        if isinstance(o, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in o.items()})
        if isinstance(o, (list, tuple)):
            return [AttrDict.wrap(v) for v in o]
        return o
