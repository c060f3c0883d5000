#!/usr/bin/env python
"""Install the UNMODIFIED reference (kristi700/ViT-SSL) into the git-ignored `baseline/_ref/`.

`baseline/_ref` is what `bench.py --impl reference` and `tests/test_reference_trainers.py` import:
it travels to the GPU box with the gpurun snapshot (git-ignored, not gpurun-ignored), whereas
`/root/reference` exists only in the build container.

Recipe (outcome recorded in DESIGN.md §6):
 1. `python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse
    --target baseline/_ref <copy of /root/reference>` — the reference ships no setup.py /
    pyproject.toml, so pip refuses ("neither 'setup.py' nor 'pyproject.toml' found");
 2. fall back to what such an install would have produced: the reference's importable top-level
    packages and entry script copied verbatim (`vit_core`, `utils`, `data`, `evaluators`,
    `configs`, `train.py`). Nothing is patched.
A manifest with the sha256 of every copied file is written next to them so tests can assert the
tree is the reference's, byte for byte.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
SRC = os.environ.get("VITSSL_REFERENCE_SRC", "/root/reference")
ITEMS = ("vit_core", "utils", "data", "evaluators", "configs", "train.py")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(verbose=True) -> str:
    if not os.path.isdir(SRC):
        if os.path.isdir(os.path.join(DEST, "vit_core")):
            return "present (source tree absent on this box; using the shipped copy)"
        return f"unavailable: {SRC} does not exist and baseline/_ref was not shipped"
    os.makedirs(DEST, exist_ok=True)
    outcome = "pip: not attempted"
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        r = subprocess.run(
            [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
             "--find-links", "/opt/wheelhouse", "--target", DEST, work],
            capture_output=True, text=True)
        if r.returncode == 0:
            outcome = "pip: installed"
        else:
            tail = (r.stderr.strip().splitlines() or ["?"])[-1]
            outcome = f"pip: failed ({tail[:160]}); copied the package tree verbatim instead"
    manifest = {}
    for item in ITEMS:
        s, d = os.path.join(SRC, item), os.path.join(DEST, item)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__"))
            for dp, _, fs in os.walk(d):
                for f in fs:
                    p = os.path.join(dp, f)
                    manifest[os.path.relpath(p, DEST)] = _sha(p)
        elif os.path.isfile(s):
            shutil.copy2(s, d)
            manifest[item] = _sha(d)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "outcome": outcome, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"[baseline/_ref] {outcome}; {len(manifest)} files")
    return outcome


if __name__ == "__main__":
    print(install())
