"""TEST / BASELINE INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference modules (``train/model.py``, ``utils.py``,
``train.py``, ``evaluate.py``) from ``/root/reference`` (build container) or from the staged copy ``oracle/_ref`` (GPU box),
with the two harness-side shims of SURVEY.md App. A and nothing else:

  1. stub modules for the wheels the reference imports at module level but that are absent here (matplotlib, seaborn,
     albumentations) -- plotting / augmentation, not on the path;
  2. ``weights_backbone=None`` so that torchvision does not try to download ImageNet weights (train/model.py:35).

The reference scripts import their siblings by bare name (``from model import ...``), so the modules are registered in
``sys.modules`` under those names.  ``shim=...`` replaces ``model`` / ``utils`` by this package's drop-in modules BEFORE
``train`` / ``evaluate`` are imported: that is exactly how a user switches.
"""
from __future__ import annotations

import functools
import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_BARE = ("model", "utils", "config", "train", "evaluate", "dataset")


def ref_dir(staged_only: bool = False) -> str | None:
    """``staged_only``: only the byte-identical copy under oracle/_ref (bench.py and the GPU tests must not read
    /root/reference at run time; the staged copy is what travels to the GPU box)."""
    cands = (os.path.join(HERE, "_ref", "train"),) if staged_only else ("/root/reference/train", os.path.join(HERE, "_ref", "train"))
    for d in cands:
        if os.path.isfile(os.path.join(d, "model.py")):
            return d
    return None


def _stubs():
    for n in ("matplotlib", "matplotlib.pyplot", "seaborn", "albumentations", "albumentations.pytorch"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["albumentations.pytorch"].ToTensorV2 = object
    if "wandb" not in sys.modules:
        try:
            importlib.import_module("wandb")
        except Exception:  # noqa: BLE001  (optional logger of train/train.py:219-224)
            sys.modules["wandb"] = types.ModuleType("wandb")


def _load(d, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    sys.path.insert(0, d)  # siblings the module imports by bare name and that were not loaded explicitly (e.g. `dataset`)
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(d)
    return mod


def load_reference(names=("config", "model", "utils"), shim=None, staged_only: bool = False):
    """Returns {name: module}.  ``shim``: {bare name: replacement module} installed before the remaining names are imported.
    Previously registered bare-name modules are dropped first, so a shimmed and an unshimmed load do not see each other."""
    d = ref_dir(staged_only)
    if d is None:
        raise RuntimeError("the reference is neither at /root/reference nor staged in oracle/_ref (python oracle/make_ref.py)")
    _stubs()
    for n in _BARE:
        sys.modules.pop(n, None)
    out = {}
    for n, mod in (shim or {}).items():
        sys.modules[n] = mod
        out[n] = mod
    for n in names:
        if n in out:
            continue
        mod = _load(d, n)
        if n == "model":
            mod.lraspp_mobilenet_v3_large = functools.partial(mod.lraspp_mobilenet_v3_large, weights_backbone=None)
        out[n] = mod
    return out


def unload():
    for n in _BARE:
        sys.modules.pop(n, None)
