"""TEST / BASELINE INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference into ``oracle/_ref/``.

``/root/reference`` exists only in the build container; the GPU box gets a snapshot of this repository.  ``oracle/_ref/`` is
git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so the byte-identical copies made here
travel to the GPU box like the built ``.so`` does.  ``__graft_entry__.build()`` runs this; nothing is edited or generated --
``shutil.copy2`` of the files listed below, plus a manifest with their sha256 so a test can prove they are unmodified.

Used by: ``bench.py --impl reference`` (the reference arm runs the reference's own ``create_model``), the ``torch_eager_gpu``
bench leg (the reference model through PyTorch/cuDNN on the same B200), ``tests/test_reference_drivers.py`` (the reference's
own ``train_epoch`` / ``validate_epoch`` / ``ModelEvaluator`` driving this package).  The product package never imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = [  # the path of SURVEY.md §8(a): model, loss/metrics/checkpoints, config, the two drivers (+ the dataset module they import)
    "train/model.py", "train/utils.py", "train/config.py", "train/train.py", "train/evaluate.py", "train/dataset.py",
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make_ref(verbose: bool = True) -> str | None:
    """Copy the reference files (when /root/reference is present); returns the staged directory or None."""
    if not os.path.isdir(SRC):
        return DST if os.path.exists(os.path.join(DST, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy2(src, dst)
        manifest[rel] = _sha(dst)
        assert manifest[rel] == _sha(src)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"staged {len(FILES)} unmodified reference files into {DST}")
    return DST


if __name__ == "__main__":
    make_ref()
