#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: the UNMODIFIED reference (pure Python) made importable next to the oracle.
TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py          # needs /root/reference (the build container); no-op elsewhere

The reference has no setup.py / pyproject for `source/` and its requirements are not installable offline, so the
"build" is a verbatim tree copy of the packages the hot path lives in:

    /root/reference/source/              -> oracle/_ref/source/
    /root/reference/train_settings/      -> oracle/_ref/train_settings/
    /root/reference/third_party/{pytorch_ssim,ATE}/ -> oracle/_ref/third_party/...   (imported by training/base.py and
                                                         utils/geometry/align_trajectories.py)

`oracle/_ref/` is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so it travels to
the GPU box like the built `.so`.  Consumers (only `tests/`, `bench.py --impl reference` and bench.py's
`torch_gpu_baseline` leg) go through `oracle/ref_loader.py`, which adds the two import shims shipped under
`tests/golden/_shims` (easydict, lpips) and MagicMock stand-ins for plotting / PDC-Net / COLMAP leaf modules
(SURVEY.md appendix A).  Nothing under `sparf_b200/` ever imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
TREES = ["source", "train_settings", os.path.join("third_party", "pytorch_ssim"), os.path.join("third_party", "ATE")]


def build(force: bool = False) -> str:
    """Copy the reference packages into oracle/_ref/.  Returns the directory ('' if the reference is absent and no
    earlier copy exists)."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(os.path.join(DST, "source")) else ""
    stamp = os.path.join(DST, ".built_from")
    if not force and os.path.exists(stamp) and os.path.isdir(os.path.join(DST, "source")):
        return DST
    for t in TREES:
        dst = os.path.join(DST, t)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, t), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    tp_init = os.path.join(DST, "third_party", "__init__.py")
    if not os.path.exists(tp_init) and os.path.exists(os.path.join(SRC, "third_party", "__init__.py")):
        shutil.copy(os.path.join(SRC, "third_party", "__init__.py"), tp_init)
    with open(stamp, "w") as f:
        f.write(SRC + "\n")
    return DST


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out or "reference not present: nothing built")
