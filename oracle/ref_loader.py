"""Import the UNMODIFIED reference (from /root/reference in the build container, else from the git-ignored copy
`oracle/_ref/` made by `oracle/build_ref.py`).  TEST INFRASTRUCTURE ONLY: used by `tests/`, `tests/golden/*`,
`bench.py --impl reference` and bench.py's `torch_gpu_baseline` leg; never by `sparf_b200/`.

    root = ref_root()                      # '' when neither location exists
    mods = load(stack="renderer")          # source.models.renderer, frequency_nerf, camera, poses_models
    mods = load(stack="trainer")           # + source.training.joint_pose_nerf_trainer, loss_factory (leaf modules the
                                           #   image lacks are MagicMock stand-ins)
    mods = load(stack="trainer", shadow_renderer=True)
                                           # same, but `source.models.renderer` / `source.models.frequency_nerf` resolve
                                           # to sparf_b200's mirrors: the reference's trainer Graph subclass and loss
                                           # modules then run UNCHANGED on top of the CUDA path (the drop-in proof).
"""
from __future__ import annotations

import importlib
import os
import sys
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(os.path.dirname(HERE), "tests", "golden", "_shims")

_MOCKED = ["imageio", "matplotlib", "matplotlib.pyplot", "matplotlib.backends", "matplotlib.backends.backend_agg",
           "matplotlib.figure", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
           "coloredlogs", "source.models.flow_net", "source.utils.colmap_initialization.sfm",
           "source.utils.colmap_initialization.triangulation_w_known_poses",
           "third_party.DenseMatching.utils_flow.pixel_wise_mapping"]


def ref_root() -> str:
    if os.path.isdir("/root/reference/source"):
        return "/root/reference"
    cand = os.path.join(HERE, "_ref")
    return cand if os.path.isdir(os.path.join(cand, "source")) else ""


def _purge():
    for k in [k for k in sys.modules if k == "source" or k.startswith("source.") or k == "train_settings"
              or k.startswith("train_settings.") or k == "third_party" or k.startswith("third_party.")]:
        del sys.modules[k]


def load(stack: str = "renderer", shadow_renderer: bool = False) -> SimpleNamespace:
    root = ref_root()
    if not root:
        raise RuntimeError("reference not available: neither /root/reference nor oracle/_ref (run oracle/build_ref.py "
                           "in the build container)")
    for p in (root, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    _purge()   # a previous load() with a different `shadow_renderer` must not leak its class hierarchy
    out = SimpleNamespace(root=root)
    if stack == "trainer":
        from unittest.mock import MagicMock
        import torchvision  # noqa: F401  (must be imported before the mocks: torch._dynamo inspects sys.modules)
        for m in _MOCKED:
            if m not in sys.modules or not isinstance(sys.modules[m], MagicMock):
                sys.modules[m] = MagicMock()
    if shadow_renderer:
        import sparf_b200
        sparf_b200.install_as_reference_renderer()     # the product's own one-call swap
    out.renderer = importlib.import_module("source.models.renderer")
    out.frequency_nerf = importlib.import_module("source.models.frequency_nerf")
    out.camera = importlib.import_module("source.utils.camera")
    out.two_columns = importlib.import_module("source.models.poses_models.two_columns")
    if stack == "trainer":
        out.joint = importlib.import_module("source.training.joint_pose_nerf_trainer")
        out.loss_factory = importlib.import_module("source.training.core.loss_factory")
        out.base_losses = importlib.import_module("source.training.core.base_losses")
        out.sampling = importlib.import_module("source.training.core.sampling_strategies")
    return out
