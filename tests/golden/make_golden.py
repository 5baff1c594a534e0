#!/usr/bin/env python
"""Generate the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py c2_hier    # one case

The reference is pure Python/PyTorch and imports on CPU with two shims
(tests/golden/_shims: easydict, lpips).  Random draws made inside the reference
(torch.rand in renderer.py:406 and :439, torch.randn_like in frequency_nerf.py:192) are
replaced by recorded PCG64 streams so that the CUDA path can be given the very same numbers.

Each .npz stores: the recorded random tensors, every renderer output the reference produced,
the reference photometric loss (source/training/core/base_losses.py:243) and its gradients
(bias grads in full, weight grads as a strided sub-sample + fp64 sum / sum of squares,
pose-embedding grads in full).  Inputs are NOT stored: tests rebuild them from common.py.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import numpy as np
import torch

import common  # noqa: E402

from source.models import renderer as ref_renderer  # noqa: E402
from source.models.poses_models.two_columns import FirstTwoColunmnsPoseParameters  # noqa: E402
from source.training.core.base_losses import BasePhotoandReguLoss  # noqa: E402


class RandomRecorder:
    """Replaces torch.rand / torch.randn_like by PCG64 streams and records what was handed out."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed + 5000)
        self.rand_calls = []
        self.randn_calls = []

    def rand(self, *shape, device=None, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        x = torch.from_numpy(self.rng.uniform(0, 1, size=shape).astype(np.float32))
        self.rand_calls.append(x.clone())
        return x

    def randn_like(self, t, **kw):
        x = torch.from_numpy(self.rng.normal(0, 1, size=tuple(t.shape)).astype(np.float32))
        self.randn_calls.append(x.clone())
        return x

    def randperm(self, n, device=None, **kw):
        x = torch.from_numpy(self.rng.permutation(int(n)).astype(np.int64))
        self.randperm_calls.append(x.clone())
        return x

    def tensor(self, *a, **kw):
        # The reference writes `torch.tensor(0., requires_grad=True).to(device)` and later `loss += ...`
        # (base_losses.py:366-393).  On CUDA -- the only device the reference runs on -- `.to` copies, so the
        # accumulator is a non-leaf; on this CPU-only box `.to` is the identity and the in-place add on a leaf raises.
        # Reproduce the CUDA behaviour: hand out a non-leaf zero (same value, same arithmetic).
        t = self._tensor(*a, **kw)
        return t * 1.0 if (kw.get("requires_grad") and t.dim() == 0) else t

    def __enter__(self):
        self.randperm_calls = []
        self._rand, self._randn_like, self._randperm, self._tensor = torch.rand, torch.randn_like, torch.randperm, torch.tensor
        torch.rand, torch.randn_like, torch.randperm, torch.tensor = self.rand, self.randn_like, self.randperm, self.tensor
        return self

    def __exit__(self, *a):
        torch.rand, torch.randn_like, torch.randperm, torch.tensor = self._rand, self._randn_like, self._randperm, self._tensor


class PoseGraph(ref_renderer.Graph):
    """Same override as source/training/joint_pose_nerf_trainer.py:710-722 (train mode)."""

    def __init__(self, opt, device, pose_net):
        super().__init__(opt, device)
        self.pose_net = pose_net

    def get_w2c_pose(self, opt, data_dict, mode=None):
        return self.pose_net.get_w2c_poses()


class TestOptimGraph(ref_renderer.Graph):
    """The test-time-optimisation branch of joint_pose_nerf_trainer.Graph.get_w2c_pose (:720-741) with the sim(3)
    alignment left out (identity): pose = pose_refine_test o GT pose."""

    def get_w2c_pose(self, opt, data_dict, mode=None):
        from source.utils import camera as ref_camera
        return ref_camera.pose.compose([data_dict.pose_refine_test, data_dict.pose])


def run_case(name):
    c, opt, data, ray_idx, pixels, sd, sd_fine, init_w2c, depth_max = common.case_inputs(name)
    dev = torch.device("cpu")
    torch.manual_seed(0)
    se3 = None
    if c.get("test_optim"):
        from source.utils import camera as ref_camera
        net = TestOptimGraph(opt, dev)
        se3 = torch.nn.Parameter(c["se3_refine"].clone())
        data.pose_refine_test = ref_camera.lie.se3_to_SE3(se3)
    elif c.get("pose_net"):
        pose_net = FirstTwoColunmnsPoseParameters(opt, nbr_poses=c["B"], initial_poses_w2c=init_w2c, device=dev)
        net = PoseGraph(opt, dev, pose_net)
    else:
        net = ref_renderer.Graph(opt, dev)
    net.nerf.load_state_dict(sd)
    if c["fine"]:
        net.nerf_fine.load_state_dict(sd_fine)
    net.train()

    out_npz = {}
    with RandomRecorder(c["seed"]) as rec:
        if c.get("full_image"):
            with torch.no_grad():   # val_step / evaluate_full: Graph.forward -> render_by_slices
                out = net.forward(opt, data, iter=10, mode=c["mode"])
        elif c.get("test_optim"):
            out = net.forward(opt, data, iter=None, mode=c["mode"])
        elif c.get("to_max"):
            pose = net.get_w2c_pose(opt, data, mode=c["mode"])
            out = net.render_up_to_maxdepth_at_specific_pose_and_rays(
                opt, data, pose, data.intr, c["H"], c["W"], depth_max=depth_max, iter=10,
                ray_idx=ray_idx, mode=c["mode"])
        elif pixels is not None:
            out = net.render_image_at_specific_rays(opt, data, iter=10, pixels=pixels, mode=c["mode"])
        else:
            out = net.render_image_at_specific_rays(opt, data, iter=10, ray_idx=ray_idx, mode=c["mode"])

    # loss: the reference photometric module when rays come from ray_idx; for float pixels the
    # reference has no photometric target, so use a fixed linear functional of rgb/depth/opacity
    # (exercises the same gradients; the formula is restated in the tests).
    if c.get("full_image"):
        for k, v in out.items():
            if isinstance(v, torch.Tensor) and k not in ("ray_idx", "idx_img_rendered"):
                out_npz["out_" + k] = v.detach().numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out_npz)
        print("%-20s (forward only) keys=%d  %.1f KB" % (name, len(out_npz), os.path.getsize(path) / 1024))
        return
    if pixels is None and not c.get("to_max"):
        loss_mod = BasePhotoandReguLoss(opt, net, train_data=None, device=dev)
        loss_dict, _, _ = loss_mod.compute_loss(opt, data, out, iteration=10, mode=c["mode"])
        loss = loss_dict.render
        if c.get("regularisers"):   # base_losses.py:162-194 through the reference's own module
            loss = loss + loss_dict.distortion + loss_dict.depth_patch
            extra = {"loss_render": loss_dict.render, "loss_distortion": loss_dict.distortion,
                     "loss_depth_patch": loss_dict.depth_patch}
    else:
        loss = 0
        for suf in ([""] + (["_fine"] if "rgb_fine" in out else [])):
            rgb, dep, opa = out["rgb" + suf], out["depth" + suf], out["opacity" + suf]
            wr = torch.linspace(0.5, 1.5, rgb.numel()).view_as(rgb)
            wd = torch.linspace(-0.2, 0.3, dep.numel()).view_as(dep)
            loss = loss + (rgb * wr).mean() + (dep * wd).mean() + 0.1 * (opa * wd).mean()
    loss.backward()

    for k, v in out.items():
        if isinstance(v, torch.Tensor) and k not in ("ray_idx", "idx_img_rendered"):
            out_npz["out_" + k] = v.detach().numpy()
    out_npz["loss"] = np.float64(loss.item())
    if c.get("regularisers"):
        for k, v in extra.items():
            out_npz[k] = np.float64(v.item())
    for i, r in enumerate(rec.rand_calls):
        out_npz["rand_%d" % i] = r.numpy()
    for i, r in enumerate(rec.randn_calls):
        out_npz["randn_%d" % i] = r.numpy()
    for i, r in enumerate(rec.randperm_calls):
        out_npz["randperm_%d" % i] = r.numpy()
    if se3 is not None:
        out_npz["grad_se3_refine"] = se3.grad.numpy()

    nets = [("nerf", net.nerf)] + ([("nerf_fine", net.nerf_fine)] if c["fine"] else [])
    for tag, m in nets:
        for pname, p in m.named_parameters():
            if pname == "progress":
                continue
            assert p.grad is not None, (tag, pname)
            g = p.grad.numpy()
            key = "grad_%s.%s" % (tag, pname)
            if pname.endswith("bias"):
                out_npz[key] = g
            else:
                out_npz[key + ".sub"] = common.subsample(g)
                out_npz[key + ".sum"] = np.float64(g.astype(np.float64).sum())
                out_npz[key + ".sumsq"] = np.float64((g.astype(np.float64) ** 2).sum())
    if c.get("pose_net"):
        out_npz["grad_pose_embedding"] = net.pose_net.pose_embedding.grad.numpy()

    # drop the big per-sample tensors nobody consumes, sub-sample the rest
    for k in list(out_npz):
        if k.startswith("out_rgb_samples") or k.startswith("out_density_samples"):
            out_npz[k] = out_npz[k].astype(np.float32)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out_npz)
    print("%-20s loss=%.8f  keys=%d  %.1f KB" % (name, loss.item(), len(out_npz), os.path.getsize(path) / 1024))


def run_loss_case(name):
    """Full SPARF step (photometric + correspondence + depth-consistency) through the reference's own loss
    modules (source/training/core/{base_losses,corres_loss,depth_cons_loss,loss_factory}.py)."""
    from unittest.mock import MagicMock
    import torchvision  # noqa: F401  (must be imported before the mocks)
    for m in ["imageio", "matplotlib", "matplotlib.pyplot", "matplotlib.backends", "matplotlib.backends.backend_agg",
              "matplotlib.figure", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
              "coloredlogs", "source.models.flow_net", "source.utils.colmap_initialization.sfm",
              "source.utils.colmap_initialization.triangulation_w_known_poses",
              "third_party.DenseMatching.utils_flow.pixel_wise_mapping"]:
        sys.modules.setdefault(m, MagicMock())
    from source.training.core.loss_factory import define_loss
    from easydict import EasyDict as edict

    c, opt, data, ray_idx, sd, sd_fine, init_w2c = common.loss_case_inputs(name)
    dev = torch.device("cpu")
    pose_net = FirstTwoColunmnsPoseParameters(opt, nbr_poses=c["B"], initial_poses_w2c=init_w2c, device=dev)
    net = PoseGraph(opt, dev, pose_net)
    net.nerf.load_state_dict(sd)
    if c["fine"]:
        net.nerf_fine.load_state_dict(sd_fine)
    net.train()
    train_data = edict(all=data)
    train_data.__class__.__len__ = lambda self: c["B"]
    flow = common.FakeFlowNet(c["B"], c["H"], c["W"])
    np.random.seed(c["seed"])
    with RandomRecorder(c["seed"]) as rec:
        loss_module = define_loss(opt.loss_type, opt, net, train_data, dev, flow_net=flow)
        data["iter"] = c["iteration"]
        out = net.render_image_at_specific_rays(opt, data, iter=c["iteration"], ray_idx=ray_idx, mode="train")
        data.poses_w2c = net.get_w2c_pose(opt, data, mode="train")
        loss_dict, stats, _ = loss_module.compute_loss(opt, data, out, iteration=c["iteration"], mode="train")
    loss_dict["all"].backward()
    out_npz = {"loss_" + k: np.float64(v.item()) for k, v in loss_dict.items() if torch.is_tensor(v) and v.dim() == 0}
    for i, r in enumerate(rec.rand_calls):
        out_npz["rand_%d" % i] = r.numpy()
    for i, r in enumerate(rec.randperm_calls):
        out_npz["randperm_%d" % i] = r.numpy()
    for tag, m in [("nerf", net.nerf)] + ([("nerf_fine", net.nerf_fine)] if c["fine"] else []):
        for pname, p in m.named_parameters():
            if pname == "progress":
                continue
            if p.grad is None:      # fine network still switched off by the schedule: no gradient
                continue
            g = p.grad.numpy()
            key = "grad_%s.%s" % (tag, pname)
            if pname.endswith("bias"):
                out_npz[key] = g
            else:
                out_npz[key + ".sub"] = common.subsample(g)
                out_npz[key + ".sum"] = np.float64(g.astype(np.float64).sum())
                out_npz[key + ".sumsq"] = np.float64((g.astype(np.float64) ** 2).sum())
    out_npz["grad_pose_embedding"] = net.pose_net.pose_embedding.grad.numpy()
    out_npz["out_rgb"] = out["rgb"].detach().numpy()
    if "depth_fine" in out:
        out_npz["out_depth_fine"] = out["depth_fine"].detach().numpy()
    out_npz["out_depth"] = out["depth"].detach().numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out_npz)
    print("%-20s %s  %.1f KB" % (name, {k: round(float(v), 6) for k, v in out_npz.items() if k.startswith("loss_")},
                                  os.path.getsize(path) / 1024))


if __name__ == "__main__":
    names = sys.argv[1:] or (list(common.CASES) + list(common.LOSS_CASES))
    for n in names:
        if n in common.LOSS_CASES:
            run_loss_case(n)
        else:
            run_case(n)
