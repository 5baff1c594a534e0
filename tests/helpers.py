"""Helpers shared by the parity tests: replay a golden case through the oracle."""
import os

import numpy as np
import torch

import common  # tests/golden/common.py (on sys.path via conftest)
from oracle import sparf_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def linear_probe_loss(out):
    """The fixed linear functional make_golden.py uses when there is no photometric target."""
    loss = 0
    for suf in ([""] + (["_fine"] if "rgb_fine" in out else [])):
        rgb, dep, opa = out["rgb" + suf], out["depth" + suf], out["opacity" + suf]
        wr = torch.linspace(0.5, 1.5, rgb.numel(), device=rgb.device).view_as(rgb).to(rgb.dtype)
        wd = torch.linspace(-0.2, 0.3, dep.numel(), device=rgb.device).view_as(dep).to(rgb.dtype)
        loss = loss + (rgb * wr).mean() + (dep * wd).mean() + 0.1 * (opa * wd).mean()
    return loss


def case_randoms(c, gold):
    """Map the recorded reference draws to (rand, noise, noise_fine, grid_fine)."""
    rand = torch.from_numpy(gold["rand_0"]) if "rand_0" in gold else None
    grid = torch.from_numpy(gold["rand_1"]) if "rand_1" in gold else None
    noise = torch.from_numpy(gold["randn_0"]) if "randn_0" in gold else None
    noise_f = torch.from_numpy(gold["randn_1"]) if "randn_1" in gold else None
    return rand, noise, noise_f, grid


def replay_oracle(name, dtype=torch.float32):
    """Run the oracle on a golden case.  Returns (out, loss, grads dict, extras)."""
    c, opt, data, ray_idx, pixels, sd, sd_fine, init_w2c, depth_max = common.case_inputs(name)
    gold = load_golden(name)
    cast = lambda x: x.to(dtype) if torch.is_floating_point(x) else x
    params = {k: cast(v).clone().requires_grad_(k != "progress") for k, v in sd.items()}
    params_f = {k: cast(v).clone().requires_grad_(k != "progress") for k, v in sd_fine.items()} if sd_fine else None
    emb = se3 = None
    if init_w2c is not None:
        emb = O.pose_to_d9(init_w2c).to(dtype).clone().requires_grad_(True)
        pose = O.d9_to_pose(emb)
    elif c.get("test_optim"):    # pose = GT pose o exp(se3 refinement)  (joint_pose_nerf_trainer.py:381-404, :737-739)
        se3 = c["se3_refine"].to(dtype).clone().requires_grad_(True)
        pose = O.compose_pair(O.se3_to_SE3(se3), data.pose.to(dtype))
        ray_idx = torch.from_numpy(gold["randperm_0"])[: opt.nerf.rand_rays // c["B"]]   # Graph.forward's draw
    else:
        pose = data.pose.to(dtype)
    if c.get("full_image"):
        ray_idx = None
    intr = data.intr.to(dtype)
    if pixels is not None:
        center, ray = O.rays_at_pixels(pose, intr, pixels.to(dtype))
    else:
        center, ray = O.rays_from_ray_idx(pose, intr, c["H"], c["W"], ray_idx)
    rand, noise, noise_f, grid = case_randoms(c, gold)
    drange = opt.nerf.depth.range if opt.nerf.depth.param == "inverse" else data.depth_range[0]
    if c.get("full_image"):     # val / eval: nothing random is drawn, slices of a full image = one big batch
        with torch.no_grad():
            out = O.render(opt, params, params_f, center, ray, drange, mode=c["mode"], iteration=10)
        return out, None, {}, gold
    out = O.render(opt, params, params_f, center, ray, drange, mode=c["mode"], rand=rand,
                   noise=noise, noise_fine=noise_f, grid_fine=grid, iteration=None if c.get("test_optim") else 10,
                   depth_max=depth_max.to(dtype) if depth_max is not None else None)
    if pixels is None and not c.get("to_max"):
        loss = O.photometric_loss(out, data.image.to(dtype), ray_idx)
        if c.get("regularisers"):
            reg = O.regularization_losses(out, True, True, opt.depth_regu_patch_size)
            loss = loss + reg["distortion"] + reg["depth_patch"]
            out["loss_distortion"], out["loss_depth_patch"] = reg["distortion"].detach(), reg["depth_patch"].detach()
    else:
        loss = linear_probe_loss(out)
    loss.backward()
    grads = {}
    for tag, p in (("nerf", params), ("nerf_fine", params_f)):
        if p is None:
            continue
        for k, v in p.items():
            if k != "progress":
                grads["grad_%s.%s" % (tag, k)] = v.grad
    if emb is not None:
        grads["grad_pose_embedding"] = emb.grad
    if se3 is not None:
        grads = {"grad_se3_refine": se3.grad}     # the only quantity test-time optimisation updates
    return out, loss, grads, gold


def rel_err(a, b):
    """max|a-b| / max|b| (the parity metric of BASELINE.md §2)."""
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach().cpu() if torch.is_tensor(b) else b), dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check_grads(grads, gold, tol, report=None, norm="max"):
    """Compare a grads dict against the golden's bias grads / sub-sampled weight grads.  norm="max": max|a-b| / max|b|
    per tensor; norm="l2": ||a-b|| / ||b|| (for the ill-conditioned inverse-depth cases, where single entries of two
    correct fp32 evaluations differ like phase noise and the max over a tensor is a heavy-tailed statistic)."""
    def dist(a, b):
        if norm == "l2":
            return np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30)
        return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    worst = 0.0
    for k, g in grads.items():
        g = g.detach().cpu().double().numpy()
        if k in gold:
            ref = gold[k]
            e = dist(g, ref)
        elif k + ".sub" in gold:
            ref = gold[k + ".sub"]
            e = dist(common.subsample(g), ref)
            ss = np.sqrt((g ** 2).sum())
            e = max(e, abs(ss - np.sqrt(gold[k + ".sumsq"])) / max(np.sqrt(gold[k + ".sumsq"]), 1e-30))
        else:
            raise KeyError(k)
        if report is not None:
            report[k] = e
        worst = max(worst, e)
        assert e < tol, (k, e)
    return worst


# ------------------------------------------------------------------------------------------------
# CUDA path replay (GPU box only)
# ------------------------------------------------------------------------------------------------
class RandomReplayer:
    """Feeds the reference's recorded draws to our Graph, which issues the same torch.rand /
    torch.randn_like calls in the same order (see tests/golden/make_golden.py: RandomRecorder)."""

    def __init__(self, gold):
        self.rand = [torch.from_numpy(gold[k]) for k in sorted(k for k in gold if k.startswith("rand_"))]
        self.randn = [torch.from_numpy(gold[k]) for k in sorted(k for k in gold if k.startswith("randn_"))]
        key = lambda k: int(k.split("_")[1])
        self.rand = [torch.from_numpy(gold[k]) for k in sorted((k for k in gold if k.startswith("rand_")), key=key)]
        self.randn = [torch.from_numpy(gold[k]) for k in sorted((k for k in gold if k.startswith("randn_")), key=key)]
        self.perm = [torch.from_numpy(gold[k]) for k in sorted((k for k in gold if k.startswith("randperm_")), key=key)]

    def _rand(self, *shape, device=None, **kw):
        x = self.rand.pop(0)
        return x.to(device) if device is not None else x

    def _randn_like(self, t, **kw):
        return self.randn.pop(0).to(t.device)

    def _randperm(self, n, device=None, **kw):
        x = self.perm.pop(0)
        assert x.numel() == int(n), (x.numel(), n)
        return x.to(device) if device is not None else x

    def __enter__(self):
        self._o = (torch.rand, torch.randn_like, torch.randperm)
        torch.rand, torch.randn_like, torch.randperm = self._rand, self._randn_like, self._randperm
        return self

    def __exit__(self, *a):
        torch.rand, torch.randn_like, torch.randperm = self._o


def build_graph(name, device="cuda"):
    """Our Graph (+ pose net) loaded with the golden case's deterministic weights."""
    from sparf_b200.renderer import Graph
    from sparf_b200.poses_models import FirstTwoColunmnsPoseParameters

    c, opt, data, ray_idx, pixels, sd, sd_fine, init_w2c, depth_max = common.case_inputs(name)
    dev = torch.device(device)
    for k in ("image", "intr", "pose", "depth_range", "idx"):
        data[k] = data[k].to(dev)
    if init_w2c is not None:
        pose_net = FirstTwoColunmnsPoseParameters(opt, nbr_poses=c["B"], initial_poses_w2c=init_w2c.to(dev), device=dev)

        class PoseGraph(Graph):
            def __init__(self, opt, device, pose_net):
                super().__init__(opt, device)
                self.pose_net = pose_net

            def get_w2c_pose(self, opt, data_dict, mode=None):
                return self.pose_net.get_w2c_poses()

        net = PoseGraph(opt, dev, pose_net.to(dev))
    elif c.get("test_optim"):
        from sparf_b200 import camera as our_camera

        class TestOptimGraph(Graph):   # the test-optim branch of joint_pose_nerf_trainer.Graph.get_w2c_pose
            def get_w2c_pose(self, opt, data_dict, mode=None):
                return our_camera.pose.compose([data_dict.pose_refine_test, data_dict.pose])

        net = TestOptimGraph(opt, dev)
    else:
        net = Graph(opt, dev)
    net.nerf.load_state_dict(sd)
    if c["fine"]:
        net.nerf_fine.load_state_dict(sd_fine)
    net.to(dev).train()
    mv = lambda x: x.to(dev) if x is not None else None
    return net, c, opt, data, mv(ray_idx), mv(pixels), mv(depth_max)


def replay_graph(name, engine="simt_fp32"):
    """Run the CUDA path on a golden case exactly as make_golden.py ran the reference."""
    import sparf_b200
    from oracle import sparf_oracle as O

    sparf_b200.set_engine(engine)
    gold = load_golden(name)
    net, c, opt, data, ray_idx, pixels, depth_max = build_graph(name)
    se3 = None
    if c.get("test_optim"):
        from sparf_b200 import camera as our_camera
        se3 = torch.nn.Parameter(c["se3_refine"].clone().cuda())
        data.pose_refine_test = our_camera.lie.se3_to_SE3(se3)
    with RandomReplayer(gold):
        if c.get("full_image"):
            with torch.no_grad():
                out = net.forward(opt, data, iter=10, mode=c["mode"])
            return out, None, {}, gold
        elif c.get("test_optim"):
            out = net.forward(opt, data, iter=None, mode=c["mode"])
            ray_idx = out.ray_idx
        elif c.get("to_max"):
            pose = net.get_w2c_pose(opt, data, mode=c["mode"])
            out = net.render_up_to_maxdepth_at_specific_pose_and_rays(
                opt, data, pose, data.intr, c["H"], c["W"], depth_max=depth_max, iter=10, ray_idx=ray_idx, mode=c["mode"])
        elif pixels is not None:
            out = net.render_image_at_specific_rays(opt, data, iter=10, pixels=pixels, mode=c["mode"])
        else:
            out = net.render_image_at_specific_rays(opt, data, iter=10, ray_idx=ray_idx, mode=c["mode"])
    if pixels is None and not c.get("to_max"):
        loss = O.photometric_loss(out, data.image, ray_idx)   # the loss formula is the checker's, the render is ours
        if c.get("regularisers"):   # ... except the regularisers, whose product path (kernel + Loss mirror) is under test
            from sparf_b200.losses import BasePhotoandReguLoss
            mod = BasePhotoandReguLoss(opt, net, train_data=None, device=data.image.device)
            reg = mod.compute_regularization_losses(opt, out, {})
            loss = loss + reg["distortion"] + reg["depth_patch"]
            out["loss_distortion"], out["loss_depth_patch"] = reg["distortion"].detach(), reg["depth_patch"].detach()
    else:
        loss = linear_probe_loss(out)
    loss.backward()
    grads = {}
    nets = [("nerf", net.nerf)] + ([("nerf_fine", net.nerf_fine)] if c["fine"] else [])
    for tag, m in nets:
        for pname, p in m.named_parameters():
            if pname != "progress":
                grads["grad_%s.%s" % (tag, pname)] = p.grad
    if hasattr(net, "pose_net"):
        grads["grad_pose_embedding"] = net.pose_net.pose_embedding.grad
    if se3 is not None:
        grads = {"grad_se3_refine": se3.grad}
    return out, loss, grads, gold
