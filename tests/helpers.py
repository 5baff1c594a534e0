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
    emb = None
    if init_w2c is not None:
        emb = O.pose_to_d9(init_w2c).to(dtype).clone().requires_grad_(True)
        pose = O.d9_to_pose(emb)
    else:
        pose = data.pose.to(dtype)
    intr = data.intr.to(dtype)
    if pixels is not None:
        center, ray = O.rays_at_pixels(pose, intr, pixels.to(dtype))
    else:
        center, ray = O.rays_from_ray_idx(pose, intr, c["H"], c["W"], ray_idx)
    rand, noise, noise_f, grid = case_randoms(c, gold)
    drange = opt.nerf.depth.range if opt.nerf.depth.param == "inverse" else data.depth_range[0]
    out = O.render(opt, params, params_f, center, ray, drange, mode=c["mode"], rand=rand,
                   noise=noise, noise_fine=noise_f, grid_fine=grid, iteration=10,
                   depth_max=depth_max.to(dtype) if depth_max is not None else None)
    if pixels is None and not c.get("to_max"):
        loss = O.photometric_loss(out, data.image.to(dtype), ray_idx)
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
    return out, loss, grads, gold


def rel_err(a, b):
    """max|a-b| / max|b| (the parity metric of BASELINE.md §2)."""
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach().cpu() if torch.is_tensor(b) else b), dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check_grads(grads, gold, tol, report=None):
    """Compare a grads dict against the golden's bias grads / sub-sampled weight grads."""
    worst = 0.0
    for k, g in grads.items():
        g = g.detach().cpu().double().numpy()
        if k in gold:
            ref = gold[k]
            e = np.abs(g - ref).max() / max(np.abs(ref).max(), 1e-30)
        elif k + ".sub" in gold:
            ref = gold[k + ".sub"]
            e = np.abs(common.subsample(g) - ref).max() / max(np.abs(ref).max(), 1e-30)
            ss = np.sqrt((g ** 2).sum())
            e = max(e, abs(ss - np.sqrt(gold[k + ".sumsq"])) / max(np.sqrt(gold[k + ".sumsq"]), 1e-30))
        else:
            raise KeyError(k)
        if report is not None:
            report[k] = e
        worst = max(worst, e)
        assert e < tol, (k, e)
    return worst
