"""Shared, reference-free builders for the golden cases.

Used by BOTH tests/golden/make_golden.py (which feeds them to the unmodified
reference imported from /root/reference, in the build container only) and by the
parity tests (which feed the same objects to the oracle and to the CUDA path on
the GPU box, where /root/reference does not exist).  Nothing here imports the
reference.

Everything is generated from numpy's PCG64 `default_rng(seed)` so that the inputs
are bit-reproducible across machines and torch versions.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))

try:  # pragma: no cover - depends on the image
    from easydict import EasyDict as edict
except ImportError:  # the image has no easydict: use the shim shipped with the goldens
    sys.path.insert(0, os.path.join(_HERE, "_shims"))
    from easydict import EasyDict as edict


# --------------------------------------------------------------------------------------
# options (the subset of train_settings/default_config.py:84-135,247-273 the renderer reads)
# --------------------------------------------------------------------------------------
def make_opt(*, S=128, S_fine=128, fine=False, depth_param="metric", depth_range=(1.2, 5.2),
             rand_rays=1024, stratified=False, noise=False, barf_c2f=None, setbg=False,
             max_iter=1000, ratio_start_fine=None, width=256, depth_layers=8, L_3D=10, L_view=4):
    opt = edict()
    opt.arch = edict()
    opt.arch.layers_feat = [None] + [width] * depth_layers
    opt.arch.layers_feat_fine = None
    opt.arch.layers_rgb = [None, width // 2, 3]
    opt.arch.skip = [4]
    opt.arch.posenc = edict(include_pi_in_posenc=True, add_raw_3D_points=True, add_raw_rays=True,
                            log_sampling=True, L_3D=L_3D, L_view=L_view)
    opt.arch.density_activ = "softplus"
    opt.arch.tf_init = True
    opt.nerf = edict()
    opt.nerf.view_dep = True
    opt.nerf.depth = edict(param=depth_param, range=list(depth_range))
    opt.nerf.sample_intvs = S
    opt.nerf.sample_stratified = stratified
    opt.nerf.fine_sampling = fine
    opt.nerf.sample_intvs_fine = S_fine
    opt.nerf.rand_rays = rand_rays
    opt.nerf.density_noise_reg = noise
    opt.nerf.setbg_opaque = setbg
    if ratio_start_fine is not None:
        opt.nerf.ratio_start_fine_sampling_at_x = ratio_start_fine
    opt.camera = edict(model="perspective", ndc=False, pose_parametrization="two_columns",
                       optimize_c2w=False, optimize_trans=True, optimize_rot=True,
                       optimize_relative_poses=False, n_first_fixed_poses=0)
    opt.barf_c2f = list(barf_c2f) if barf_c2f is not None else None
    opt.apply_cf_pe = True
    opt.mask_img = False
    opt.max_iter = max_iter
    opt.huber_loss_for_photometric = True
    opt.start_iter = edict(photometric=0, corres=0, depth_cons=0)
    opt.loss_weight = edict(parametrization="exp", equalize_losses=False, render=0, fg_mask=None,
                            distortion=None, depth_patch=None, corres=None, depth_cons=None,
                            render_matches=None)
    opt.depth_regu_patch_size = 2
    # loss-layer options (train_settings/default_config.py:130-202)
    opt.precrop_frac, opt.precrop_iters, opt.sampled_fraction_in_center = 0.5, 0, 0.0
    opt.start_ratio = edict(photometric=None, corres=None, depth_cons=None)
    opt.gradually_decrease_corres_weight = False
    opt.ratio_start_decrease_corres_weight = None
    opt.iter_start_decrease_corres_weight = 0
    opt.corres_weight_reduct_at_x_iter = 10000
    opt.gradually_decrease_depth_cons_loss = False
    opt.depth_cons_loss_reduct_at_x_iter = 10000
    opt.use_homography_flow = False
    opt.min_nbr_matches = 500
    opt.matching_pair_generation = "all_to_all"
    opt.loss_type = "photometric"
    return opt


# --------------------------------------------------------------------------------------
# deterministic MLP weights with the reference's state_dict keys and shapes
# (frequency_nerf.py:87-134: mlp_feat.{0..7}, mlp_rgb.{0,1}, progress)
# --------------------------------------------------------------------------------------
def layer_shapes(opt):
    in3 = 3 + 6 * opt.arch.posenc.L_3D
    inv = 3 + 6 * opt.arch.posenc.L_view
    feat = opt.arch.layers_feat
    shapes = []
    n = len(feat) - 1
    for li in range(n):
        k_in = in3 if li == 0 else feat[li]
        if li in opt.arch.skip:
            k_in += in3
        k_out = feat[li + 1] + (1 if li == n - 1 else 0)
        shapes.append(("mlp_feat.%d" % li, k_out, k_in))
    rgb = opt.arch.layers_rgb
    for li in range(len(rgb) - 1):
        k_in = feat[-1] + inv if li == 0 else rgb[li]
        shapes.append(("mlp_rgb.%d" % li, rgb[li + 1], k_in))
    return shapes


def det_weights(opt, seed, *, peaky=False, progress=None, sigma_bias=-6.0):
    """Xavier-uniform-like weights (ReLU gain) and NON-zero biases from PCG64(seed).

    peaky=True scales the density row and colour head so that densities/colours vary
    strongly along a ray (precision errors only become visible then)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, k_out, k_in in layer_shapes(opt):
        bound = math.sqrt(2.0) * math.sqrt(6.0 / (k_in + k_out))
        w = rng.uniform(-bound, bound, size=(k_out, k_in)).astype(np.float32)
        b = rng.uniform(-0.1, 0.1, size=(k_out,)).astype(np.float32)
        if peaky and name == "mlp_feat.%d" % (len(opt.arch.layers_feat) - 2):
            w[0] *= 3.0
            b[0] = sigma_bias
        if peaky and name.startswith("mlp_rgb.1"):
            w *= 2.0
        sd[name + ".weight"] = torch.from_numpy(w)
        sd[name + ".bias"] = torch.from_numpy(b)
    if progress is None:
        progress = 1.0 if opt.barf_c2f is None else 0.0
    sd["progress"] = torch.tensor(float(progress))
    return sd


# --------------------------------------------------------------------------------------
# synthetic scene with the loaders' shapes
# --------------------------------------------------------------------------------------
def look_at_w2c(cam_pos, target=(0.0, 0.0, 0.0), up=(0.0, -1.0, 0.0)):
    """OpenCV-convention w2c [3,4]: camera at cam_pos looking at target (+z forward)."""
    c = np.asarray(cam_pos, np.float64)
    z = np.asarray(target, np.float64) - c
    z /= np.linalg.norm(z)
    x = np.cross(z, np.asarray(up, np.float64))
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], 0)  # rows = camera axes in world coords  -> w2c rotation
    t = -R @ c
    return np.concatenate([R, t[:, None]], 1).astype(np.float32)


def make_scene(seed, B, H, W, *, radius=3.0, focal=None, identity=False):
    rng = np.random.default_rng(seed + 1000)
    focal = float(focal if focal is not None else max(H, W))
    poses = []
    for b in range(B):
        if identity:
            poses.append(np.concatenate([np.eye(3), np.zeros((3, 1))], 1).astype(np.float32))
        else:
            ang = 0.5 * (b - (B - 1) / 2.0)
            pos = [radius * math.sin(ang), 0.3 * (b % 2) - 0.1, -radius * math.cos(ang)]
            poses.append(look_at_w2c(pos))
    intr = np.array([[focal, 0, W / 2.0], [0, focal, H / 2.0], [0, 0, 1]], np.float32)
    data = edict()
    data.idx = torch.arange(B)
    data.image = torch.from_numpy(rng.uniform(0, 1, size=(B, 3, H, W)).astype(np.float32))
    data.intr = torch.from_numpy(np.stack([intr] * B))
    data.pose = torch.from_numpy(np.stack(poses))
    return data


def perturb_poses(pose_w2c, seed, sigma=0.05):
    """Small deterministic rigid perturbation of [B,3,4] w2c poses (noisy-GT initialisation)."""
    rng = np.random.default_rng(seed + 2000)
    out = []
    for P in pose_w2c.numpy():
        w = rng.normal(0, sigma, 3)
        th = np.linalg.norm(w)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        Rn = np.eye(3) + math.sin(th) / th * K + (1 - math.cos(th)) / th**2 * K @ K
        tn = rng.normal(0, sigma, 3)
        R = Rn @ P[:, :3].astype(np.float64)
        t = Rn @ P[:, 3].astype(np.float64) + tn
        out.append(np.concatenate([R, t[:, None]], 1).astype(np.float32))
    return torch.from_numpy(np.stack(out))


def subsample(x, n=257):
    """Deterministic strided subsample of a flattened array (keeps goldens small)."""
    flat = np.asarray(x).reshape(-1)
    if flat.size <= 4096:
        return flat.copy()
    step = max(1, flat.size // n)
    return flat[::step].copy()


class FakeFlowNet:
    """Stand-in for source/models/flow_net.FlowSelectionWrapper (PDC-Net needs a checkpoint we do not have):
    deterministic synthetic correspondence + confidence maps with the duck type the correspondence loss uses
    (SURVEY.md §8b): `.combi_list`, `.compute_flow_and_confidence_map_of_combi_list`, `.visualize_mapping_combinations`."""

    def __init__(self, n_views, H, W):
        self.H, self.W = H, W
        pairs = [[i, j] for i in range(n_views) for j in range(n_views) if i != j]
        self.combi_list = torch.tensor(pairs).T            # 2 x N (target, source)

    def compute_flow_and_confidence_map_of_combi_list(self, images, combi_list_tar_src, plot=False, use_homography=False):
        H, W = self.H, self.W
        N = combi_list_tar_src.shape[1]
        yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
        maps, confs = [], []
        for k in range(N):
            dx, dy = 1.5 + 0.25 * k, -0.75 + 0.2 * k
            maps.append(torch.stack([xx + dx + 0.02 * yy, yy + dy - 0.01 * xx], 0))
            conf = torch.full((1, H, W), 0.99)
            conf[:, :2] = 0.5
            conf[:, :, -3:] = 0.6
            confs.append(conf - 0.0005 * k)
        dev = images.device
        return torch.stack(maps).to(dev), torch.stack(confs).to(dev), None

    def visualize_mapping_combinations(self, **kw):
        return np.zeros((self.H, self.W, 3), np.uint8)


LOSS_CASES = {
    # BASELINE configs 3/4 flavour: joint pose + NeRF, photometric + correspondence + depth-consistency,
    # hierarchical sampling, BARF c2f mid-schedule, stratified jitter
    "c7_sparf_losses": dict(seed=7, B=3, H=24, W=32, n_rays=32, S=32, S_fine=32, fine=True, barf_c2f=(0.4, 0.7),
                            progress=0.5, depth_range=(1.2, 5.2), peaky=True, sigma_bias=-2.0, stratified=True,
                            rand_rays=96, min_nbr_matches=50, iteration=10),
    # BASELINE config 4 flavour (LLFF): inverse depth, coarse network only, no c2f mask left (progress = 1), other seed
    "c7b_sparf_losses_inverse": dict(seed=17, B=3, H=24, W=32, n_rays=32, S=48, S_fine=32, fine=False, barf_c2f=(0.4, 0.7),
                                     progress=1.0, depth_param="inverse", depth_range=(1, 0), data_depth_range=(0.5, 8.0),
                                     peaky=True, sigma_bias=-2.0, stratified=True, rand_rays=96, min_nbr_matches=50,
                                     iteration=10),
    # 4 views, metric depth, hierarchical, stage 1 of the joint schedule: the fine network is still off
    # (ratio_start_fine_sampling_at_x), so the losses see coarse depth only
    "c7c_sparf_losses_stage1": dict(seed=27, B=4, H=24, W=32, n_rays=24, S=32, S_fine=32, fine=True, barf_c2f=(0.1, 0.5),
                                    progress=0.2, depth_range=(1.0, 4.5), peaky=True, sigma_bias=-1.0, stratified=True,
                                    rand_rays=96, min_nbr_matches=50, iteration=10, ratio_start_fine=0.3),
    # DS-NeRF sparse-depth loss (base_losses.py:326-402) next to the photometric term: per-image render at the pixels
    # that have a COLMAP depth, more valid pixels than rand_rays // B in image 0 (randperm subsampling), none in image 2
    "c9_colmap_depth": dict(seed=37, B=3, H=24, W=32, n_rays=24, S=32, S_fine=32, fine=True, barf_c2f=None,
                            progress=None, depth_range=(1.2, 5.2), peaky=True, sigma_bias=-2.0, stratified=True,
                            rand_rays=96, min_nbr_matches=50, iteration=10, colmap=True,
                            loss_type="photometric_and_SparseCOLMAPDepthLoss"),
}


def loss_case_inputs(name):
    c = dict(LOSS_CASES[name])
    opt = make_opt(S=c["S"], S_fine=c["S_fine"], fine=c["fine"], depth_range=c["depth_range"], stratified=c["stratified"],
                   barf_c2f=c["barf_c2f"], rand_rays=c["rand_rays"], depth_param=c.get("depth_param", "metric"),
                   ratio_start_fine=c.get("ratio_start_fine"))
    opt.loss_type = c.get("loss_type", "photometric_and_corres_and_depth_cons")
    opt.loss_weight.corres = -3.0
    opt.loss_weight.depth_cons = -3.0
    opt.loss_weight.colmap_depth = 0
    opt.min_nbr_matches = c["min_nbr_matches"]
    data = make_scene(c["seed"], c["B"], c["H"], c["W"])
    data.depth_range = torch.tensor([list(map(float, c.get("data_depth_range", c["depth_range"])))] * c["B"])
    if c.get("colmap"):     # sparse "COLMAP" depth + confidence maps: [B,1,H,W], zero = no triangulated point
        rng_c = np.random.default_rng(c["seed"] + 4000)
        dm = rng_c.uniform(c["depth_range"][0] + 0.5, c["depth_range"][1] - 0.5, size=(c["B"], 1, c["H"], c["W"]))
        keep = rng_c.uniform(size=dm.shape) < np.array([0.2, 0.02, 0.0]).reshape(-1, 1, 1, 1)[: c["B"]]
        data.colmap_depth = torch.from_numpy((dm * keep).astype(np.float32))
        data.colmap_conf = torch.from_numpy(rng_c.uniform(0.2, 1.0, size=dm.shape).astype(np.float32))
    rng = np.random.default_rng(c["seed"] + 3000)
    ray_idx = torch.from_numpy(rng.permutation(c["H"] * c["W"])[: c["n_rays"]].astype(np.int64))
    sd = det_weights(opt, c["seed"], peaky=c["peaky"], progress=c["progress"], sigma_bias=c["sigma_bias"])
    sd_fine = det_weights(opt, c["seed"] + 77, peaky=c["peaky"], progress=c["progress"], sigma_bias=c["sigma_bias"])
    init_w2c = perturb_poses(data.pose, c["seed"], sigma=0.02)
    return c, opt, data, ray_idx, sd, sd_fine, init_w2c


# --------------------------------------------------------------------------------------
# golden case table: name -> kwargs.  Sizes are small (CPU reference finishes in seconds).
# --------------------------------------------------------------------------------------
CASES = {
    # BASELINE config 1: single 32x32 view, identity pose, 64 coarse samples, fixed poses
    "c1_coarse": dict(seed=1, B=1, H=32, W=32, identity=True, n_rays=96, S=64, fine=False,
                      depth_range=(0.5, 2.5), peaky=False, mode="train"),
    # BASELINE config 2 (shrunk image): 3 views, hierarchical 128 + 128, GT poses, peaky densities
    "c2_hier": dict(seed=2, B=3, H=30, W=40, n_rays=24, S=128, S_fine=128, fine=True,
                    depth_range=(1.2, 5.2), peaky=True, mode="train"),
    # BASELINE config 3: joint pose-NeRF, BARF c2f mid-schedule, stratified jitter + sigma noise
    "c3_barf_pose": dict(seed=3, B=3, H=30, W=40, n_rays=24, S=128, fine=False, barf_c2f=(0.4, 0.7),
                         progress=0.55, depth_range=(1.2, 5.2), peaky=True, mode="train",
                         pose_net=True, stratified=True, noise=True, sigma_bias=-2.0),
    # BASELINE config 4 flavour: inverse depth, float pixel locations (correspondence-loss entry)
    "c4_inverse_pixels": dict(seed=4, B=2, H=36, W=48, n_rays=20, S=128, fine=False,
                              depth_param="inverse", depth_range=(1, 0), peaky=True, mode="train",
                              pixels=True, pose_net=True, barf_c2f=(0.4, 0.7), progress=1.0),
    # eval mode + background compositing + hierarchical with pose grads (BARF runs fine net + poses)
    # default-off regularisers switched on: hierarchical, distortion (both nets) + depth-patch terms
    "c8_hier_regularisers": dict(seed=8, B=2, H=24, W=32, n_rays=24, S=64, S_fine=64, fine=True,
                                 depth_range=(1.0, 4.5), peaky=True, mode="train", regularisers=True),
    "c5_hier_pose_bg": dict(seed=5, B=2, H=24, W=32, n_rays=16, S=64, S_fine=64, fine=True,
                            depth_range=(0.8, 4.0), peaky=True, mode="train", pose_net=True,
                            setbg=True, barf_c2f=(0.1, 0.5), progress=0.3, sigma_bias=-0.5),
    # render_to_max (depth-consistency visibility pass): per-ray far bound, both nets on same samples
    "c6_to_max": dict(seed=6, B=2, H=24, W=32, n_rays=16, S=64, S_fine=64, fine=True,
                      depth_range=(0.8, 4.0), peaky=True, mode="train", to_max=True, sigma_bias=-3.0),
    # full-image inference through Graph.forward -> render_by_slices (renderer.py:347-381) in VAL mode: no stratified
    # jitter, no sigma noise although both flags are on, deterministic fine grid (renderer.py:326, 404, 435)
    "c10_val_full_image": dict(seed=10, B=2, H=12, W=16, n_rays=0, S=64, S_fine=64, fine=True, depth_range=(1.0, 4.5),
                               peaky=True, mode="val", full_image=True, rand_rays=80, stratified=True, noise=True,
                               sigma_bias=-2.0),
    # same in EVAL mode, coarse only, inverse depth, opaque background, c2f mask mid-schedule
    "c11_eval_full_image": dict(seed=11, B=1, H=12, W=16, n_rays=0, S=64, fine=False, depth_param="inverse",
                                depth_range=(1, 0), peaky=True, mode="eval", full_image=True, rand_rays=64,
                                stratified=True, setbg=True, barf_c2f=(0.1, 0.5), progress=0.35, sigma_bias=-2.0),
    # test-time pose optimisation (joint_pose_nerf_trainer.py:381-404): mode "test-optim", random rays from
    # Graph.forward, stratified + random fine grid, pose = se3 refinement o GT pose, gradient w.r.t. the 6-vector only
    "c12_test_optim": dict(seed=12, B=1, H=24, W=32, n_rays=0, S=64, S_fine=64, fine=True, depth_range=(1.0, 4.5),
                           peaky=True, mode="test-optim", test_optim=True, rand_rays=48, stratified=True,
                           sigma_bias=-2.0),
    # BASELINE config 5 flavour: 9 views, per-image (B,n) ray indices (RaySamplingStrategy, sampling_strategies.py:132),
    # hierarchical, photometric loss gathers per image (base_losses.py:283-291)
    "c13_b9_per_image_idx": dict(seed=13, B=9, H=20, W=24, n_rays=12, S=64, S_fine=64, fine=True,
                                 depth_range=(0.1, 4.5), peaky=True, mode="train", per_image_idx=True,
                                 sigma_bias=-2.0),
}


def case_inputs(name):
    """Everything a test needs to replay a case; shared with the tests (no reference here)."""
    c = dict(CASES[name])
    opt = make_opt(S=c["S"], S_fine=c.get("S_fine", 128), fine=c["fine"],
                          depth_param=c.get("depth_param", "metric"), depth_range=c["depth_range"],
                          stratified=c.get("stratified", False), noise=c.get("noise", False),
                          barf_c2f=c.get("barf_c2f"), setbg=c.get("setbg", False), rand_rays=c.get("rand_rays", 1024))
    data = make_scene(c["seed"], c["B"], c["H"], c["W"], identity=c.get("identity", False))
    data.depth_range = torch.tensor([list(map(float, c["depth_range"]))] * c["B"])
    rng = np.random.default_rng(c["seed"] + 3000)
    HW = c["H"] * c["W"]
    ray_idx = torch.from_numpy(rng.permutation(HW)[: c["n_rays"]].astype(np.int64))
    if c.get("per_image_idx"):
        ray_idx = torch.from_numpy(np.stack([rng.permutation(HW)[: c["n_rays"]] for _ in range(c["B"])]).astype(np.int64))
    if c.get("depth_param") == "inverse" and c.get("full_image"):   # the dataset still carries a metric range; the renderer ignores it
        data.depth_range = torch.tensor([[0.5, 8.0]] * c["B"])
    pixels = None
    if c.get("pixels"):
        px = rng.uniform([1, 1], [c["W"] - 2, c["H"] - 2], size=(c["B"], c["n_rays"], 2))
        pixels = torch.from_numpy(px.astype(np.float32))
    sb = c.get("sigma_bias", -6.0)
    sd = det_weights(opt, c["seed"], peaky=c["peaky"], progress=c.get("progress"), sigma_bias=sb)
    sd_fine = det_weights(opt, c["seed"] + 77, peaky=c["peaky"], progress=c.get("progress"), sigma_bias=sb) \
        if c["fine"] else None
    if c.get("regularisers"):
        opt.loss_weight.distortion = 0
        opt.loss_weight.depth_patch = 0
    init_w2c = perturb_poses(data.pose, c["seed"]) if c.get("pose_net") else None
    depth_max = None
    if c.get("to_max"):
        dm = rng.uniform(c["depth_range"][0] + 0.3, c["depth_range"][1], size=(c["B"], c["n_rays"]))
        depth_max = torch.from_numpy(dm.astype(np.float32))
    if c.get("test_optim"):     # non-zero se3 refinement so that the Lie exponential and its gradient are exercised
        c["se3_refine"] = torch.from_numpy(rng.normal(0, 0.02, size=(1, 6)).astype(np.float32))
    return c, opt, data, ray_idx, pixels, sd, sd_fine, init_w2c, depth_max
