"""Host-side mirror of the reference loss layer for the hot path (source/training/core/):

    Loss                              base_losses.py:26     aggregation, all = sum 10^w * loss
    BasePhotoandReguLoss              base_losses.py:243    photometric Huber / MSE (+ fg-mask)
    CorrespondencesPairRenderDepthAndGet3DPtsAndReproject   corres_loss.py:29 + base_corres_loss.py:28
    DepthConsistencyLoss              depth_cons_loss.py:32
    SparseCOLMAPDepthLoss             base_losses.py:326    DS-NeRF sparse-depth term
    define_loss                       loss_factory.py:25

Same constructor arguments, `compute_loss(opt, data_dict, output_dict, iteration, mode, plot)` ->
`(loss_dict, stats_dict, plotting_dict)`, same option names, same random draws (np.random.randint /
np.random.rand / torch.randperm in the same order).  Every render these losses trigger goes through
`Graph.render_image_at_specific_pose_and_rays` / `render_up_to_maxdepth_...`, i.e. the CUDA kernels;
what remains here is the tiny projective geometry on <= ~1e3 points per step (SURVEY.md §2 row 7:
"keep in torch first") and the photometric Huber reduction, which runs on the huber2 kernel.

The reference's own loss modules also work unchanged on top of `sparf_b200.renderer.Graph`
(INTEGRATION.md); this mirror exists so the path is usable and testable without the reference.
"""
from __future__ import annotations

from typing import Any, Dict, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .utils.edict import edict


# ------------------------------------------------------------------------------------------------
# small geometry (source/utils/geometry/batched_geometry_utils.py:42-49, 199-266; camera.py:37-64)
# ------------------------------------------------------------------------------------------------
def to_homogeneous(p: torch.Tensor) -> torch.Tensor:
    return torch.cat([p, torch.ones_like(p[..., :1])], dim=-1)


def from_homogeneous(p: torch.Tensor) -> torch.Tensor:
    return p[..., :-1] / (p[..., -1:] + 1e-6)


def _bottom_row(device, dtype=torch.float32) -> torch.Tensor:
    """[0, 0, 0, 1] built ON the device (a python-list constructor is a pageable host->device copy: a synchronisation,
    and illegal inside a CUDA-graph capture)."""
    return torch.cat((torch.zeros(3, device=device, dtype=dtype), torch.ones(1, device=device, dtype=dtype)))


def pose_inverse_4x4(mat: torch.Tensor) -> torch.Tensor:
    """[...,4,4] rigid inverse without a matrix inverse: [R^T | -R^T t]."""
    R, t = mat[..., :3, :3], mat[..., :3, 3:]
    Rt = R.transpose(-1, -2)
    top = torch.cat((Rt, -Rt @ t), dim=-1)
    bottom = _bottom_row(mat.device, mat.dtype).expand(*mat.shape[:-2], 1, 4)    # (no scalar setitem: that is a host copy)
    return torch.cat((top, bottom), dim=-2)


def batch_backproject_to_3d(kpi, di, Ki, T_itoj):
    """pixels [N,2] with depth [N] -> 3-D points in frame j [N,3]."""
    x = to_homogeneous(kpi) @ torch.linalg.inv_ex(Ki).inverse.transpose(-1, -2)   # (inv_ex: no host-side singularity check)
    x = x * di[..., None]
    return from_homogeneous(to_homogeneous(x) @ T_itoj.transpose(-1, -2))


def batch_project(x_i, T_itoj, Kj, return_depth=False):
    """3-D points in frame i -> pixels (and depth) in image j."""
    x_j = from_homogeneous(to_homogeneous(x_i) @ T_itoj.transpose(-1, -2))
    uv = from_homogeneous(x_j @ Kj.transpose(-1, -2))
    return (uv, x_j[..., -1]) if return_depth else uv


def batch_project_to_other_img(kpi, di, Ki, Kj, T_itoj, return_depth=False):
    if di.dim() == kpi.dim():
        di = di.squeeze(-1)
    x_j = batch_backproject_to_3d(kpi, di, Ki, T_itoj)
    uv = from_homogeneous(x_j @ Kj.transpose(-1, -2))
    return (uv, x_j[..., -1]) if return_depth else uv


def generate_pair_list(n_views: int) -> torch.Tensor:
    """Unordered exhaustive pairs as a 2 x N tensor (correspondence_utils.py:213-221)."""
    pairs = [[i, j] for i in range(n_views) for j in range(i + 1, n_views)]
    return torch.from_numpy(np.array(pairs).T)


def get_nearest_pose_ids(tar_pose_c2w: np.ndarray, ref_poses_c2w: np.ndarray, tar_id: int) -> int:
    """Closest other camera by the angle between camera-position vectors (data_utils.py:267-311,
    angular_dist_method='vector', scene centre at the origin)."""
    tiny = 1e-6   # TINY_NUMBER, data_utils.py:27
    a = tar_pose_c2w[:3, 3][None].repeat(len(ref_poses_c2w), 0)
    b = ref_poses_c2w[:, :3, 3]
    au = a / (np.linalg.norm(a, axis=1, keepdims=True) + tiny)
    bu = b / (np.linalg.norm(b, axis=1, keepdims=True) + tiny)
    d = np.arccos(np.clip((au * bu).sum(-1), -1.0, 1.0))
    if tar_id >= 0:
        d[tar_id] = 1e3
    return int(np.argsort(d)[0])


from .sampling_strategies import sample_rays  # noqa: E402,F401  (sampling_strategies.py:250-295)


def _with_defaults(defaults: Dict[str, Any], opt) -> edict:
    out = edict(defaults)
    for k, v in opt.items():
        out[k] = v
    return out


# ------------------------------------------------------------------------------------------------
# aggregation
# ------------------------------------------------------------------------------------------------
class Loss:
    """Runs every loss module and combines them (base_losses.py:26-135)."""

    # The reference asserts every term finite on the host each step (base_losses.py:118-124): one device
    # synchronisation per loss key.  `check_finite = False` (or a CUDA-graph capture in progress) skips the asserts so
    # that a step stays sync-free; the fused optimiser's non-finite guard (csrc/optim.cu) then covers the same failure.
    check_finite = True

    def __init__(self, loss_modules):
        self.loss_modules = loss_modules

    def compute_loss(self, opt, data_dict, output_dict, iteration, mode=None, plot=False, **kwargs):
        loss, stats, plots = edict(), {}, {}
        for m in self.loss_modules:
            l, s, p = m.compute_loss(opt, data_dict, output_dict, iteration=iteration, mode=mode, plot=plot, **kwargs)
            loss.update(l)
            stats.update(s)
            plots.update(p)
        if opt.loss_weight.equalize_losses:
            loss = self.summarize_loss_w_equal_weights(opt, loss)
        else:
            loss = self.summarize_loss_w_predefined_weights(opt, loss)
        return loss, stats, plots

    def compute_flow(self, train_data, opt):
        for m in self.loss_modules:
            if hasattr(m, "compute_flow"):
                m.compute_flow(train_data, opt)

    def plot_something(self):
        out = {}
        for m in self.loss_modules:
            if hasattr(m, "plot_something"):
                out.update(m.plot_something())
        return out

    def get_flow_metrics(self):
        out = {}
        for m in self.loss_modules:
            if hasattr(m, "get_flow_metrics"):
                out.update(m.get_flow_metrics())
        return out

    def _checked(self, opt, loss_dict):
        assert "all" not in loss_dict
        check = self.check_finite and not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing())
        for key in loss_dict:
            assert key in opt.loss_weight, key
            assert loss_dict[key].shape == ()
            if opt.loss_weight[key] is not None:
                if check:
                    assert not torch.isinf(loss_dict[key]), "loss {} is Inf".format(key)
                    assert not torch.isnan(loss_dict[key]), "loss {} is NaN".format(key)
                yield key

    def summarize_loss_w_equal_weights(self, opt, loss_dict):
        total, extra = 0.0, {}
        assert "render" in loss_dict
        ref = loss_dict.render.detach()
        for key in self._checked(opt, loss_dict):
            w = ref / (loss_dict[key].detach() + 1e-6) if loss_dict[key] != 0.0 else 1.0
            extra[key + "_after_w"] = w * loss_dict[key]
            total = total + extra[key + "_after_w"]
        loss_dict.update(all=total)
        loss_dict.update(extra)
        return loss_dict

    def summarize_loss_w_predefined_weights(self, opt, loss_dict):
        total, extra = None, {}
        for key in self._checked(opt, loss_dict):
            w = 10 ** float(opt.loss_weight[key]) if opt.loss_weight.parametrization == "exp" else float(opt.loss_weight[key])
            # (w == 1, the photometric term: the same value without a multiply kernel; first term: no `0.0 + x` kernel)
            extra[key + "_after_w"] = loss_dict[key] if w == 1.0 else w * loss_dict[key]
            total = extra[key + "_after_w"] if total is None else total + extra[key + "_after_w"]
        if total is None:
            total = 0.0
        loss_dict.update(all=total)
        loss_dict.update(extra)
        return loss_dict


class BaseLoss:
    def __init__(self, device):
        self.device = device

    def L1_loss(self, pred, label):
        return (pred.contiguous() - label).abs().mean()

    def MSE_loss(self, pred, label):
        loss = (pred.contiguous() - label) ** 2
        return loss.sum() / (loss.nelement() + 1e-6)

    def huber_loss(self, pred, label, reduction="mean"):
        """2 * Huber(delta=0.5) (base_losses.py:155-156); the mean reduction runs on the CUDA kernel."""
        if reduction == "mean" and pred.is_cuda:
            return ops.huber2(pred, label)
        return nn.functional.huber_loss(pred, label, reduction=reduction, delta=0.5) * 2.0

    def compute_diff_loss(self, loss_type, diff, weights=None, var=None, mask=None, dim=-1):
        """base_losses.py:197-224."""
        lt = loss_type.lower()
        if lt == "epe":
            loss = torch.norm(diff, 2, dim, keepdim=True)
        elif lt == "l1":
            loss = torch.abs(diff)
        elif lt == "mse":
            loss = diff ** 2
        elif lt == "huber":
            loss = nn.functional.huber_loss(diff, torch.zeros_like(diff), reduction="none", delta=1.0)
        else:
            raise ValueError("Wrong loss type: {}".format(loss_type))
        if weights is not None:
            assert weights.dim() == loss.dim()
            loss = loss * weights
        if var is not None:
            eps = torch.tensor(1e-3)
            loss = loss / torch.maximum(var, eps) + torch.log(torch.maximum(var, eps))
        if mask is not None:
            assert mask.dim() == loss.dim()
            loss = loss * mask.float()
            return loss.sum() / (mask.float().sum() + 1e-6)
        return loss.sum() / (loss.nelement() + 1e-6)


# ------------------------------------------------------------------------------------------------
# photometric
# ------------------------------------------------------------------------------------------------
class BasePhotoandReguLoss(BaseLoss):
    def __init__(self, opt, nerf_net, train_data, device):
        super().__init__(device)
        self.opt, self.net, self.train_data = opt, nerf_net, train_data

    def compute_loss(self, opt, data_dict, output_dict, iteration, mode=None, plot=False, **kwargs):
        if iteration < self.opt.start_iter.photometric:
            return edict(render=torch.zeros((), device=self.device, requires_grad=True)), {}, {}
        loss_dict = edict()
        B = len(data_dict.idx)
        image = data_dict.image.reshape(B, 3, -1).permute(0, 2, 1)                       # [B,HW,3]
        fg_mask = data_dict.fg_mask.float().view(B, -1, 1) if opt.loss_weight.fg_mask is not None else None
        n_img = len(output_dict.idx_img_rendered)
        assert n_img == B
        if (hasattr(opt, "nerf") and opt.nerf.rand_rays) and mode in ["train", "test-optim"]:
            ridx = output_dict.ray_idx
            if ridx.dim() == 2 and ridx.shape[0] == n_img:
                gi = ridx.long()[..., None]
                image = torch.gather(image, 1, gi.expand(-1, -1, 3))
                if fg_mask is not None:
                    fg_mask = torch.gather(fg_mask, 1, gi)
            else:
                image = image[:, ridx]
                if fg_mask is not None:
                    fg_mask = fg_mask[:, ridx]
        crit = self.huber_loss if self.opt.huber_loss_for_photometric else self.MSE_loss
        loss_dict.render = crit(output_dict.rgb.reshape(n_img, -1, 3), image)
        if "rgb_fine" in output_dict.keys():
            loss_dict.render = loss_dict.render + crit(output_dict.rgb_fine.reshape(n_img, -1, 3), image)
        if opt.loss_weight.fg_mask is not None:
            m = 0.5 * torch.abs(fg_mask - output_dict.opacity.reshape(n_img, -1, 1)).mean()
            if "opacity_fine" in output_dict.keys():
                m = m + 0.5 * torch.abs(fg_mask - output_dict.opacity_fine.reshape(n_img, -1, 1)).mean()
            loss_dict.fg_mask = m
        loss_dict = self.compute_regularization_losses(opt, output_dict, loss_dict)
        return loss_dict, {}, {}

    def compute_regularization_losses(self, opt, output_dict, loss):
        """Distortion (mip-NeRF 360) and depth-patch smoothness terms, default off (base_losses.py:162-194;
        regularization_losses.py:20-66).  The distortion loss is one O(S) kernel per network instead of the
        reference's [S-1, S-1] matrix per ray."""
        if opt.loss_weight.distortion is not None:
            strength = 1e-3 * 2
            v = strength * ops.distortion_loss(output_dict["t"], output_dict["weights"])
            if "weights_fine" in output_dict:
                v = v + strength * ops.distortion_loss(output_dict["t_fine"], output_dict["weights_fine"])
            if "distortion" in loss.keys():
                loss["distortion"] = (loss["distortion"] + v) / 2.0
            else:
                loss["distortion"] = v
        if opt.loss_weight.depth_patch is not None:
            strength = 0.01 * 2

            def patch(depths):      # a few thousand elements: plain device tensor algebra
                B = depths.shape[0]
                d = depths.reshape(B, -1, self.opt.depth_regu_patch_size ** 2)
                return torch.sqrt((d[..., None] - d[..., None, :]) ** 2 + 0.001 ** 2).mean()
            v = strength * patch(output_dict["depth"])
            if "depth_fine" in output_dict.keys():
                v = v + strength * patch(output_dict["depth_fine"])
            if "depth_patch" in loss.keys():
                loss["depth_patch"] = (loss["depth_patch"] + v) / 2.0
            else:
                loss["depth_patch"] = v
        return loss


# ------------------------------------------------------------------------------------------------
# multi-view correspondence loss
# ------------------------------------------------------------------------------------------------
class CorrespondencesPairRenderDepthAndGet3DPtsAndReproject(BaseLoss):
    """Re-projection error between pre-computed dense correspondences, using the RENDERED depth and the
    current pose estimates (corres_loss.py:29-220, base_corres_loss.py:28-375)."""

    DEFAULTS = dict(matching_pair_generation="all", min_nbr_matches=500, pairing_angle_threshold=30,
                    filter_corr_w_cc=False, min_conf_valid_corr=0.95, min_conf_cc_valid_corr=1.0 / 2.5,
                    diff_loss_type="huber", compute_photo_on_matches=False,
                    renderrepro_do_pixel_reprojection_check=False, renderrepro_do_depth_reprojection_check=False,
                    renderrepro_pixel_reprojection_thresh=10.0, renderrepro_depth_reprojection_thresh=0.1,
                    use_gt_depth=False, use_gt_correspondences=False, use_dummy_all_one_confidence=False)

    # device_side = True: the per-step work takes NO host decision and has fixed shapes, so a whole SPARF
    # step is sync-free and capturable into one CUDA graph: the image pair is drawn with the device generator, the
    # valid correspondences are sub-sampled by a top-k over random keys (a uniformly random subset of the valid set,
    # like the reference's randperm) into a fixed-capacity buffer of rand_rays // 2 points + a validity mask, and the
    # loss normalises by the device-side count.  The random stream differs from the reference's; the arithmetic per
    # selected point does not (tests/test_losses.py checks both modes against each other with injected choices).
    device_side = False

    def __init__(self, opt, nerf_net, flow_net, train_data, device):
        super().__init__(device)
        self.opt = _with_defaults(self.DEFAULTS, opt)
        self.net, self.flow_net, self.train_data = nerf_net, flow_net, train_data
        H, W = train_data.all.image.shape[-2:]
        xx = torch.arange(0, W).view(1, -1).repeat(H, 1)
        yy = torch.arange(0, H).view(-1, 1).repeat(1, W)
        self.grid = torch.stack((xx, yy), dim=-1).to(device).float()                     # [H,W,2]
        self.grid_flat = (self.grid[:, :, 1] * W + self.grid[:, :, 0]).to(device).long()
        self.compute_correspondences(train_data)

    @torch.no_grad()
    def compute_correspondences(self, train_data):
        """Dense correspondence + confidence maps for the view pairs, once (base_corres_loss.py:65-149)."""
        images = train_data.all["image"]
        H, W = images.shape[-2:]
        n_views = images.shape[0]
        how = self.opt.matching_pair_generation
        if how == "all":
            combi = generate_pair_list(n_views)
        elif how == "all_to_all":
            combi = self.flow_net.combi_list
        else:
            raise NotImplementedError("matching_pair_generation=%r" % how)
        if combi.shape[1] == 0:
            self.corres_maps = self.conf_maps = self.mask_valid_corr = None
            self.filtered_flow_pairs = []
            return
        corres, conf, _ = self.flow_net.compute_flow_and_confidence_map_of_combi_list(
            images, combi_list_tar_src=combi, plot=True, use_homography=self.opt.use_homography_flow)
        c = corres.reshape(-1, 2, H, W).permute(0, 2, 3, 1)
        inside = c[..., 0].ge(0) & c[..., 0].le(W - 1) & c[..., 1].ge(0) & c[..., 1].le(H - 1)
        mask = conf.reshape(-1, 1, H, W).ge(self.opt.min_conf_valid_corr) & inside.unsqueeze(1)   # [P,1,H,W]
        self.corres_maps, self.conf_maps, self.mask_valid_corr = corres, conf, mask
        self.flow_pairs = combi.cpu().numpy().T.tolist()
        self.filtered_flow_pairs = [(i, p[0], p[1]) for i, p in enumerate(self.flow_pairs)
                                    if mask[i].sum() > self.opt.min_nbr_matches]

    def sample_valid_image_pair(self):
        k = np.random.randint(len(self.filtered_flow_pairs))
        i, id_self, id_other = self.filtered_flow_pairs[k]
        return (id_self, id_other, self.corres_maps[i].permute(1, 2, 0)[:, :, :2], self.conf_maps[i].permute(1, 2, 0), None,
                self.mask_valid_corr[i].permute(1, 2, 0))

    def compute_loss(self, opt, data_dict, output_dict, iteration, mode=None, plot=False, **kwargs):
        if mode != "train":
            return {}, {}, {}
        loss_dict, stats, plots = self.compute_loss_pairwise(opt, data_dict, output_dict, iteration, mode, plot)
        if self.opt.gradually_decrease_corres_weight:
            start = self.opt.ratio_start_decrease_corres_weight * self.opt.max_iter \
                if self.opt.ratio_start_decrease_corres_weight is not None else self.opt.iter_start_decrease_corres_weight
            gamma = 1.0 if iteration < start else 2 ** ((iteration - start) // self.opt.corres_weight_reduct_at_x_iter)
            loss_dict["corres"] = loss_dict["corres"] / gamma
        return loss_dict, stats, plots

    # ---- device-side mode -------------------------------------------------------------------------------------
    def _rand_pair(self):
        """index into filtered_flow_pairs as a [1] device tensor"""
        return torch.randint(len(self.filtered_flow_pairs), (1,), device=self.device)

    def _rand_keys(self, n):
        return torch.rand(n, device=self.device)

    def _device_tables(self):
        if not hasattr(self, "_pair_tab"):
            tab = torch.tensor([[i, a, b] for i, a, b in self.filtered_flow_pairs], dtype=torch.long, device=self.device)
            self._pair_tab = tab.reshape(-1, 3)
            self._grid_px = self.grid.reshape(-1, 2)
        return self._pair_tab

    def _compute_loss_pairwise_device(self, opt, data_dict, iteration):
        B, _, H, W = data_dict.image.shape
        tab = self._device_tables()
        row = tab.index_select(0, self._rand_pair())[0]
        i_map, id_self, id_other = row[0:1], row[1:2], row[2:3]
        corres = self.corres_maps.index_select(0, i_map)[0].permute(1, 2, 0)[:, :, :2].reshape(-1, 2).detach()
        conf = self.conf_maps.index_select(0, i_map)[0].permute(1, 2, 0).reshape(-1, 1).detach()
        mask = self.mask_valid_corr.index_select(0, i_map)[0, 0]
        if iteration < self.opt.precrop_iters:
            dH, dW = int(H // 2 * self.opt.precrop_frac), int(W // 2 * self.opt.precrop_frac)
            center = torch.zeros_like(mask)
            center[H // 2 - dH:H // 2 + dH - 1, W // 2 - dW:W // 2 + dW - 1].fill_(1)   # (fill_: no host scalar copy)
            mask = mask & center
        mask = mask.reshape(-1)
        enough = (mask.sum() >= self.opt.min_nbr_matches).float()        # corres_loss / base_corres_loss early return
        n = min(self.opt.nerf.rand_rays // 2, H * W)
        keys = torch.where(mask, self._rand_keys(H * W), torch.full((), 2.0, device=self.device))
        val, idx = torch.topk(keys, n, largest=False)                    # random subset of the valid pixels, padded
        valid = (val < 1.5)[:, None]                                     # [n,1]
        px_self = self._grid_px.index_select(0, idx)
        px_other = torch.where(valid, corres.index_select(0, idx), px_self)   # padding: any finite in-image pixel
        conf_v = conf.index_select(0, idx)
        poses = data_dict.poses_w2c
        bottom = _bottom_row(poses.device)[None]
        P_self = torch.cat((poses.index_select(0, id_self)[0], bottom), 0)
        P_other = torch.cat((poses.index_select(0, id_other)[0], bottom), 0)
        K_self, K_other = data_dict.intr.index_select(0, id_self)[0], data_dict.intr.index_select(0, id_other)[0]
        loss_dict, stats, _ = self.compute_loss_on_image_pair(data_dict, P_self, P_other, K_self, K_other, None, None, None,
                                                              {"render_matches": torch.zeros((), device=self.device)}, {}, {},
                                                              selected=(px_self, px_other, conf_v, valid))
        loss_dict["corres"] = loss_dict["corres"] * enough
        stats["perc_valid_corr_mask"] = mask.sum() / (mask.nelement() + 1e-6)
        return loss_dict, stats, {}

    def compute_loss_pairwise(self, opt, data_dict, output_dict, iteration, mode=None, plot=False):
        loss_dict = {"corres": torch.zeros((), device=self.device, requires_grad=True),
                     "render_matches": torch.zeros((), device=self.device, requires_grad=True)}
        if mode != "train" or iteration < self.opt.start_iter.corres or len(self.filtered_flow_pairs) == 0:
            return loss_dict, {}, {}
        if self.device_side:
            return self._compute_loss_pairwise_device(opt, data_dict, iteration)
        id_self, id_other, corres_map, conf_map, _, mask = self.sample_valid_image_pair()
        if iteration < self.opt.precrop_iters:
            H, W = data_dict.image.shape[-2:]
            dH, dW = int(H // 2 * self.opt.precrop_frac), int(W // 2 * self.opt.precrop_frac)
            center = torch.zeros_like(mask)
            center[H // 2 - dH:H // 2 + dH - 1, W // 2 - dW:W // 2 + dW - 1] = 1
            mask = mask & center
        return self.compute_loss_at_given_img_indexes(opt, data_dict, id_self, id_other, corres_map, conf_map, None, mask,
                                                      loss_dict, {}, {}, plot)

    def compute_loss_at_given_img_indexes(self, opt, data_dict, id_self, id_matching_view, corres_map_self_to_other_,
                                          conf_map_self_to_other_, variance_self_to_other_, mask_correct_corr, loss_dict,
                                          stats_dict, plotting_dict, plot=False, skip_verif=True):
        B, _, H, W = data_dict.image.shape
        poses = data_dict.poses_w2c
        P_self = torch.eye(4).to(poses.device)
        P_self[:3, :4] = poses[id_self]
        P_other = torch.eye(4).to(poses.device)
        P_other[:3, :4] = poses[id_matching_view]
        corres = corres_map_self_to_other_.detach()
        conf = conf_map_self_to_other_.detach()
        mask = mask_correct_corr.detach().squeeze(-1)
        if mask.sum() < self.opt.min_nbr_matches:
            return loss_dict, stats_dict, plotting_dict
        stats_dict["perc_valid_corr_mask"] = mask.sum() / (mask.nelement() + 1e-6)
        return self.compute_loss_on_image_pair(data_dict, P_self, P_other, data_dict.intr[id_self],
                                               data_dict.intr[id_matching_view], corres, conf, mask, loss_dict,
                                               stats_dict, plotting_dict)

    def _reprojection(self, px_i, depth_i, K_i, px_j, depth_j, K_j, T_i2j, conf, stats, valid0=None):
        """compute_render_and_repro_loss_w_repro_thres (corres_loss.py:50-95)."""
        proj, depth_proj = batch_project_to_other_img(px_i.float(), depth_i, K_i, K_j, T_i2j, return_depth=True)
        err = torch.norm(proj - px_j, dim=-1, keepdim=True)
        valid = torch.ones_like(err).bool() if valid0 is None else valid0
        if self.opt.renderrepro_do_pixel_reprojection_check:
            ok = err.detach().le(self.opt.renderrepro_pixel_reprojection_thresh)
            valid = valid & ok
            stats["perc_val_pix_rep"] = ok.sum().float() / (ok.nelement() + 1e-6)
        if self.opt.renderrepro_do_depth_reprojection_check:
            ok = (torch.abs(depth_j - depth_proj) / (depth_j + 1e-6)).detach().le(self.opt.renderrepro_depth_reprojection_thresh)
            valid = valid & ok.unsqueeze(-1)
            stats["perc_val_depth_rep"] = ok.sum().float() / (ok.nelement() + 1e-6)
        return self.compute_diff_loss(self.opt.diff_loss_type, proj - px_j, weights=conf, mask=valid, dim=-1)

    def compute_loss_on_image_pair(self, data_dict, P_self, P_other, K_self, K_other, corres, conf, mask, loss_dict,
                                   stats_dict, plotting_dict, selected=None):
        iteration = data_dict["iter"]
        H, W = data_dict.image.shape[-2:]
        valid0 = None
        if selected is not None:             # device-side mode: fixed-capacity selection + validity mask
            px_self, px_other, conf_v, valid0 = selected
        else:
            px_self = self.grid[mask]
            px_other = corres[mask]
            conf_v = conf[mask]
            half = self.opt.nerf.rand_rays // 2
            if px_self.shape[0] > half:                                      # corres_loss.py:149-157
                sel = torch.randperm(px_self.shape[0], device=self.device)[:half].to(px_self.device)
                px_self, px_other, conf_v = px_self[sel], px_other[sel], conf_v[sel]
        ret_self = self.net.render_image_at_specific_pose_and_rays(self.opt, data_dict, P_self[:3], K_self, H, W,
                                                                   pixels=px_self, mode="train", iter=iteration)
        ret_other = self.net.render_image_at_specific_pose_and_rays(self.opt, data_dict, P_other[:3], K_other, H, W,
                                                                    pixels=px_other, mode="train", iter=iteration)
        if self.opt.compute_photo_on_matches:
            raise NotImplementedError("compute_photo_on_matches (default off)")
        T_s2o = P_other @ pose_inverse_4x4(P_self)
        T_o2s = pose_inverse_4x4(T_s2o)
        keys = ["depth"] + (["depth_fine"] if "depth_fine" in ret_other.keys() else [])
        stats_dict["depth_in_corr_loss"] = ret_self.depth.detach().mean()
        total = 0.0
        for k in keys:
            d_s, d_o = ret_self[k].squeeze(0).squeeze(-1), ret_other[k].squeeze(0).squeeze(-1)
            total = total + self._reprojection(px_self, d_s, K_self, px_other, d_o, K_other, T_s2o, conf_v, stats_dict, valid0)
            total = total + self._reprojection(px_other, d_o, K_other, px_self, d_s, K_self, T_o2s, conf_v, stats_dict, valid0)
        loss_dict["corres"] = total / (2.0 * len(keys))
        return loss_dict, stats_dict, plotting_dict


# ------------------------------------------------------------------------------------------------
# depth-consistency loss
# ------------------------------------------------------------------------------------------------
class DepthConsistencyLoss(BaseLoss):
    """Pseudo-depth supervision at an unseen pose interpolated between two training poses
    (depth_cons_loss.py:32-321)."""

    DEFAULTS = dict(gradually_decrease_geo_sampling_loss=False, geo_sampling_loss_reduct_at_x_iter=10000,
                    diff_loss_type="huber")

    # device_side = True: no host decision, fixed shapes (see the correspondence loss): the reference view, the
    # interpolation weight and the pixels are drawn with the device generator, the nearest camera is an argmin on the
    # device, and the two data-dependent filters (inside the virtual image & in front of the near plane; visibility
    # >= 0.2) become a validity mask over ALL sampled points (invalid ones are moved to a safe dummy point, rendered and
    # weighted 0); the mean divides by the device-side count of valid points.
    device_side = False

    def __init__(self, opt, nerf_net, device):
        super().__init__(device)
        self.opt = _with_defaults(self.DEFAULTS, opt)
        self.net = nerf_net

    # ---- device-side mode -------------------------------------------------------------------------------------
    def _rand_image(self, B):
        return torch.randint(B, (1,), device=self.device)

    def _rand_weight(self):
        return torch.rand((), device=self.device)

    def _rand_pixels(self, H, W, n):
        return sample_rays(H, W, nbr=n, fraction_in_center=self.opt.sampled_fraction_in_center, device=self.device)[0]

    def _compute_loss_device(self, opt, data_dict, iteration):
        B, _, H, W = data_dict.image.shape
        bottom = _bottom_row(self.device).reshape(1, 1, -1).repeat(B, 1, 1)
        poses_w2c = torch.cat((data_dict.poses_w2c.detach(), bottom.to(data_dict.poses_w2c.dtype)), dim=1)
        poses_c2w = pose_inverse_4x4(poses_w2c)
        id_self = self._rand_image(B)
        px_ref = self._rand_pixels(H, W, max(1024, self.opt.nerf.rand_rays)).reshape(-1, 2)
        K_ref, P_ref = data_dict.intr.index_select(0, id_self)[0], poses_w2c.index_select(0, id_self)[0]
        ret_ref = self.net.render_image_at_specific_pose_and_rays(self.opt, data_dict, pose=P_ref[:3], intr=K_ref, H=H, W=W,
                                                                  pixels=px_ref, mode="train", iter=iteration)
        use_fine = "depth_fine" in ret_ref.keys()
        if use_fine and hasattr(opt.nerf, "ratio_start_fine_sampling_at_x") and opt.nerf.ratio_start_fine_sampling_at_x is not None \
                and iteration < opt.max_iter * (opt.nerf.ratio_start_fine_sampling_at_x + 0.05):
            use_fine = False
        depth_ref = (ret_ref.depth_fine if use_fine else ret_ref.depth).squeeze(0).squeeze(-1)
        c2w_self = poses_c2w.index_select(0, id_self)[0]
        pts_w = batch_backproject_to_3d(px_ref, depth_ref, K_ref, c2w_self)
        # nearest other camera by the angle between camera-position vectors (data_utils.py:267-311), on the device
        pos = poses_c2w[:, :3, 3]
        unit = pos / (pos.norm(dim=1, keepdim=True) + 1e-6)
        ang = torch.acos((unit * unit.index_select(0, id_self)).sum(-1).clamp(-1.0, 1.0))
        ang = ang.scatter(0, id_self, torch.full((1,), 1e3, device=self.device))
        id_other = torch.argmin(ang, dim=0, keepdim=True)
        w = self._rand_weight()
        unseen_c2w = w * c2w_self.detach() + (1 - w) * poses_c2w.index_select(0, id_other)[0].detach()
        P_unseen = pose_inverse_4x4(unseen_c2w)
        K = K_ref.clone()
        near = data_dict.depth_range[0][0]
        with torch.no_grad():
            px0, z0 = batch_project(pts_w, P_unseen, K, return_depth=True)
            ok = px0[:, 0].ge(0.0) & px0[:, 1].ge(0.0) & px0[:, 0].le(W - 1) & px0[:, 1].le(H - 1) & z0.ge(near)
            # a dummy world point for the rejected ones: on the virtual camera's axis, one unit behind the near plane
            safe = (unseen_c2w[:3, :3] @ torch.stack([torch.zeros_like(near), torch.zeros_like(near), near + 1.0]) + unseen_c2w[:3, 3])
        pts_s = torch.where(ok[:, None], pts_w, safe[None].expand_as(pts_w))
        px, z = batch_project(pts_s, P_unseen, K, return_depth=True)      # finite everywhere, grad only through valid points
        with torch.no_grad():
            vis_ret = self.net.render_up_to_maxdepth_at_specific_pose_and_rays(
                self.opt, data_dict, P_unseen[:3], K, H, W, depth_max=z, pixels=px, mode="train", iter=iteration)
            key = "all_cumulated_fine" if "all_cumulated_fine" in vis_ret.keys() else "all_cumulated"
            vis = vis_ret[key].squeeze(0).unsqueeze(-1)
        valid = ok & vis.ge(0.2).reshape(-1)
        ret = self.net.render_image_at_specific_pose_and_rays(self.opt, data_dict, P_unseen[:3], K, H, W, pixels=px,
                                                              mode="train", iter=iteration)
        total = 0.0
        for suf in [""] + (["_fine"] if "rgb_fine" in ret.keys() else []):
            depth = ret["depth" + suf].squeeze().reshape(-1)
            wgt = vis * ret["opacity" + suf].squeeze(0).detach()
            total = total + self.compute_diff_loss(self.opt.diff_loss_type, diff=z.view(-1) - depth.view(-1), weights=wgt.view(-1),
                                                   mask=valid)
        stats = {"nbr_px_sampling": valid.sum(), "avg_vis_weight": (wgt.view(-1) * valid).sum() / (valid.sum() + 1e-6)}
        loss = {"depth_cons": total}
        if self.opt.gradually_decrease_depth_cons_loss:
            loss["depth_cons"] = loss["depth_cons"] / (2 ** (iteration // self.opt.depth_cons_loss_reduct_at_x_iter))
        return loss, stats, {}

    def sample_pose(self, poses_c2w, id_self, pose_w2c_self):
        """w * own + (1-w) * nearest other camera, on the 4x4 matrices (depth_cons_loss.py:45-63).  The
        reference moves the poses to the host for the neighbour search; so do we (a [B,4,4] copy)."""
        host = poses_c2w.detach().cpu().numpy()
        id_other = get_nearest_pose_ids(host[id_self], host, tar_id=id_self)
        w = np.random.rand()
        own = pose_inverse_4x4(pose_w2c_self).detach()
        unseen = w * own + (1 - w) * poses_c2w[id_other].detach()
        return pose_inverse_4x4(unseen)

    def _start_iter(self, opt):
        return opt.start_ratio.depth_cons * opt.max_iter if opt.start_ratio.depth_cons is not None else opt.start_iter.depth_cons

    def compute_loss(self, opt, data_dict, output_dict, iteration, mode=None, plot=False, **kwargs):
        if mode != "train" or iteration < self._start_iter(opt):
            return {}, {}, {}
        if self.device_side:
            return self._compute_loss_device(opt, data_dict, iteration)
        B, _, H, W = data_dict.image.shape
        bottom = torch.tensor([0, 0, 0, 1], device=self.device).reshape(1, 1, -1).repeat(B, 1, 1)
        poses_w2c = torch.cat((data_dict.poses_w2c.detach(), bottom.to(data_dict.poses_w2c.dtype)), dim=1)
        poses_c2w = pose_inverse_4x4(poses_w2c)
        id_self = np.random.randint(B)
        px_ref, _ = sample_rays(H, W, nbr=max(1024, self.opt.nerf.rand_rays),
                                fraction_in_center=self.opt.sampled_fraction_in_center)
        px_ref = px_ref.reshape(-1, 2).to(self.device)
        K_ref, P_ref = data_dict.intr[id_self], poses_w2c[id_self]
        ret_ref = self.net.render_image_at_specific_pose_and_rays(self.opt, data_dict, pose=P_ref[:3], intr=K_ref, H=H, W=W,
                                                                  pixels=px_ref, mode="train", iter=iteration)
        use_fine = "depth_fine" in ret_ref.keys()
        if use_fine and hasattr(opt.nerf, "ratio_start_fine_sampling_at_x") and opt.nerf.ratio_start_fine_sampling_at_x is not None \
                and iteration < opt.max_iter * (opt.nerf.ratio_start_fine_sampling_at_x + 0.05):
            use_fine = False
        depth_ref = (ret_ref.depth_fine if use_fine else ret_ref.depth).squeeze(0).squeeze(-1)
        pts_w = batch_backproject_to_3d(px_ref, depth_ref, K_ref, poses_c2w[id_self])
        P_unseen = self.sample_pose(poses_c2w, id_self, P_ref)
        return self.compute_loss_at_sampled_pose(data_dict, P_unseen, K_ref.clone(), pts_w, H, W, px_ref)

    def compute_loss_at_sampled_pose(self, data_dict, pose_w2c_at_unseen, intr_at_unseen, pseudo_gt_3dpts_in_w, H, W,
                                     pixels_in_ref):
        iteration = data_dict["iter"]
        stats = {"nbr_px_sampling": pseudo_gt_3dpts_in_w.shape[0]}
        zero = {"depth_cons": torch.tensor(0.0, requires_grad=True).to(self.device)}
        px, z = batch_project(pseudo_gt_3dpts_in_w, pose_w2c_at_unseen, intr_at_unseen, return_depth=True)
        ok = px[:, 0].ge(0.0) & px[:, 1].ge(0.0) & px[:, 0].le(W - 1) & px[:, 1].le(H - 1) & z.ge(data_dict.depth_range[0][0])
        px, z = px[ok], z[ok]
        if z.shape[0] == 0:
            return zero, {}, {}
        with torch.no_grad():   # visibility: transmittance up to the pseudo depth (depth_cons_loss.py:266-277)
            vis_ret = self.net.render_up_to_maxdepth_at_specific_pose_and_rays(
                self.opt, data_dict, pose_w2c_at_unseen[:3], intr_at_unseen, H, W, depth_max=z, pixels=px, mode="train",
                iter=iteration)
            key = "all_cumulated_fine" if "all_cumulated_fine" in vis_ret.keys() else "all_cumulated"
            vis = vis_ret[key].squeeze(0).unsqueeze(-1)
            assert vis.le(1.0).all()
        keep = vis.ge(0.2).reshape(-1)
        px, z, vis = px[keep], z[keep], vis[keep]
        if z.shape[0] == 0:
            return zero, {}, {}
        ret = self.net.render_image_at_specific_pose_and_rays(self.opt, data_dict, pose_w2c_at_unseen[:3], intr_at_unseen,
                                                              H, W, pixels=px, mode="train", iter=iteration)
        total = 0.0
        for suf in [""] + (["_fine"] if "rgb_fine" in ret.keys() else []):
            depth = ret["depth" + suf].squeeze().reshape(-1)
            wgt = vis * ret["opacity" + suf].squeeze(0).detach()
            total = total + self.compute_diff_loss(self.opt.diff_loss_type, diff=z.view(-1) - depth.view(-1), weights=wgt.view(-1))
        stats["avg_vis_weight"] = wgt.sum() / (wgt.nelement() + 1e-6)
        loss = {"depth_cons": total}
        if self.opt.gradually_decrease_depth_cons_loss:
            loss["depth_cons"] = loss["depth_cons"] / (2 ** (iteration // self.opt.depth_cons_loss_reduct_at_x_iter))
        return loss, stats, {}


# ------------------------------------------------------------------------------------------------
# DS-NeRF sparse-depth loss
# ------------------------------------------------------------------------------------------------
class SparseCOLMAPDepthLoss(BaseLoss):
    """Weighted squared error between the rendered depth and sparse COLMAP depth, rendered per image at the pixels
    that carry a triangulated point (base_losses.py:326-402).  `output_dict` is unused, as in the reference."""

    def __init__(self, opt, nerf_net, device):
        super().__init__(device)
        self.opt, self.net = opt, nerf_net

    def compute_loss(self, opt, data_dict, output_dict, iteration, mode=None, plot=False, **kwargs):
        if mode != "train":
            return {}, {}, {}
        H, W = data_dict.image.shape[-2:]
        pose = self.net.get_w2c_pose(opt, data_dict, mode=mode)
        intr = data_dict.intr
        depth_maps = data_dict.colmap_depth.view(-1, H, W)
        conf_maps = data_dict.colmap_conf.view(-1, H, W)
        B = pose.shape[0]
        stats = {"perc_col_depth": (depth_maps > 0).sum() / depth_maps.nelement()}
        per_image = opt.nerf.rand_rays // B
        total = torch.zeros((), device=self.device)
        for i in range(B):
            ys, xs = torch.where(depth_maps[i] > 1e-6)
            if len(ys) == 0:
                continue
            if len(ys) > per_image:
                sel = torch.randperm(len(ys), device=self.device)[:per_image]
                ys, xs = ys[sel], xs[sel]
            d_ref = depth_maps[i][ys, xs].reshape(-1)
            w_ref = conf_maps[i][ys, xs].reshape(-1)
            ret = self.net.render_image_at_specific_pose_and_rays(self.opt, data_dict, pose[i], intr[i], H, W,
                                                                  ray_idx=ys * W + xs, mode="train", iter=iteration)
            total = total + torch.mean(((d_ref - ret.depth.reshape(-1)) ** 2) * w_ref)
            if "depth_fine" in ret.keys():
                total = total + torch.mean(((d_ref - ret.depth_fine.reshape(-1)) ** 2) * w_ref)
        return edict(colmap_depth=0.1 * total / B), stats, {}      # DS-NeRF's weighting


def define_loss(loss_type: str, opt, nerf_net, train_data, device, flow_net=None, device_side: bool = False) -> Loss:
    """loss_factory.py:25-42.  device_side=True switches the correspondence / depth-consistency modules to their
    sync-free fixed-shape mode and drops the per-key host asserts of the aggregator (one CUDA graph per SPARF step)."""
    mods = []
    if "photometric" in loss_type:
        mods.append(BasePhotoandReguLoss(opt, nerf_net, train_data=train_data, device=device))
    if "SparseCOLMAPDepthLoss" in loss_type:
        mods.append(SparseCOLMAPDepthLoss(opt, nerf_net, device=device))
    if "corres" in loss_type:
        mods.append(CorrespondencesPairRenderDepthAndGet3DPtsAndReproject(opt, nerf_net, flow_net=flow_net,
                                                                          train_data=train_data, device=device))
    if "depth_cons" in loss_type:
        mods.append(DepthConsistencyLoss(opt, nerf_net, device=device))
    agg = Loss(mods)
    if device_side:
        agg.check_finite = False
        for m in mods:
            if hasattr(m, "device_side"):
                m.device_side = True
    return agg
