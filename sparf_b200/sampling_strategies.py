"""Mirror of source/training/core/sampling_strategies.py: `RaySamplingStrategy` (the per-step ray-batch sampler of
`nerf_trainer.train_iteration`, :132-214) and the free function `sample_rays` (:250-295) the depth-consistency loss uses.

Same constructor / call signatures, same pools (all pixels, centre box, dilated foreground mask), same number and
order of `torch.randperm` draws, so a seeded run picks the same rays as the reference.  Everything the per-step call
touches lives on the device (the pools are built once); there is no host synchronisation in `__call__`, so the sampler
can sit inside the CUDA graph of a training step.  `sample_rays(..., device=...)` adds a device-side variant of the
reference's CPU function for the same purpose.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch


class RaySamplingStrategy:
    def __init__(self, opt: Dict[str, Any], data_dict: Dict[str, Any], device: torch.device):
        self.opt = opt
        self.device = device
        self.nbr_images, _, self.H, self.W = data_dict.image.shape
        self.all_possible_pixels = self.get_all_samples(data_dict)       # [HW', 2] (x, y), long
        self.all_center_pixels = self.get_all_center_pixels(data_dict)   # [N, 2]
        if self.opt.sample_fraction_in_fg_mask > 0.0:
            self.in_mask_pixels, self.min_nbr_in_mask = self.samples_in_mask(data_dict)
        p = self.opt.depth_regu_patch_size
        ys, xs = torch.meshgrid(torch.arange(p, dtype=torch.long, device=device),
                                torch.arange(p, dtype=torch.long, device=device), indexing="ij")
        self.dxdy = torch.stack([xs, ys], dim=-1).view(-1, 2)

    def samples_in_mask(self, data_dict):
        """Pixels inside the foreground masks dilated by 10 px (sampling_strategies.py:56-87; host work, once)."""
        import cv2
        assert "fg_mask" in data_dict.keys()
        B, _, H, W = data_dict.image.shape
        masks = []
        for b in range(B):
            m = data_dict.fg_mask[b].squeeze(0).cpu().numpy().astype(np.float32)
            masks.append(torch.from_numpy(cv2.dilate(m, np.ones((3, 3)), iterations=10) > 0))
        masks = torch.stack(masks, dim=0).to(self.device)
        p = self.opt.depth_regu_patch_size
        inner = torch.zeros_like(masks).bool()
        inner[:, :H - p - 1, :W - p - 1] = True
        ib, ih, iw = torch.where(masks & inner)
        per_image, smallest = [], float("inf")
        for b in range(B):
            sel = ib == b
            px = torch.stack((iw[sel], ih[sel]), dim=-1)
            smallest = min(smallest, len(px))
            per_image.append(px)
        return per_image, smallest

    def get_all_samples(self, data_dict) -> torch.Tensor:
        H, W = data_dict.image.shape[-2:]
        if self.opt.loss_weight.depth_patch is not None:      # keep the patch inside the image
            H, W = H - self.opt.depth_regu_patch_size - 1, W - self.opt.depth_regu_patch_size - 1
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.long, device=self.device),
                                torch.arange(W, dtype=torch.long, device=self.device), indexing="ij")
        return torch.stack([xs, ys], dim=-1).view(-1, 2).long()

    def get_all_center_pixels(self, data_dict) -> torch.Tensor:
        H, W = data_dict.image.shape[-2:]
        dH, dW = int(H // 2 * self.opt.precrop_frac), int(W // 2 * self.opt.precrop_frac)
        ys, xs = torch.meshgrid(torch.linspace(H // 2 - dH, H // 2 + dH - 1, 2 * dH),
                                torch.linspace(W // 2 - dW, W // 2 + dW - 1, 2 * dW), indexing="ij")
        return torch.stack([xs, ys], -1).view(-1, 2).long().to(self.device)

    def compute_pixel_coords_for_patch(self, pixel_coords: torch.Tensor) -> torch.Tensor:
        p2 = self.opt.depth_regu_patch_size ** 2
        shape = pixel_coords.shape[:-1]
        x = (pixel_coords.view(-1, 2)[..., 0][:, None].repeat(1, p2) + self.dxdy[:, 0]).reshape(-1)
        y = (pixel_coords.view(-1, 2)[..., 1][:, None].repeat(1, p2) + self.dxdy[:, 1]).reshape(-1)
        return torch.stack([x, y], dim=-1).reshape(shape + (p2, -1))

    def __call__(self, nbr_pixels: int, sample_in_center: bool = False, idx_imgs: Optional[List[int]] = None) -> torch.Tensor:
        """-> flat ray indices, (n,) shared by all images or (B, n) per image (foreground-mask sampling)."""
        nbr_images = self.nbr_images if idx_imgs is None else len(idx_imgs)
        per_img = nbr_pixels // nbr_images
        n_rand = nbr_pixels // nbr_images
        if self.opt.loss_weight.depth_patch is not None:
            per_img //= self.opt.depth_regu_patch_size ** 2
            n_rand //= self.opt.depth_regu_patch_size ** 2
        in_mask = in_center = None
        if self.opt.sample_fraction_in_fg_mask > 0.0:
            n_mask = min(self.min_nbr_in_mask, int(n_rand * self.opt.sample_fraction_in_fg_mask))
            n_rand -= n_mask
            ids = np.arange(nbr_images) if idx_imgs is None else idx_imgs
            in_mask = torch.stack([self.in_mask_pixels[i][torch.randperm(len(self.in_mask_pixels[i]), device=self.device)[:n_mask]]
                                   for i in ids], dim=0)
        elif self.opt.sampled_fraction_in_center > 0:
            n_mask = int(n_rand * self.opt.sampled_fraction_in_center)
            n_rand -= n_mask
            in_center = self.all_center_pixels[torch.randperm(len(self.all_center_pixels), device=self.device)[:n_mask]]
        pool = self.all_center_pixels if sample_in_center else self.all_possible_pixels
        px = pool[torch.randperm(len(pool), device=self.device)[:n_rand]]
        if in_mask is not None:
            px = torch.cat((px.unsqueeze(0).repeat(nbr_images, 1, 1), in_mask), dim=1)
        if in_center is not None:
            px = torch.cat((px, in_center), dim=0)
        if self.opt.loss_weight.depth_patch is not None:
            px = self.compute_pixel_coords_for_patch(px)
            px = px.reshape(nbr_images, -1, 2) if px.dim() == 4 else px.reshape(-1, 2)
        return px[..., 1] * self.W + px[..., 0]


def sample_rays(H: int, W: int, precrop_frac: float = 0.5, fraction_in_center: float = 0.0, nbr: Optional[int] = None,
                device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Random pixels of the (H-1) x (W-1) grid as float [N,2] (+ flat indices), sampling_strategies.py:250-295.
    device=None reproduces the reference (index pools and `torch.randperm` on the CPU generator); a CUDA device draws
    the permutations there instead (no host->device copy, capturable)."""
    ys, xs = torch.meshgrid(torch.arange(H - 1, device=device), torch.arange(W - 1, device=device), indexing="ij")
    x_ind, y_ind = xs.reshape(-1), ys.reshape(-1)
    if fraction_in_center > 0.0:
        dH, dW = int(H // 2 * precrop_frac), int(W // 2 * precrop_frac)
        Yc, Xc = torch.meshgrid(torch.linspace(H // 2 - dH, H // 2 + dH - 1, 2 * dH, device=device),
                                torch.linspace(W // 2 - dW, W // 2 + dW - 1, 2 * dW, device=device), indexing="ij")
        center = torch.stack([Xc, Yc], -1).view(-1, 2)
        if nbr is not None:
            n_c = int(nbr * fraction_in_center)
            idx = torch.randperm(len(x_ind), device=x_ind.device)[:nbr - n_c]
            x_ind, y_ind = x_ind[idx], y_ind[idx]
            idx = torch.randperm(len(center), device=x_ind.device)[:n_c]
            x_ind = torch.cat((x_ind, center[idx][..., 0]))
            y_ind = torch.cat((y_ind, center[idx][..., 1]))
    elif nbr is not None:
        idx = torch.randperm(len(x_ind), device=x_ind.device)[:nbr]
        x_ind, y_ind = x_ind[idx], y_ind[idx]
    px = torch.stack([x_ind, y_ind], dim=-1).reshape(len(x_ind), -1)
    return px.float(), px[..., 1] * W + px[..., 0]
