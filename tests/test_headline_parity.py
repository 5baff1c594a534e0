"""Graph-level parity at the HEADLINE shape (BASELINE config 2/3: 3 views 300x400, 3 x 341 = 1023 rays x 128 samples):
outputs, photometric loss and the gradients of all 20 MLP tensors + the 9-D pose embedding, through the public API
(`Graph.render_image_at_specific_rays` + the `BasePhotoandReguLoss` mirror), against the oracle.

The golden fixtures are small (the CPU reference has to finish in seconds); here the checker is the oracle itself, run
on the GPU box in fp64 (exact for this purpose) and in fp32 (= the reference's arithmetic, pinned bit-exact to it by
tests/test_oracle_vs_golden.py).  Gate: our distance from the exact result may not exceed twice the distance of the
reference's own fp32 arithmetic from it (floors: 1e-4 on outputs -- north_star's bound --, 2e-3 on gradients).
"""
import numpy as np
import pytest
import torch

import common
from helpers import rel_err

pytestmark = pytest.mark.gpu


def _oracle_step(opt, sd, data, init_w2c, ray_idx, dtype, device):
    from oracle import sparf_oracle as O
    cast = lambda x: x.to(device=device, dtype=dtype) if torch.is_floating_point(x) else x.to(device)
    params = {k: cast(v).clone().requires_grad_(k != "progress") for k, v in sd.items()}
    emb = O.pose_to_d9(init_w2c).to(device=device, dtype=dtype).clone().requires_grad_(True)
    pose = O.d9_to_pose(emb)
    center, ray = O.rays_from_ray_idx(pose, cast(data.intr), data.image.shape[-2], data.image.shape[-1], ray_idx.to(device))
    out = O.render(opt, params, None, center, ray, cast(data.depth_range[0]), mode="train", iteration=10)
    loss = O.photometric_loss(out, cast(data.image), ray_idx.to(device))
    loss.backward()
    grads = {k: v.grad.detach().double().cpu() for k, v in params.items() if k != "progress"}
    grads["pose_embedding"] = emb.grad.detach().double().cpu()
    outs = {k: out[k].detach().double().cpu() for k in ("rgb", "depth", "opacity")}
    return outs, float(loss), grads


@pytest.mark.parametrize("per_image", [False, True])
@pytest.mark.parametrize("engine", ["simt_fp32", "tc_3x"])
def test_headline_graph_loss_and_grads(engine, per_image):
    import sparf_b200
    from sparf_b200.losses import BasePhotoandReguLoss
    from sparf_b200.poses_models import FirstTwoColunmnsPoseParameters
    from sparf_b200.renderer import Graph

    sparf_b200.set_engine(engine)
    dev = torch.device("cuda")
    B, H, W, n, S = 3, 300, 400, 341, 128
    opt = common.make_opt(S=S, rand_rays=1024, barf_c2f=(0.1, 0.5))
    sd = common.det_weights(opt, 21, peaky=True, sigma_bias=-3.0, progress=0.35)
    data = common.make_scene(21, B, H, W, focal=400.0)
    data.depth_range = torch.tensor([[1.2, 5.2]] * B)
    init_w2c = common.perturb_poses(data.pose, 21, sigma=0.02)
    rng = np.random.default_rng(21)
    if per_image:    # RaySamplingStrategy-style per-image indices (B,n)
        ray_idx = torch.from_numpy(np.stack([rng.permutation(H * W)[:n] for _ in range(B)]).astype(np.int64))
    else:
        ray_idx = torch.from_numpy(rng.permutation(H * W)[:n].astype(np.int64))

    exact_out, exact_loss, exact_g = _oracle_step(opt, sd, data, init_w2c, ray_idx, torch.float64, dev)
    ref_out, ref_loss, ref_g = _oracle_step(opt, sd, data, init_w2c, ray_idx, torch.float32, dev)

    for k in ("image", "intr", "pose", "depth_range", "idx"):
        data[k] = data[k].to(dev)
    pose_net = FirstTwoColunmnsPoseParameters(opt, nbr_poses=B, initial_poses_w2c=init_w2c.to(dev), device=dev).to(dev)

    class PoseGraph(Graph):
        def get_w2c_pose(self, opt, data_dict, mode=None):
            return pose_net.get_w2c_poses()

    net = PoseGraph(opt, dev)
    net.nerf.load_state_dict(sd)
    net.to(dev).train()
    out = net.render_image_at_specific_rays(opt, data, iter=10, ray_idx=ray_idx.to(dev), mode="train")
    loss_mod = BasePhotoandReguLoss(opt, net, train_data=None, device=dev)
    loss = loss_mod.compute_loss(opt, data, out, iteration=10, mode="train")[0].render
    loss.backward()
    torch.cuda.synchronize()

    rep = {}
    for k in ("rgb", "depth", "opacity"):
        ours = rel_err(out[k].detach().cpu().reshape(exact_out[k].shape), exact_out[k])
        ref = rel_err(ref_out[k], exact_out[k])
        rep[k] = (ours, ref)
        assert ours <= max(2 * ref, 1e-4), (k, ours, ref)
    assert abs(float(loss) - exact_loss) <= max(2 * abs(ref_loss - exact_loss), 1e-5 * abs(exact_loss)), (float(loss), exact_loss, ref_loss)
    got = {k: p.grad.detach().double().cpu() for k, p in net.nerf.named_parameters() if k != "progress"}
    got["pose_embedding"] = pose_net.pose_embedding.grad.detach().double().cpu()
    for k, ex in exact_g.items():
        scale = float(ex.abs().max().clamp_min(1e-30))
        ours = float((got[k] - ex).abs().max()) / scale
        ref = float((ref_g[k] - ex).abs().max()) / scale
        rep["grad " + k] = (ours, ref)
        assert ours <= max(2 * ref, 2e-3), (k, ours, ref)
    worst = max(rep.items(), key=lambda kv: kv[1][0] / max(kv[1][1], 1e-12))
    print(engine, "per-image idx" if per_image else "shared idx",
          "outputs", {k: "%.1e/%.1e" % rep[k] for k in ("rgb", "depth", "opacity")},
          "worst (ours/reference-fp32 vs fp64): %s %.1e/%.1e" % (worst[0], worst[1][0], worst[1][1]))
