"""Fused clip + Adam + ExponentialLR update (sparf_b200.optim) against the reference's recipe:
torch.nn.utils.clip_grad_norm_ + torch.optim.Adam(betas=(0.9, 0.999)) + ExponentialLR (+ pose warm-up, + skipping a
non-finite gradient while the schedule keeps advancing)."""
import pytest
import torch


def _modules(seed, dev):
    torch.manual_seed(seed)
    a = torch.nn.Sequential(torch.nn.Linear(63, 256), torch.nn.Linear(256, 257)).to(dev)
    b = torch.nn.Linear(9, 3).to(dev)
    return a, b


@pytest.mark.gpu
@pytest.mark.parametrize("max_norm,warmup", [(None, 0), (0.1, 0), (0.05, 7)])
def test_fused_adam_matches_torch(max_norm, warmup):
    from sparf_b200.optim import FlatParameters, FusedAdam
    dev = torch.device("cuda")
    lr, lr_end, max_iter = 1e-3, 1e-4, 40
    gamma = (lr_end / lr) ** (1.0 / max_iter)
    ref_mods, our_mods = _modules(0, dev), _modules(0, dev)
    ref_params = [p for m in ref_mods for p in m.parameters()]
    ref_opt = torch.optim.Adam(ref_params, lr=lr, betas=(0.9, 0.999))
    sched = torch.optim.lr_scheduler.ExponentialLR(ref_opt, gamma=gamma)
    flat = FlatParameters(our_mods)
    ours = FusedAdam(flat, lr=lr, gamma=gamma, warmup_steps=warmup, max_norm=max_norm)
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(1, 31):
        grads = [torch.randn(p.shape, device=dev, generator=g) * (10.0 ** float(torch.randint(-4, 1, (1,)))) for p in ref_params]
        bad = it in (11, 12)            # two iterations with a NaN gradient: update skipped, schedule advances
        if bad:
            grads[1][0] = float("nan")
        for p, q, gr in zip(ref_params, flat.params, grads):
            p.grad = gr.clone()
            q.grad.copy_(gr)
        # ---- reference recipe (iter_based_trainer.py:128-147, joint_pose_nerf_trainer.py:513-549)
        if warmup:
            lr_orig = ref_opt.param_groups[0]["lr"]
            ref_opt.param_groups[0]["lr"] = lr_orig * min(1, it / warmup)
        finite = all(torch.isfinite(p.grad).all() for p in ref_params)
        if finite:
            if max_norm:
                torch.nn.utils.clip_grad_norm_(ref_params, max_norm)
            ref_opt.step()
        ref_opt.zero_grad()
        if warmup:
            ref_opt.param_groups[0]["lr"] = lr_orig
        sched.step()
        # ---- ours
        ours.step()
        flat.zero_grad()
        for p, q in zip(ref_params, flat.params):
            assert torch.isfinite(q).all()
            err = (p.data - q.data).abs().max().item()
            assert err <= 2e-6 * max(1.0, p.data.abs().max().item()), "iteration %d: %g" % (it, err)
    assert ours.steps.tolist() == [28, 30]


@pytest.mark.gpu
def test_full_iteration_in_one_cuda_graph():
    """render -> loss -> backward -> fused Adam captured once; replays follow the eager trajectory."""
    import common
    from helpers import build_graph
    from sparf_b200 import ops
    from sparf_b200.graphs import GraphedStep
    from sparf_b200.optim import FlatParameters, FusedAdam
    dev = torch.device("cuda")

    def make():
        net, _c, opt, data, _ridx, _px, _dm = build_graph("c1_coarse", dev)
        net.train()
        flat = FlatParameters([net])
        return opt, data, net, flat, FusedAdam(flat, lr=5e-4, gamma=0.999, max_norm=0.1)

    def body_of(opt, data, net, flat, adam, image_flat):
        def body(ray_idx):
            flat.zero_grad()
            out = net.render_image_at_specific_rays(opt, data, iter=0, ray_idx=ray_idx, mode="val")
            loss = ops.huber2(out.rgb, image_flat[:, ray_idx])
            loss.backward()
            adam.step()
            return loss.detach()
        return body

    runs = []
    for graphed in (False, True):
        opt, data, net, flat, adam = make()
        B, _, H, W = data.image.shape
        image_flat = data.image.reshape(B, 3, -1).permute(0, 2, 1).contiguous()
        body = body_of(opt, data, net, flat, adam, image_flat)
        g = torch.Generator(device="cpu").manual_seed(3)
        idxs = [torch.randperm(H * W, generator=g)[:64].to(dev) for _ in range(8)]
        if graphed:
            snap = (flat.flat_param.clone(), adam.exp_avg.clone(), adam.exp_avg_sq.clone(), adam.steps.clone())
            step = GraphedStep(body, (idxs[0].clone(),), warmup=2)
            flat.flat_param.copy_(snap[0]); adam.exp_avg.copy_(snap[1]); adam.exp_avg_sq.copy_(snap[2]); adam.steps.copy_(snap[3])
            adam.scratch.zero_()
        else:
            step = body
        losses = [float(step(i)) for i in idxs]
        runs.append((losses, flat.flat_param.clone()))
    # Adam's first updates are ~ lr * sign(g): for the many weights whose gradient is pure rounding noise of the atomic
    # accumulation the sign is arbitrary, so two runs of the SAME eager code differ by up to 2 * lr * steps in those
    # weights.  The check is that the replay follows the same trajectory where it is determined: the losses.
    for a, b in zip(*[r[0] for r in runs]):
        assert abs(a - b) <= 1e-2 * max(abs(a), 1e-3)
    d = (runs[0][1] - runs[1][1]).abs()
    assert d.max() <= 2 * 5e-4 * 8 + 1e-6 and (d <= 1e-4).float().mean() > 0.9
