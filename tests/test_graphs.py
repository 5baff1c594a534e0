"""CUDA-graph replay of a training step (sparf_b200.graphs) reproduces the eager step."""
import pytest
import torch

import common
from sparf_b200 import _lib, ops


@pytest.mark.gpu
def test_graphed_mlp_step_matches_eager():
    from sparf_b200.graphs import GraphedStep
    R, S = 96, 64
    opt = common.make_opt(S=S)
    sd = common.det_weights(opt, 0)
    keys = sum([["mlp_feat.%d.weight" % i, "mlp_feat.%d.bias" % i] for i in range(8)], []) + \
        ["mlp_rgb.0.weight", "mlp_rgb.0.bias", "mlp_rgb.1.weight", "mlp_rgb.1.bias"]
    params = [sd[k].cuda().requires_grad_(True) for k in keys]
    for p in params:
        p.grad = torch.zeros_like(p)
    spec = ops.MLPSpec()
    g = torch.Generator(device="cuda").manual_seed(0)
    d = torch.nn.functional.normalize(torch.randn(R, 3, device="cuda", generator=g), dim=-1)
    t = torch.sort(torch.rand(R, S, device="cuda", generator=g) * 4 + 1.2, dim=1).values
    gs, gc = torch.randn(R, S, device="cuda", generator=g), torch.randn(R, S, 3, device="cuda", generator=g)

    def step(o):
        for p in params:
            p.grad.zero_()
        s, c = ops.mlp_forward(spec, o, d, t, params, engine=_lib.ENGINE_TC_3X)
        torch.autograd.backward([s, c], [gs, gc])
        return s.detach(), c.detach()

    o_static = torch.zeros(R, 3, device="cuda")
    graphed = GraphedStep(step, (o_static,))
    for seed in (1, 2):
        o = torch.randn(R, 3, device="cuda", generator=g) * 0.3
        s_g, c_g = [x.clone() for x in graphed(o)]
        grads_g = [p.grad.clone() for p in params]
        s_e, c_e = step(o)
        assert torch.equal(s_g, s_e) and torch.equal(c_g, c_e)
        for a, b in zip(grads_g, [p.grad for p in params]):
            assert (a - b).abs().max() <= 1e-5 * b.abs().max().clamp_min(1e-20) + 1e-12
