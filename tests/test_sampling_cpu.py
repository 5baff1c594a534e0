"""RaySamplingStrategy / sample_rays mirrors (sparf_b200/sampling_strategies.py) against the reference's own classes
(oracle/ref_loader.py): same pools, same torch.randperm draws in the same order => identical rays under one seed."""
import pytest
import torch

import common


def _opt(**kw):
    opt = common.make_opt(rand_rays=96)
    opt.sample_fraction_in_fg_mask = 0.0
    opt.sampled_fraction_in_center = 0.0
    opt.depth_regu_patch_size = 2
    for k, v in kw.items():
        if k == "depth_patch":
            opt.loss_weight.depth_patch = v
        else:
            opt[k] = v
    return opt


@pytest.mark.parametrize("case", [dict(), dict(sampled_fraction_in_center=0.25), dict(depth_patch=0),
                                  dict(depth_patch=0, sampled_fraction_in_center=0.5)])
def test_ray_sampling_strategy_matches_reference(case):
    from oracle import ref_loader
    if not ref_loader.ref_root():
        pytest.skip("reference not available")
    ref = ref_loader.load("trainer")
    try:
        from sparf_b200.sampling_strategies import RaySamplingStrategy, sample_rays
        opt = _opt(**case)
        data = common.make_scene(3, 3, 24, 32)
        dev = torch.device("cpu")
        ours = RaySamplingStrategy(opt, data, dev)
        theirs = ref.sampling.RaySamplingStrategy(opt, data_dict=data, device=dev)
        for center in (False, True):
            torch.manual_seed(5)
            a = ours(96, sample_in_center=center)
            torch.manual_seed(5)
            b = theirs(96, sample_in_center=center)
            assert a.shape == b.shape and torch.equal(a, b)
        for kw in (dict(nbr=50), dict(nbr=64, fraction_in_center=0.25), dict()):
            torch.manual_seed(9)
            pa, ra = sample_rays(24, 32, **kw)
            torch.manual_seed(9)
            pb, rb = ref.sampling.sample_rays(24, 32, **kw)
            assert torch.equal(pa, pb) and torch.equal(ra, rb)
    finally:
        ref_loader._purge()
