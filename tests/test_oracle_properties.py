"""Size-independent properties of the CPU oracle (hypothesis): they hold for the reference algorithm by construction and
are the same invariants the full-size GPU parity tests check on the kernels."""
import math

import torch
from hypothesis import given, settings, strategies as st

from oracle import sparf_oracle as O

settings.register_profile("fast", max_examples=25, deadline=None)
settings.load_profile("fast")


@given(st.integers(2, 96), st.integers(1, 4), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_composite_partition_of_unity(S, R, seed, white):
    g = torch.Generator().manual_seed(seed)
    t = torch.sort(torch.rand(1, R, S, generator=g, dtype=torch.float64) * 4 + 0.5, dim=-1).values
    dens = torch.rand(1, R, S, generator=g, dtype=torch.float64) * 5
    rgb = torch.rand(1, R, S, 3, generator=g, dtype=torch.float64)
    ray = torch.randn(1, R, 3, generator=g, dtype=torch.float64)
    out = O.composite(ray, dens, rgb, t, white_bg=white)
    w = out["weights"][..., 0]
    assert (w >= 0).all() and torch.allclose(w.sum(-1, keepdim=True), out["opacity"])
    # the 1e10 last interval makes the ray opaque whenever the last density is positive: opacity + T_end == 1
    assert (out["opacity"] <= 1 + 1e-12).all()
    if white:
        assert torch.allclose(out["rgb"], (w[..., None] * rgb).sum(-2) + (1 - out["opacity"]))
    assert ((out["depth"] >= t.min() * out["opacity"] - 1e-9) & (out["depth"] <= t.max() * out["opacity"] + 1e-9)).all()


@given(st.integers(2, 64), st.integers(1, 64), st.integers(0, 2 ** 31 - 1))
def test_pdf_samples_stay_in_range_and_follow_the_mass(S, S_fine, seed):
    g = torch.Generator().manual_seed(seed)
    w = torch.rand(1, 3, S, generator=g, dtype=torch.float64) ** 4 + 1e-3
    near, far = 1.0, 5.0
    ts = O.sample_pdf(w, S, S_fine, (near, far))
    assert ts.shape == (1, 3, S_fine)
    assert (ts >= near - 1e-9).all() and (ts <= far + 1e-9).all()
    assert (ts[..., 1:] >= ts[..., :-1] - 1e-12).all()          # deterministic grid: monotone in u
    # all the mass in one bin -> every sample falls inside that bin
    k = int(torch.randint(0, S, (1,), generator=g))
    one = torch.zeros(1, 1, S, dtype=torch.float64)
    one[..., k] = 1.0
    ts1 = O.sample_pdf(one, S, S_fine, (near, far))
    lo, hi = near + (far - near) * k / S, near + (far - near) * (k + 1) / S
    assert (ts1 >= lo - 1e-6).all() and (ts1 <= hi + 1e-6).all()


@given(st.floats(0.0, 1.0), st.floats(0.0, 0.5), st.floats(0.55, 1.0))
def test_c2f_weights_are_a_monotone_ramp(progress, start, end):
    w = O.c2f_weights(10, progress, (start, end))
    assert ((w >= 0) & (w <= 1)).all()
    assert (w[:-1] >= w[1:] - 1e-7).all()                          # low bands open first
    w_later = O.c2f_weights(10, min(1.0, progress + 0.1), (start, end))
    assert (w_later >= w - 1e-6).all()                             # and never close again
    if progress <= start:
        assert w.abs().max() <= 1e-6
    if progress >= end:
        assert (w - 1).abs().max() <= 1e-6


@given(st.integers(0, 2 ** 31 - 1))
def test_pose_inversion_and_d9_round_trip(seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(4, 3, 3, generator=g, dtype=torch.float64)
    Q, _ = torch.linalg.qr(A)
    Q = Q * torch.sign(torch.linalg.det(Q))[:, None, None]
    P = torch.cat([Q, torch.randn(4, 3, 1, generator=g, dtype=torch.float64)], dim=-1)
    assert torch.allclose(O.invert_pose(O.invert_pose(P)), P, atol=1e-12)
    assert torch.allclose(O.d9_to_pose(O.pose_to_d9(P)), P, atol=1e-10)


@given(st.integers(2, 64), st.integers(0, 2 ** 31 - 1))
def test_distortion_loss_is_translation_invariant_and_quadratic(S, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.sort(torch.rand(1, 2, S, 1, generator=g, dtype=torch.float64) * 3 + 1, dim=2).values
    w = torch.rand(1, 2, S, 1, generator=g, dtype=torch.float64)
    base = O.distortion_loss(t, w)
    assert torch.allclose(O.distortion_loss(t + 7.5, w), base, rtol=1e-9, atol=1e-12)
    assert torch.allclose(O.distortion_loss(t, 3 * w), 9 * base, rtol=1e-9, atol=1e-12)
    assert base >= 0


# ------------------------------------------------------------------------------------------------ pose algebra (round 2)
@settings(max_examples=40, deadline=None)
@given(st.lists(st.floats(-1.5, 1.5, allow_nan=False, width=32), min_size=6, max_size=6))
def test_se3_exponential_is_rigid_and_composes(v):
    """se3_to_SE3 (camera.py:142-157 restated): R is a rotation, exp(0) = identity, and composing a pose with its inverse
    (compose_pair, camera.py:108-115) gives the identity."""
    wu = torch.tensor(v, dtype=torch.float64)[None]
    P = O.se3_to_SE3(wu)[0]
    R, t = P[:, :3], P[:, 3]
    assert torch.allclose(R @ R.T, torch.eye(3, dtype=torch.float64), atol=1e-9)
    assert abs(float(torch.linalg.det(R)) - 1.0) < 1e-9
    Pinv = O.invert_pose(P[None])[0]
    I = O.compose_pair(P[None], Pinv[None])[0]
    assert torch.allclose(I, torch.cat([torch.eye(3, dtype=torch.float64), torch.zeros(3, 1, dtype=torch.float64)], 1), atol=1e-9)
    Z = O.se3_to_SE3(torch.zeros(1, 6, dtype=torch.float64))[0]
    assert torch.allclose(Z, torch.cat([torch.eye(3, dtype=torch.float64), torch.zeros(3, 1, dtype=torch.float64)], 1))
    # translation part: for a pure translation generator (w = 0) V = I
    wu0 = torch.cat([torch.zeros(3, dtype=torch.float64), torch.tensor(v[3:], dtype=torch.float64)])[None]
    assert torch.allclose(O.se3_to_SE3(wu0)[0][:, 3], wu0[0, 3:], atol=1e-12)
