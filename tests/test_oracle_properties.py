"""Size-independent properties of the oracle's restatement (CPU; part of `-m "not gpu"`): the invariants the domain
offers — sortedness of the merged samples, probability bounds of the compositing weights, rotation-group properties of
the pose maths, linearity of the reprojection error in the weights — on seeded random inputs of several shapes.  The
golden fixtures pin values; these pin structure at sizes the fixtures do not cover."""
import math

import pytest
import torch

from oracle import neus_oracle as O


def _rays(B, g):
    o = torch.tensor([0.0, 0.0, -3.0]).repeat(B, 1) + 0.05 * torch.randn(B, 3, generator=g)
    d = torch.nn.functional.normalize(torch.tensor([0.0, 0.0, 1.0]) + 0.15 * torch.randn(B, 3, generator=g), dim=-1)
    near, far = O.near_far_from_sphere(o, d)
    return o, d, near, far


@pytest.mark.parametrize("B,S,n", [(1, 2, 1), (7, 17, 16), (33, 112, 16), (5, 64, 64)])
def test_sample_pdf_stays_sorted_and_inside_the_bins(B, S, n):
    g = torch.Generator().manual_seed(B * 1000 + S)
    bins = torch.sort(torch.rand(B, S, generator=g) * 2.0 + 1.0, dim=-1)[0]
    w = torch.rand(B, S - 1, generator=g) ** 6                      # peaky, with near-empty bins
    w[:, ::3] = 0.0
    z = O.sample_pdf_det(bins, w, n)
    assert z.shape == (B, n)
    assert bool((z[:, 1:] >= z[:, :-1]).all())
    assert bool((z >= bins[:, :1] - 1e-6).all()) and bool((z <= bins[:, -1:] + 1e-6).all())
    # all-equal weights: the deterministic u grid maps to a uniform resample of equally spaced bins
    lin = torch.linspace(1.0, 3.0, S)[None].repeat(B, 1)
    zu = O.sample_pdf_det(lin, torch.ones(B, S - 1), n)
    u = torch.linspace(0.5 / n, 1 - 0.5 / n, n)
    torch.testing.assert_close(zu, (1.0 + 2.0 * u)[None].repeat(B, 1), atol=2e-5, rtol=0)


@pytest.mark.parametrize("B,n,m,steps", [(9, 16, 16, 2), (4, 64, 64, 4), (3, 32, 48, 3)])
def test_hierarchical_sampling_merges_sorted_and_keeps_sdf_aligned(B, n, m, steps):
    g = torch.Generator().manual_seed(n + m)
    o, d, near, far = _rays(B, g)
    sdf_fn = lambda p: (p.norm(dim=-1, keepdim=True) - 0.5)          # analytic sphere instead of the MLP
    z = O.coarse_z(near, far, n, torch.rand(B, 1, generator=g))
    sdf = sdf_fn((o[:, None] + d[:, None] * z[..., None]).reshape(-1, 3)).reshape(B, n)
    for i in range(steps):
        new_z = O.up_sample(o, d, z, sdf, m // steps, 64 * 2 ** i)
        assert new_z.shape == (B, m // steps) and bool((new_z[:, 1:] >= new_z[:, :-1]).all())
        z, sdf = O.cat_z_vals(sdf_fn, o, d, z, new_z, sdf, last=False)
        assert bool((z[:, 1:] >= z[:, :-1]).all())
        ref = sdf_fn((o[:, None] + d[:, None] * z[..., None]).reshape(-1, 3)).reshape(B, -1)
        torch.testing.assert_close(sdf, ref, atol=1e-6, rtol=0)       # the permutation carried the SDF along
    assert z.shape == (B, n + (m // steps) * steps)
    # importance samples concentrate around the surface crossing |x| = 0.5 of rays that hit the sphere
    hit = (o + d * (-(o * d).sum(-1, keepdim=True))).norm(dim=-1) < 0.4
    if hit.any():
        r = (o[:, None] + d[:, None] * z[..., None]).norm(dim=-1)
        near_surface = ((r - 0.5).abs() < 0.1).float().mean(dim=-1)
        assert float(near_surface[hit].mean()) > 0.25


@pytest.mark.parametrize("B,S,car,bg", [(6, 32, 1.0, False), (5, 128, 0.3, True), (2, 3, 0.0, True)])
def test_compositing_weights_are_a_sub_probability_and_colour_is_convex(B, S, car, bg):
    g = torch.Generator().manual_seed(S)
    o, d, near, far = _rays(B, g)
    z = O.coarse_z(near, far, S, torch.rand(B, 1, generator=g))
    pts = (o[:, None] + d[:, None] * z[..., None]).reshape(-1, 3)
    sdf = pts.norm(dim=-1, keepdim=True) - 0.5 + 0.02 * torch.randn(B * S, 1, generator=g)
    grad = torch.nn.functional.normalize(pts, dim=-1) * (1.0 + 0.1 * torch.randn(B * S, 1, generator=g))
    rgb = torch.rand(B, S, 3, generator=g)
    background = torch.ones(1, 3) if bg else None
    out = O.composite(o, d, z, sdf, grad, rgb, torch.tensor(20.0), 2.0 / S, background_rgb=background,
                      cos_anneal_ratio=car)
    w = out["weights"]
    assert bool((w >= 0).all()) and bool((w.sum(-1) <= 1.0 + 1e-4).all())
    assert bool((out["alpha"] >= 0).all()) and bool((out["alpha"] <= 1).all())
    assert bool((out["color"] >= -1e-6).all()) and bool((out["color"] <= 1.0 + 1e-4).all())
    assert bool(((out["inside_sphere"] == 0) | (out["inside_sphere"] == 1)).all())
    assert float(out["gradient_error"]) >= 0.0
    # transmittance form: w_j = alpha_j * prod_{k<j} (1 - alpha_k + 1e-7)
    T = torch.cumprod(torch.cat([torch.ones(B, 1), 1.0 - out["alpha"] + 1e-7], -1), -1)[:, :-1]
    torch.testing.assert_close(w, out["alpha"] * T, atol=1e-7, rtol=0)


def test_pose_maths_stays_in_the_rotation_group():
    g = torch.Generator().manual_seed(4)
    r = torch.cat([torch.randn(16, 3, generator=g), 1e-9 * torch.randn(2, 3, generator=g), torch.zeros(1, 3)])
    R = O.rodrigues_exp(r)
    eye = torch.eye(3)[None].expand_as(R)
    torch.testing.assert_close(R @ R.transpose(1, 2), eye, atol=2e-6, rtol=0)
    torch.testing.assert_close(torch.linalg.det(R), torch.ones(R.shape[0]), atol=2e-6, rtol=0)
    torch.testing.assert_close(O.rodrigues_exp(-r), R.transpose(1, 2), atol=2e-6, rtol=0)
    # rotation angle = |r| (mod 2 pi): trace(R) = 1 + 2 cos|r|
    torch.testing.assert_close(R.diagonal(dim1=1, dim2=2).sum(-1), 1.0 + 2.0 * torch.cos(r.norm(dim=-1)), atol=5e-6, rtol=0)
    wu = torch.cat([torch.zeros(5, 3), torch.randn(5, 3, generator=g)], dim=-1)          # pure translation
    Rt = O.se3_to_SE3(wu)
    torch.testing.assert_close(Rt[:, :, :3], torch.eye(3)[None].expand(5, 3, 3), atol=1e-7, rtol=0)
    torch.testing.assert_close(Rt[:, :, 3], wu[:, 3:], atol=1e-7, rtol=0)
    wu = torch.randn(8, 6, generator=g) * 0.5
    Rt = O.se3_to_SE3(wu)
    torch.testing.assert_close(Rt[:, :, :3] @ Rt[:, :, :3].transpose(1, 2), torch.eye(3)[None].expand(8, 3, 3), atol=5e-6, rtol=0)
    # compose_pair with the identity is the identity map on poses
    ident = torch.cat([torch.eye(3), torch.zeros(3, 1)], dim=1)[None].expand(8, 3, 4)
    torch.testing.assert_close(O.compose_pair(Rt, ident), Rt, atol=1e-7, rtol=0)
    torch.testing.assert_close(O.compose_pair(ident, Rt), Rt, atol=1e-7, rtol=0)


def test_reprojection_error_vanishes_for_consistent_geometry_and_is_linear_in_weights():
    """Points seen from frame A, projected into frame B, land on the pixels that B's own rays through those points
    have: the flow error is zero for a delta weight there, and it is linear in the weights."""
    g = torch.Generator().manual_seed(8)
    K = torch.tensor([[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1.0]])
    pose_a = torch.cat([O.rodrigues_exp(torch.tensor([[0.05, -0.1, 0.02]]))[0], torch.tensor([[0.1], [0.0], [-3.0]])], dim=1)
    pose_b = torch.cat([O.rodrigues_exp(torch.tensor([[-0.04, 0.12, 0.01]]))[0], torch.tensor([[-0.2], [0.1], [-2.9]])], dim=1)
    B, S = 6, 12
    px = torch.rand(B, generator=g) * 200 + 220
    py = torch.rand(B, generator=g) * 200 + 140
    ro, rd = O.gen_rays(pose_a, torch.linalg.inv(K), px, py)
    z = torch.sort(torch.rand(B, S, generator=g) * 2 + 2, dim=-1)[0]
    sd = 0.25
    pts = O.sample_points(ro, rd, z, sd)
    j = 5
    # pixel of sample j in frame B
    w2c = torch.linalg.inv(torch.cat([pose_b, torch.tensor([[0.0, 0, 0, 1]])], 0))[:3]
    cam = pts[:, j] @ w2c[:, :3].T + w2c[:, 3]
    pix = cam @ K.T
    xy = pix[:, :2] / pix[:, 2:]
    w = torch.zeros(B, S)
    w[:, j] = 1.0
    err = O.reprojection_error(pts, w, pose_b, K, xy)
    assert float(err.abs().max()) < 2e-3                      # pixels (fp32 projection of ~600 px coordinates)
    w1, w2 = torch.rand(B, S, generator=g), torch.rand(B, S, generator=g)
    e12 = O.reprojection_error(pts, 2.0 * w1 - 0.5 * w2, pose_b, K, xy)
    e1, e2 = O.reprojection_error(pts, w1, pose_b, K, xy), O.reprojection_error(pts, w2, pose_b, K, xy)
    torch.testing.assert_close(e12, 2.0 * e1 - 0.5 * e2, atol=2e-2, rtol=1e-4)
    # sample_points: last sample uses sample_dist, the others the midpoint
    mid = torch.cat([(z[:, :-1] + z[:, 1:]) * 0.5, z[:, -1:] + sd * 0.5], dim=-1)
    torch.testing.assert_close(pts, ro[:, None] + rd[:, None] * mid[..., None], atol=1e-6, rtol=0)


def test_lr_schedule_shape():
    from fmov_pose_b200.train import LRSchedule
    s = LRSchedule(learning_rate=5e-4, learning_rate_alpha=0.05, warm_up_end=5000, end_iter=300000)
    lrs = [s.net_lr(i) for i in range(0, 300001, 2500)]
    assert lrs[0] == 0.0 and abs(lrs[2] - 5e-4) < 1e-12                    # linear warm-up ends at the base rate
    assert all(a >= b - 1e-15 for a, b in zip(lrs[2:], lrs[3:]))           # then monotone cosine decay
    assert abs(lrs[-1] - 5e-4 * 0.05) < 1e-12                              # down to alpha * lr
    assert math.isclose(s.pose_mlp_lr(0), 5e-4) and math.isclose(s.pose_mlp_lr(1000), 5e-4 * 0.5)
