"""GPU parity of the flow / reprojection loss, unit-sphere loss, differentiable `pts` and sub-pixel ray generation
 against fixtures made by executing the reference's own exp_runner.py lines (605-688, 714-724)
and against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import neus_oracle as O
from tests._util import load_golden, t

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


@pytest.mark.parametrize("name", ["flow_half", "flow_quarter_detach"])
def test_flow_and_unit_sphere_kernels_vs_reference_lines(name):
    from fmov_pose_b200 import flow
    d = load_golden(name)
    lv = {k: t(d, k).to(DEV).requires_grad_(True) for k in ("rays_o", "rays_d", "z", "weights", "c2w_0", "c2w_1")}
    K = t(d, "intrinsics").to(DEV)
    sd = float(d["sample_dist"])
    out = {"z_vals": lv["z"], "weights": lv["weights"]}
    fl = flow.flow_loss(out, lv["rays_o"], lv["rays_d"], lv["c2w_1"], lv["c2w_0"], K[1], K[0], t(d, "pixels_xy").to(DEV),
                        t(d, "pixels_xy_corr").to(DEV), sd, float(d["flow_weight"]),
                        maintain_shape=bool(d["maintain_shape"]), detach_flow_on_sdf=bool(d["detach_flow_on_sdf"]))
    assert abs(float(fl.detach()) - float(d["flow_loss"])) <= 1e-4 * abs(float(d["flow_loss"]))      # fp32 sums
    g = torch.autograd.grad(fl, list(lv.values()), allow_unused=True)
    for (k, v), gi in zip(lv.items(), g):
        ref = d["gflow_" + k]
        got = np.zeros_like(ref) if gi is None else gi.cpu().numpy()
        assert rel(got, ref) <= 1e-3 if np.abs(ref).max() > 0 else np.abs(got).max() == 0, (k, rel(got, ref))
    ul = flow.unit_sphere_loss(out, lv["rays_o"], lv["rays_d"], sd, float(d["unit_sphere_weight"]))
    np.testing.assert_allclose(float(ul.detach()), float(d["unit_sphere_loss"]), rtol=1e-5)
    gu, = torch.autograd.grad(ul, [lv["weights"]])
    np.testing.assert_allclose(gu.cpu().numpy(), d["gunit_weights"], rtol=1e-5, atol=1e-9)


def test_w2c_closed_form_equals_matrix_inverse():
    from fmov_pose_b200 import flow
    torch.manual_seed(3)
    c2w = torch.eye(4, device=DEV)
    c2w[:3, :3] = O.rodrigues_exp(torch.tensor([[0.4, -0.3, 0.2]]))[0].to(DEV) * 1.01     # not exactly orthonormal
    c2w[:3, 3] = torch.tensor([0.3, -0.1, -2.5], device=DEV)
    np.testing.assert_allclose(flow.w2c_from_c2w(c2w[:3]).cpu().numpy(), torch.linalg.inv(c2w.cpu())[:3].numpy(), atol=2e-6)


def test_raygen_with_subpixel_coordinates():
    from fmov_pose_b200.models.dataset import _RayGenFn
    g = torch.Generator().manual_seed(2)
    B = 333
    px = torch.rand(B, generator=g) * 639
    py = torch.rand(B, generator=g) * 479
    intr_inv = torch.linalg.inv(torch.tensor([[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1.0]])).contiguous()
    pose = torch.eye(4)[:3].clone()
    pose[:3, :3] = O.rodrigues_exp(torch.tensor([[0.1, -0.2, 0.05]]))[0]
    pose[:, 3] = torch.tensor([0.1, -0.2, -3.0])
    pose_c = pose.clone().requires_grad_(True)
    ro_ref, rd_ref = O.gen_rays(pose_c, intr_inv, px, py)
    Go, Gd = torch.randn(B, 3, generator=g), torch.randn(B, 3, generator=g)
    ((ro_ref * Go).sum() + (rd_ref * Gd).sum()).backward()
    pose_g = pose.to(DEV).requires_grad_(True)
    ro, rd = _RayGenFn.apply(pose_g, intr_inv.to(DEV), px.to(DEV), py.to(DEV))
    np.testing.assert_allclose(rd.detach().cpu().numpy(), rd_ref.detach().numpy(), atol=2e-6)
    np.testing.assert_allclose(ro.detach().cpu().numpy(), ro_ref.detach().numpy(), atol=0)
    ((ro * Go.to(DEV)).sum() + (rd * Gd.to(DEV)).sum()).backward()
    assert rel(pose_g.grad.cpu().numpy(), pose_c.grad.numpy()) <= 1e-4


def _scene_with_flows(n_samples, n_importance, seed=5):
    from fmov_pose_b200 import synthetic
    sc = synthetic.build_scene(device=DEV, n_images=6, n_samples=n_samples, n_importance=n_importance,
                               up_sample_steps=2, pose_type="seg")
    with torch.no_grad():          # move the per-frame poses apart so that the reprojection is not an identity
        for i, m in enumerate(sc["pose_network"].pose_mlps):
            for p in m.parameters():
                if p.requires_grad:
                    p.add_(torch.randn(p.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(seed + i)) * 0.02)
    rng = np.random.RandomState(seed)
    n = 400
    xa, ya = rng.uniform(200, 440, n), rng.uniform(120, 360, n)
    matches = {(2, 3): (xa, ya, xa + rng.normal(0, 4, n), ya + rng.normal(0, 4, n)),
               (2, 4): (xa, ya, xa + rng.normal(0, 6, n), ya + rng.normal(0, 6, n))}
    sc["dataset"].set_flows(matches)
    return sc


def test_pair_sampler_contract():
    sc = _scene_with_flows(16, 0)
    ds, pn = sc["dataset"], sc["pose_network"]
    assert ds.gen_random_ray_pairs_at(0, 8, pn, 6)[0] is None                  # frame without matches
    assert ds.gen_random_ray_pairs_at(2, 8, pn, 3, interval=10)[0] is None     # matched frames not yet registered
    assert ds.gen_random_ray_pairs_at(2, 8, pn, 6, interval=0)[0] is None      # outside the flow interval
    idx = np.arange(8)
    data, xy, xy_corr, img_id, depth = ds.gen_random_ray_pairs_at(2, 8, pn, 6, interval=1, indexs=idx)
    assert int(img_id) == 3 and data.shape == (16, 10) and xy.shape == (8, 2) and depth.shape == (16,)
    xs1, ys1, xs2, ys2 = ds.loftr_interval_flows["0002_0003"]
    np.testing.assert_array_equal(xy_corr.cpu().numpy(), np.stack([xs1[idx], ys1[idx]], -1))
    np.testing.assert_array_equal(xy.cpu().numpy(), np.stack([xs2[idx], ys2[idx]], -1))
    # first half: rays of img_id_corr (frame 2) through its own pose; colours at the truncated pixel; mask = 1
    ro_ref, rd_ref = O.gen_rays(pn(2)[:3].detach().cpu(), ds.intrinsics_all_inv[2].cpu(), torch.from_numpy(xs1[idx]),
                                torch.from_numpy(ys1[idx]))
    np.testing.assert_allclose(data[:8, 3:6].detach().cpu().numpy(), rd_ref.numpy(), atol=3e-6)
    np.testing.assert_allclose(data[:8, 0:3].detach().cpu().numpy(), ro_ref.numpy(), atol=1e-6)
    col = ds.images[3][torch.from_numpy(ys2[idx]).long(), torch.from_numpy(xs2[idx]).long()]
    np.testing.assert_array_equal(data[8:, 6:9].detach().cpu().numpy(), col.cpu().numpy())
    assert float(data[:, 9].detach().min()) == 1.0


@pytest.mark.parametrize("cfg", [dict(n_samples=16, n_importance=16, maintain_shape=True),
                                 dict(n_samples=32, n_importance=0, maintain_shape=False)])
def test_flow_iteration_fused_kernels_equal_autograd_on_pts(cfg):
    """A whole `use_flow` iteration: the fused reprojection kernels (no [P,3] gradient tensor) give the same loss and the
    same gradients on every network and pose parameter as the reference's formulation — torch ops on the differentiable
    `render_out["pts"]` (exp_runner.py:605-688 transcribed below) — on the same render."""
    from fmov_pose_b200.train import TrainStep
    B = 128
    g = torch.Generator().manual_seed(9)
    idx = torch.randint(0, 400, [B // 2], generator=g).numpy()
    add_px = (torch.randint(200, 440, [B], generator=g).to(DEV), torch.randint(120, 360, [B], generator=g).to(DEV))
    n_rays = B + (B if cfg["maintain_shape"] else 0)
    tr = torch.rand(n_rays, 1, generator=g).to(DEV)
    res = []
    for fused in (True, False):
        sc = _scene_with_flows(cfg["n_samples"], cfg["n_importance"])
        ts = TrainStep(sc, mask_weight=5.0, optimizer=False, flow_weight=0.1, unit_sphere_weight=0.05,
                       maintain_shape=cfg["maintain_shape"])
        kw = dict(img_id=3, indexs=idx, additional_img_id=1 if cfg["maintain_shape"] else None,
                  add_pixels=add_px if cfg["maintain_shape"] else None, t_rand=tr)
        if fused:
            ls, out, img_id = ts.flow_forward_backward(2, B, 6, interval=1, **kw)
            assert int(img_id) == 3
        else:
            ds, rend, pn = sc["dataset"], sc["renderer"], sc["pose_network"]
            data, xy, xy_corr, img_id, _ = ds.gen_random_ray_pairs_at(2, B // 2, pn, 6, 1, img_id=3, indexs=idx)
            if cfg["maintain_shape"]:
                add, _ = ds.gen_random_rays_at(1, B, pn(1)[:3], pixels=add_px)
                data = torch.cat([data, add], 0)
            rays_o, rays_d = data[:, :3], data[:, 3:6]
            near, far = ds.near_far_from_sphere(rays_o, rays_d)
            out = rend.render(rays_o, rays_d, near, far, cos_anneal_ratio=1.0, t_rand=tr)
            ls = ts.losses(out, data[:, 6:9], data[:, 9:10])
            pts, w = out["pts"], out["weights"]
            S = w.shape[1]
            n = n_rays // 4 if cfg["maintain_shape"] else n_rays // 2
            parts = ((pts[: n * S], w[:n], 3, xy), (pts[n * S: 2 * n * S], w[n: 2 * n], 2, xy_corr))
            fl = 0.0
            for p_, w_, frame, target in parts:
                c2w = torch.eye(4, device=DEV)
                c2w = torch.cat([pn(frame)[:3], c2w[3:]], 0)
                w2c = torch.inverse(c2w)[:3]
                cam = p_ @ w2c[:, :3].T + w2c[:, 3]
                pix = cam @ ds.intrinsics_all[frame][:3, :3].T
                uv = (pix[:, :2] / pix[:, 2:]).reshape(-1, S, 2)
                err = ((uv - target[:, None, :]) * w_[:, :, None]).sum(1)
                fl = fl + err.abs().mean() * 0.1
            outside = (pts.norm(dim=-1) > 1.0).detach()
            ul = w.reshape(-1)[outside].abs().mean() * 0.05
            ls["flow_loss"], ls["unit_sphere_loss"] = fl, ul
            ls["loss"] = ls["loss"] + fl + ul
            ls["loss"].backward()
        res.append((ls, [None if p.grad is None else p.grad.detach().clone() for p in ts.all_params],
                    [n_ for n_, _ in list(sc["sdf_network"].named_parameters())]))
    (la, ga, _), (lb, gb, _) = res
    for k in ("loss", "flow_loss", "unit_sphere_loss", "color_loss"):
        np.testing.assert_allclose(float(la[k].detach()), float(lb[k].detach()), rtol=2e-4, err_msg=k)
    assert float(la["flow_loss"].detach()) > 0 and float(la["unit_sphere_loss"].detach()) > 0
    n_checked = 0
    for a, b in zip(ga, gb):
        assert (a is None) == (b is None)
        if a is not None and float(b.abs().max()) > 0:
            assert rel(a.cpu().numpy(), b.cpu().numpy()) <= 2e-3
            n_checked += 1
    assert n_checked > 40
