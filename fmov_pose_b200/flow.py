"""Flow / reprojection loss, unit-sphere loss and a differentiable `pts` on the fused kernels of
csrc/flow.cu.

The reference differentiates the per-sample points `render_out["pts"]` ([P,3], exp_runner.py:607, 715): here the
points are never materialised with a gradient — `fmov_flow_fwd/_bwd` recompute p = o + d*mid_z per ray and reduce the
backward directly to d rays_o, d rays_d (-> pose parameters of the rendered frame), d z (n_importance == 0), d weights
(-> all networks through the compositing backward) and d w2c (-> pose parameters of the matched frame)."""
import torch

from . import _lib as L
from . import ops as _ops


def w2c_from_c2w(c2w34):
    """rows 0..2 of inverse([c2w; 0 0 0 1]) (exp_runner.py:634-636 `torch.inverse(c2w)[:3]`) in closed form
    (adjugate / determinant of the 3x3 block): differentiable, no cuSOLVER call and no host sync, so it is legal inside
    a CUDA-graph capture."""
    R, t = c2w34[:3, :3], c2w34[:3, 3]
    c0, c1, c2 = R[:, 0], R[:, 1], R[:, 2]
    adj = torch.stack([torch.linalg.cross(c1, c2), torch.linalg.cross(c2, c0), torch.linalg.cross(c0, c1)])
    Rinv = adj / torch.dot(c0, torch.linalg.cross(c1, c2))
    return torch.cat([Rinv, -(Rinv @ t)[:, None]], dim=1)


class FlowReprojFunction(torch.autograd.Function):
    """(rays_o [B,3], rays_d [B,3], z [B,S], weights [B,S], w2c34 [3,4], K33 [3,3], xy [B,2], sample_dist)
    -> err [B,2] = sum_j w_j (pi(K (R_w p_j + t_w)) - xy)      (exp_runner.py:637-654 / 670-687)"""

    @staticmethod
    def forward(ctx, rays_o, rays_d, z, weights, w2c34, K33, xy, sample_dist):
        rays_o, rays_d, z, weights = L.f32c(rays_o), L.f32c(rays_d), L.f32c(z), L.f32c(weights)
        w2c34, xy = L.f32c(w2c34), L.f32c(xy)
        K33 = K33.float()
        if K33.stride(1) != 1:
            K33 = K33.contiguous()
        B, S = z.shape
        if not (rays_o.shape == (B, 3) and rays_d.shape == (B, 3) and weights.shape == (B, S) and xy.shape == (B, 2)
                and w2c34.shape == (3, 4) and K33.shape == (3, 3)):
            raise ValueError(f"flow reprojection: inconsistent shapes rays {tuple(rays_o.shape)} z {tuple(z.shape)} "
                             f"weights {tuple(weights.shape)} xy {tuple(xy.shape)}")
        err = torch.empty(B, 2, dtype=torch.float32, device=z.device)
        L.check(L.lib().fmov_flow_fwd(L.c_ll(B), S, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.c_float(sample_dist),
                                      L.ptr(weights), L.ptr(w2c34), L.ptr(K33), K33.stride(0), L.ptr(xy), L.ptr(err),
                                      L.stream()), "fmov_flow_fwd")
        ctx.save_for_backward(rays_o, rays_d, z, weights, w2c34, K33, xy)
        ctx.sample_dist = float(sample_dist)
        return err

    @staticmethod
    def backward(ctx, g_err):
        rays_o, rays_d, z, weights, w2c34, K33, xy = ctx.saved_tensors
        B, S = z.shape
        dev = z.device
        need = ctx.needs_input_grad
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        d_o, d_d, d_w2c = e(B, 3), e(B, 3), e(3, 4)
        d_w = e(B, S) if need[3] else None
        d_z = e(B, S) if need[2] else None
        L.check(L.lib().fmov_flow_bwd(L.c_ll(B), S, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.c_float(ctx.sample_dist),
                                      L.ptr(weights), L.ptr(w2c34), L.ptr(K33), K33.stride(0), L.ptr(xy),
                                      L.ptr(L.f32c(g_err)), L.ptr(d_w), L.ptr(d_o), L.ptr(d_d), L.ptr(d_z), L.ptr(d_w2c),
                                      L.stream()), "fmov_flow_bwd")
        return (d_o if need[0] else None, d_d if need[1] else None, d_z, d_w, d_w2c if need[4] else None, None, None,
                None)


class UnitSphereFunction(torch.autograd.Function):
    """(rays_o, rays_d, z, weights, sample_dist) -> mean |w| over samples with |o + d*mid_z| > 1
    (exp_runner.py:714-724; the mask is detached there, so only `weights` receives a gradient)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, z, weights, sample_dist, group=None):
        rays_o, rays_d, z, weights = L.f32c(rays_o), L.f32c(rays_d), L.f32c(z), L.f32c(weights)
        B, S = z.shape
        part = torch.empty(2, dtype=torch.float32, device=z.device)
        L.check(L.lib().fmov_unit_sphere_fwd_bwd(L.c_ll(B), S, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z),
                                                 L.c_float(sample_dist), L.ptr(weights), L.ptr(None), L.ptr(part),
                                                 L.ptr(None), L.stream()), "fmov_unit_sphere_fwd_bwd")
        if group is not None:          # ray-sharded: the mean runs over the outside samples of ALL ranks
            torch.distributed.all_reduce(part, group=group)
        ctx.save_for_backward(rays_o, rays_d, z, weights, part)
        ctx.sample_dist = float(sample_dist)
        return part[0] / part[1]          # F.l1_loss(weights[outside], 0): nan when nothing is outside, as the reference

    @staticmethod
    def backward(ctx, g):
        rays_o, rays_d, z, weights, part = ctx.saved_tensors
        B, S = z.shape
        d_w = torch.empty(B, S, dtype=torch.float32, device=z.device)
        g_scale = (g / part[1]).reshape(1).float().contiguous()
        L.check(L.lib().fmov_unit_sphere_fwd_bwd(L.c_ll(B), S, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z),
                                                 L.c_float(ctx.sample_dist), L.ptr(weights), L.ptr(g_scale), L.ptr(None),
                                                 L.ptr(d_w), L.stream()), "fmov_unit_sphere_fwd_bwd")
        return None, None, None, d_w, None, None


class PtsFunction(torch.autograd.Function):
    """Gives `render_out["pts"]` its autograd link pts = o + d*mid_z (models/renderer.py:269-272) without a second
    forward: returns the points the compositing kernel wrote, backward = fmov_ray_reduce_bwd."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, z, pts, sample_dist):
        ctx.save_for_backward(L.f32c(rays_d), L.f32c(z))
        ctx.sample_dist = float(sample_dist)
        return pts.view_as(pts)

    @staticmethod
    def backward(ctx, g_pts):
        rays_d, z = ctx.saved_tensors
        B, S = z.shape
        want_dz = ctx.needs_input_grad[2]
        zero3 = torch.zeros(B, 3, dtype=torch.float32, device=z.device)
        zeroS = torch.zeros(B, S, dtype=torch.float32, device=z.device) if want_dz else None
        d_o, d_d, d_z = _ops.ray_reduce_bwd(L.f32c(g_pts.reshape(-1, 3)), None, zero3, zeroS, zeroS, rays_d, z,
                                            ctx.sample_dist, want_dz)
        return d_o, d_d, d_z, None, None


def flow_loss(render_out, rays_o, rays_d, c2w_1, c2w_0, K_1, K_0, pixels_xy, pixels_xy_corr, sample_dist, flow_weight,
              maintain_shape=False, detach_flow_on_sdf=False, detach_ref=False, n_terms=None):
    """exp_runner.py:605-688.  The batch is [rays of frame 0 (= img_id_corr) | rays of frame 1 (= img_id) | ...]:
    the first part's points are projected into frame 1 (c2w_1, K_1) against `pixels_xy`, the second part's into frame 0
    against `pixels_xy_corr`; parts are quarters of the batch with maintain_shape (the second half being the
    additional frame, exp_runner.py:512-548), halves otherwise.  `n_terms` overrides the l1 mean's element count
    (global count when rays are sharded)."""
    z = render_out["z_vals"]
    weights = render_out["weights"]
    if detach_flow_on_sdf:
        weights = weights.detach()
    B = z.shape[0]
    n = B // 4 if maintain_shape else B // 2
    sl0, sl1 = (slice(0, n), slice(n, 2 * n)) if maintain_shape else (slice(0, n), slice(n, B))
    if detach_ref:
        c2w_1, c2w_0 = c2w_1.detach(), c2w_0.detach()
    total = 0.0
    for sl, c2w, K, xy in ((sl0, c2w_1, K_1, pixels_xy), (sl1, c2w_0, K_0, pixels_xy_corr)):
        err = FlowReprojFunction.apply(rays_o[sl], rays_d[sl], z[sl], weights[sl], w2c_from_c2w(c2w[:3, :4]), K[:3, :3],
                                       xy.float(), sample_dist)
        cnt = err.numel() if n_terms is None else n_terms
        total = total + err.abs().sum() / cnt * flow_weight
    return total


def unit_sphere_loss(render_out, rays_o, rays_d, sample_dist, unit_sphere_weight, group=None):
    """exp_runner.py:714-724; `group`: ray-sharded step — (sum, count) are all-reduced so every rank optimises the
    single-GPU objective (mean over the outside samples of the whole batch)"""
    return UnitSphereFunction.apply(rays_o, rays_d, render_out["z_vals"], render_out["weights"],
                                    sample_dist, group) * unit_sphere_weight
