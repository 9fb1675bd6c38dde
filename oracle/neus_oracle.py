"""CPU oracle for the FMOV NeuS train-step hot path.  TEST INFRASTRUCTURE ONLY.

A plain, functional PyTorch (fp32 or fp64, CPU) restatement of the reference algorithm for
the path named in BASELINE.json `north_star` / SURVEY.md §8.  It is the checker for the CUDA
path, never the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  Nothing under `fmov_pose_b200/`
imports this module.

Every function cites the reference file:line (relative to /root/reference) it follows.
Parameters are passed as plain dicts keyed exactly like the reference modules'
`state_dict()` (`lin{l}.weight_g`, `lin{l}.weight_v`, `lin{l}.bias`, `variance`), so a
state dict of either the reference modules or of the B200 modules can be fed in.

Pinning: the reference ships no tests and no golden vectors.  The oracle is
pinned against outputs of the *imported reference itself* (run in the build container by
`oracle/gen_golden.py`, fixtures committed under `tests/golden/`): see
`tests/test_oracle_golden.py`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------
# Positional encoding  (models/embedder.py:12-37, :40-55; models/barf_embedder.py:39-56 —
# the BARF variant computes coarse-to-fine weights but never applies them, so it is the
# same function)
# ----------------------------------------------------------------------------------------
def embed(x: torch.Tensor, multires: int) -> torch.Tensor:
    out = [x]
    for k in range(multires):
        f = 2.0 ** k
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


# ----------------------------------------------------------------------------------------
# weight-norm effective weight (nn.utils.weight_norm, dim=0; models/fields.py:81-82)
# ----------------------------------------------------------------------------------------
def eff_weight(p: Params, prefix: str, l: int) -> torch.Tensor:
    key = f"{prefix}lin{l}."
    if key + "weight_v" in p:
        v = p[key + "weight_v"]
        g = p[key + "weight_g"]
        return v * (g / v.norm(dim=1, keepdim=True))  # same op order as torch._weight_norm
    return p[key + "weight"]


def n_lin(p: Params, prefix: str = "") -> int:
    n = 0
    while f"{prefix}lin{n}.bias" in p:
        n += 1
    return n


def softplus100(x: torch.Tensor) -> torch.Tensor:
    # nn.Softplus(beta=100), default threshold=20 (models/fields.py:86)
    return F.softplus(x, beta=100.0, threshold=20.0)


# ----------------------------------------------------------------------------------------
# SDFNetwork.forward / .sdf / .gradient   (models/fields.py:88-124,
# models/barf_fields.py:99-138)
# ----------------------------------------------------------------------------------------
def sdf_forward(p: Params, x: torch.Tensor, multires: int = 6, skip_in=(4,), scale: float = 1.0,
                prefix: str = "") -> torch.Tensor:
    nl = n_lin(p, prefix)
    inputs = x * scale
    if multires > 0:
        inputs = embed(inputs, multires)
    h = inputs
    for l in range(nl):
        if l in skip_in:
            h = torch.cat([h, inputs], 1) / math.sqrt(2.0)
        h = F.linear(h, eff_weight(p, prefix, l), p[f"{prefix}lin{l}.bias"])
        if l < nl - 1:
            h = softplus100(h)
    return torch.cat([h[:, :1] / scale, h[:, 1:]], dim=-1)


def sdf_value(p: Params, x: torch.Tensor, **kw) -> torch.Tensor:
    return sdf_forward(p, x, **kw)[:, :1]


def sdf_gradient(p: Params, x: torch.Tensor, create_graph: bool = True, **kw) -> torch.Tensor:
    """models/fields.py:112-124 — second forward + autograd.grad(create_graph=True). [N,3]"""
    with torch.enable_grad():
        if not x.requires_grad:
            x = x.detach().requires_grad_(True)
        y = sdf_value(p, x, **kw)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=create_graph,
                                   retain_graph=True, only_inputs=True)
    return g


# ----------------------------------------------------------------------------------------
# RenderingNetwork.forward, mode "idr"  (models/fields.py:166-193)
# ----------------------------------------------------------------------------------------
def color_forward(p: Params, pts, normals, dirs, feat, multires_view: int = 4, mode: str = "idr",
                  squeeze_out: bool = True, prefix: str = "") -> torch.Tensor:
    nl = n_lin(p, prefix)
    if multires_view > 0:
        dirs = embed(dirs, multires_view)
    if mode == "idr":
        h = torch.cat([pts, dirs, normals, feat], -1)
    elif mode == "no_view_dir":
        h = torch.cat([pts, normals, feat], -1)
    elif mode == "no_normal":
        h = torch.cat([pts, dirs, feat], -1)
    else:
        raise ValueError(mode)
    for l in range(nl):
        h = F.linear(h, eff_weight(p, prefix, l), p[f"{prefix}lin{l}.bias"])
        if l < nl - 1:
            h = torch.relu(h)
    if squeeze_out:
        h = torch.sigmoid(h)
    return h


def inv_s_from_variance(variance: torch.Tensor) -> torch.Tensor:
    # SingleVarianceNetwork.forward (models/fields.py:293-294) + clip (models/renderer.py:290-292)
    return torch.exp(variance * 10.0).clip(1e-6, 1e6)


# ----------------------------------------------------------------------------------------
# sample_pdf  (models/renderer.py:54-86), det=True branch (the only one the path uses, :219)
# ----------------------------------------------------------------------------------------
def sample_pdf_det(bins: torch.Tensor, weights: torch.Tensor, n_samples: int) -> torch.Tensor:
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0.0 + 0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples,
                       dtype=bins.dtype, device=bins.device)
    u = u.expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b = torch.gather(cdf, 1, below)
    cdf_a = torch.gather(cdf, 1, above)
    bins_b = torch.gather(bins, 1, below)
    bins_a = torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bins_b + t * (bins_a - bins_b)


# ----------------------------------------------------------------------------------------
# NeuSRenderer.up_sample  (models/renderer.py:168-220)
# ----------------------------------------------------------------------------------------
def up_sample(rays_o, rays_d, z_vals, sdf, n_importance: int, inv_s: float) -> torch.Tensor:
    B, n = z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
    radius = torch.linalg.norm(pts, ord=2, dim=-1)
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    sdf = sdf.reshape(B, n)
    prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
    prev_z, next_z = z_vals[:, :-1], z_vals[:, 1:]
    mid_sdf = (prev_sdf + next_sdf) * 0.5
    cos_val = (next_sdf - prev_sdf) / (next_z - prev_z + 1e-5)
    prev_cos = torch.cat([torch.zeros([B, 1], dtype=z_vals.dtype), cos_val[:, :-1]], dim=-1)
    cos_val = torch.minimum(prev_cos, cos_val)
    cos_val = cos_val.clip(-1e3, 0.0) * inside
    dist = next_z - prev_z
    prev_esti = mid_sdf - cos_val * dist * 0.5
    next_esti = mid_sdf + cos_val * dist * 0.5
    prev_cdf = torch.sigmoid(prev_esti * inv_s)
    next_cdf = torch.sigmoid(next_esti * inv_s)
    alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
    weights = alpha * torch.cumprod(
        torch.cat([torch.ones([B, 1], dtype=z_vals.dtype), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    return sample_pdf_det(z_vals, weights, n_importance).detach()


# ----------------------------------------------------------------------------------------
# NeuSRenderer.cat_z_vals  (models/renderer.py:222-242)
# ----------------------------------------------------------------------------------------
def cat_z_vals(sdf_fn, rays_o, rays_d, z_vals, new_z, sdf, last: bool):
    B, n = z_vals.shape
    m = new_z.shape[1]
    pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z[..., :, None]
    z_cat = torch.cat([z_vals, new_z], dim=-1)
    z_sorted, index = torch.sort(z_cat, dim=-1)
    if not last:
        new_sdf = sdf_fn(pts.reshape(-1, 3)).reshape(B, m)
        sdf = torch.gather(torch.cat([sdf, new_sdf], dim=-1), 1, index)
    return z_sorted, sdf


# ----------------------------------------------------------------------------------------
# NeuSRenderer.render_core  (models/renderer.py:244-372), n_outside == 0 branch
# ----------------------------------------------------------------------------------------
def render_core(sdf_p: Params, col_p: Params, variance, rays_o, rays_d, z_vals, sample_dist: float,
                background_rgb=None, cos_anneal_ratio: float = 0.0, eval: bool = False,
                multires: int = 6, multires_view: int = 4, scale: float = 1.0, skip_in=(4,)):
    B, S = z_vals.shape
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
    mid_z = z_vals + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(B, S, 3).reshape(-1, 3)

    kw = dict(multires=multires, skip_in=skip_in, scale=scale)
    out = sdf_forward(sdf_p, pts, **kw)
    sdf = out[:, :1]
    feat = out[:, 1:]
    gradients = sdf_gradient(sdf_p, pts, **kw)
    if eval:
        gradients = gradients.detach()
    sampled_color = color_forward(col_p, pts, gradients, dirs, feat,
                                  multires_view=multires_view).reshape(B, S, 3)

    ret = composite(rays_o, rays_d, z_vals, sdf, gradients, sampled_color, inv_s_from_variance(variance),
                    sample_dist, background_rgb=background_rgb, cos_anneal_ratio=cos_anneal_ratio)
    ret["sdf"] = sdf
    return ret


def composite(rays_o, rays_d, z_vals, sdf, gradients, sampled_color, inv_s_scalar, sample_dist,
              background_rgb=None, cos_anneal_ratio=0.0):
    """The per-sample -> per-ray tail of render_core (models/renderer.py:261-272, 290-372) given the
    network outputs: sdf [P,1], gradients [P,3], sampled_color [B,S,3], inv_s scalar (clipped)."""
    B, S = z_vals.shape
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
    mid_z = z_vals + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(B, S, 3).reshape(-1, 3)
    inv_s = inv_s_scalar.reshape(1, 1).expand(B * S, 1)
    true_cos = (dirs * gradients).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio)
                 + F.relu(-true_cos) * cos_anneal_ratio)
    est_next = sdf + iter_cos * dists.reshape(-1, 1) * 0.5
    est_prev = sdf - iter_cos * dists.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    p_ = prev_cdf - next_cdf
    c_ = prev_cdf
    alpha = ((p_ + 1e-5) / (c_ + 1e-5)).reshape(B, S).clip(0.0, 1.0)

    pts_norm = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).reshape(B, S)
    inside_sphere = (pts_norm < 1.0).to(z_vals.dtype).detach()
    relax_inside = (pts_norm < 1.2).to(z_vals.dtype).detach()

    weights = alpha * torch.cumprod(
        torch.cat([torch.ones([B, 1], dtype=z_vals.dtype), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    weights_sum = weights.sum(dim=-1, keepdim=True)
    color = (sampled_color * weights[:, :, None]).sum(dim=1)
    if background_rgb is not None:
        color = color + background_rgb * (1.0 - weights_sum)

    g = gradients.reshape(B, S, 3)
    gradient_error = (torch.linalg.norm(g, ord=2, dim=-1) - 1.0) ** 2
    gradient_error = (relax_inside * gradient_error).sum() / (relax_inside.sum() + 1e-5)
    return {
        "color": color, "dists": dists, "gradients": g, "s_val": 1.0 / inv_s,
        "mid_z_vals": mid_z, "weights": weights, "cdf": c_.reshape(B, S),
        "gradient_error": gradient_error, "inside_sphere": inside_sphere, "pts": pts,
        "sampled_color": sampled_color, "alpha": alpha, "relax_sum": relax_inside.sum(),
    }


# ----------------------------------------------------------------------------------------
# NeuSRenderer.render  (models/renderer.py:374-498), n_outside == 0
# `t_rand` is the host-drawn jitter `torch.rand([B,1])` (renderer.py:404) passed in so the
# CUDA path and the oracle consume the same draw (SURVEY.md §7 "RNG parity").
# ----------------------------------------------------------------------------------------
def coarse_z(near, far, n_samples: int, t_rand: Optional[torch.Tensor]):
    z = torch.linspace(0.0, 1.0, n_samples, dtype=near.dtype)
    z = near + (far - near) * z[None, :]
    if t_rand is not None:
        z = z + (t_rand - 0.5) * 2.0 / n_samples
    return z


def sample_z(sdf_p: Params, rays_o, rays_d, near, far, n_samples, n_importance, up_sample_steps,
             t_rand, multires=6, scale=1.0, skip_in=(4,)):
    """renderer.py:385-446: coarse z + hierarchical importance pass (under no_grad)."""
    z_vals = coarse_z(near, far, n_samples, t_rand)
    if n_importance > 0:
        kw = dict(multires=multires, skip_in=skip_in, scale=scale)
        with torch.no_grad():
            B = rays_o.shape[0]
            pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
            sdf = sdf_value(sdf_p, pts.reshape(-1, 3), **kw).reshape(B, n_samples)
            for i in range(up_sample_steps):
                new_z = up_sample(rays_o, rays_d, z_vals, sdf, n_importance // up_sample_steps,
                                  64 * 2 ** i)
                z_vals, sdf = cat_z_vals(lambda q: sdf_value(sdf_p, q, **kw), rays_o, rays_d,
                                         z_vals, new_z, sdf, last=(i + 1 == up_sample_steps))
    return z_vals


def render(sdf_p: Params, col_p: Params, variance, rays_o, rays_d, near, far, *, n_samples=64,
           n_importance=64, up_sample_steps=4, t_rand=None, background_rgb=None,
           cos_anneal_ratio=0.0, eval=False, multires=6, multires_view=4, scale=1.0,
           skip_in=(4,), z_vals=None):
    B = rays_o.shape[0]
    sample_dist = 2.0 / n_samples
    if z_vals is None:
        z_vals = sample_z(sdf_p, rays_o, rays_d, near, far, n_samples, n_importance,
                          up_sample_steps, t_rand, multires, scale, skip_in)
    S = z_vals.shape[1]
    ret = render_core(sdf_p, col_p, variance, rays_o, rays_d, z_vals, sample_dist,
                      background_rgb=background_rgb, cos_anneal_ratio=cos_anneal_ratio,
                      eval=eval, multires=multires, multires_view=multires_view, scale=scale,
                      skip_in=skip_in)
    weights = ret["weights"]
    return {
        "color_fine": ret["color"],
        "depth_fine": (weights * ret["mid_z_vals"]).sum(dim=-1, keepdim=True),
        "s_val": ret["s_val"].reshape(B, S).mean(dim=-1, keepdim=True),
        "cdf_fine": ret["cdf"],
        "weight_sum": weights.sum(dim=-1, keepdim=True),
        "weight_max": torch.max(weights, dim=-1, keepdim=True)[0],
        "gradients": ret["gradients"],
        "weights": weights,
        "gradient_error": ret["gradient_error"],
        "inside_sphere": ret["inside_sphere"],
        "pts": ret["pts"],
        # extras (not in the reference dict) used by stage-wise parity tests
        "z_vals": z_vals, "sdf": ret["sdf"], "sampled_color": ret["sampled_color"],
        "mid_z_vals": ret["mid_z_vals"], "dists": ret["dists"], "alpha": ret["alpha"],
    }


# ----------------------------------------------------------------------------------------
# Loss block  (exp_runner.py:562-599, combined :772-779; flow/depth/unit-sphere terms off)
# ----------------------------------------------------------------------------------------
def loss_block(out, true_rgb, mask, igr_weight: float = 0.1, mask_weight: float = 0.0):
    if mask_weight > 0.0:
        mask = (mask > 0.5).to(true_rgb.dtype)
    else:
        mask = torch.ones_like(mask)
    mask_sum = mask.sum() + 1e-5
    color_error = (out["color_fine"] - true_rgb) * mask
    color_loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum
    eik = out["gradient_error"]
    mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    loss = color_loss + eik * igr_weight + mask_loss * mask_weight
    return {"loss": loss, "color_loss": color_loss, "eikonal_loss": eik, "mask_loss": mask_loss}


# ----------------------------------------------------------------------------------------
# Pose layer
# ----------------------------------------------------------------------------------------
def vec2skew(v):
    # models/batch_lie_group_helper.py:6-16
    zero = torch.zeros(v.shape[0], 1, dtype=v.dtype)
    r0 = torch.cat([zero, -v[:, 2:3], v[:, 1:2]], dim=-1)
    r1 = torch.cat([v[:, 2:3], zero, -v[:, 0:1]], dim=-1)
    r2 = torch.cat([-v[:, 1:2], v[:, 0:1], zero], dim=-1)
    return torch.stack([r0, r1, r2], dim=1)


def rodrigues_exp(r):
    # models/batch_lie_group_helper.py:19-35  (theta = |r| + 1e-15)
    K = vec2skew(r)
    th = r.norm(dim=1, keepdim=True) + 1e-15
    eye = torch.eye(3, dtype=r.dtype).unsqueeze(0).repeat(r.shape[0], 1, 1)
    return eye + (torch.sin(th) / th)[..., None] * K + ((1 - torch.cos(th)) / th ** 2)[..., None] * (K @ K)


def pose_gf_compose(rot, trans, init_c2w34, scale=None):
    """LearnPoseGF.forward tail (models/picture_pose.py:176-186): c2w = [Exp(rot)|trans] @ init,
    init translation optionally scaled (emphasize_rot, :178-179).  rot,trans [3]; init [3,4]."""
    R = rodrigues_exp(rot[None])[0]
    t0 = init_c2w34[:3, 3]
    if scale is not None:
        t0 = t0 * scale
    R0 = init_c2w34[:3, :3]
    Rn = R @ R0
    tn = R @ t0 + trans
    return torch.cat([Rn, tn[:, None]], dim=1)


def pose_gf_mlp(p: Params, cam_id: int, emphasize_rot: bool, small_rot: bool = False):
    """LearnPoseGF.forward head (models/picture_pose.py:140-175): Gaussian-Fourier features of
    the camera index -> 2 GELU layers -> rot*pi (or pi/6), trans, (scale)."""
    b = p["b"]  # [128,1]
    cid = torch.tensor([[float(cam_id)]], dtype=b.dtype)
    a_norm = math.sqrt(b.shape[0])
    ff = torch.cat([torch.sin((2.0 * math.pi * cid) @ b.T), torch.cos((2.0 * math.pi * cid) @ b.T)],
                   dim=-1) / a_norm
    h = F.gelu(F.linear(ff, p["lin1.weight"], p["lin1.bias"]))
    h = F.gelu(F.linear(h, p["lin2.weight"], p["lin2.bias"]))
    k = math.pi / 6 if small_rot else math.pi
    if not emphasize_rot:
        o = F.linear(h, p["lin3.weight"], p["lin3.bias"])
        return o[0, :3] * k, o[0, 3:], None
    rot = F.linear(h, p["lin3_rot.weight"], p["lin3_rot.bias"])[0] * k
    trans = F.linear(h, p["lin3_trans.weight"], p["lin3_trans.bias"])[0]
    scale = F.linear(h, p["lin3_scale.weight"], p["lin3_scale.bias"])[0]
    return rot, trans, scale


def _taylor(x, kind: str, nth: int = 10):
    # models/camera.py:130-156
    ans = torch.zeros_like(x)
    denom = 1.0
    for i in range(nth + 1):
        if kind == "A":
            if i > 0:
                denom *= (2 * i) * (2 * i + 1)
        elif kind == "B":
            denom *= (2 * i + 1) * (2 * i + 2)
        else:
            denom *= (2 * i + 2) * (2 * i + 3)
        ans = ans + (-1) ** i * x ** (2 * i) / denom
    return ans


def se3_to_SE3(wu):
    # models/camera.py:89-102  [...,6] -> [...,3,4]
    w, u = wu.split([3, 3], dim=-1)
    w0, w1, w2 = w.unbind(dim=-1)
    O = torch.zeros_like(w0)
    wx = torch.stack([torch.stack([O, -w2, w1], dim=-1), torch.stack([w2, O, -w0], dim=-1),
                      torch.stack([-w1, w0, O], dim=-1)], dim=-2)
    theta = w.norm(dim=-1)[..., None, None]
    I = torch.eye(3, dtype=wu.dtype)
    A, Bc, C = _taylor(theta, "A"), _taylor(theta, "B"), _taylor(theta, "C")
    R = I + A * wx + Bc * wx @ wx
    V = I + Bc * wx + C * wx @ wx
    return torch.cat([R, V @ u[..., None]], dim=-1)


def compose_pair(pose_a, pose_b):
    # models/camera.py:53-60  pose_new(x) = pose_b o pose_a (x)
    R_a, t_a = pose_a[..., :3], pose_a[..., 3:]
    R_b, t_b = pose_b[..., :3], pose_b[..., 3:]
    return torch.cat([R_b @ R_a, R_b @ t_a + t_b], dim=-1)


def barf_poses(se3_refine, noise_poses):
    # exp_runner.py:419-424
    return compose_pair(se3_to_SE3(se3_refine), noise_poses[:, :3, :])


# ----------------------------------------------------------------------------------------
# Ray generation  (models/dataset.py:656-671) and near/far (models/dataset.py:835-842)
# ----------------------------------------------------------------------------------------
def gen_rays(pose34, intr_inv, px, py):
    p = torch.stack([px, py, torch.ones_like(py)], dim=-1).to(pose34.dtype)
    p = torch.matmul(intr_inv[None, :3, :3], p[:, :, None]).squeeze(-1)
    p_norm = torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    v = p / p_norm
    rays_v = torch.matmul(pose34[None, :3, :3], v[:, :, None]).squeeze(-1)
    rays_o = pose34[None, :3, 3].expand(rays_v.shape)
    return rays_o, rays_v


def near_far_from_sphere(rays_o, rays_d):
    a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return mid - 1.0, mid + 1.0


# ----------------------------------------------------------------------------------------
# Dense SDF grid query  (models/renderer.py:9-37 with query_func = -sdf, :506)
# ----------------------------------------------------------------------------------------
def extract_fields(sdf_p: Params, bound_min, bound_max, resolution: int, chunk: int = 64, **kw):
    X = torch.linspace(float(bound_min[0]), float(bound_max[0]), resolution).split(chunk)
    Y = torch.linspace(float(bound_min[1]), float(bound_max[1]), resolution).split(chunk)
    Z = torch.linspace(float(bound_min[2]), float(bound_max[2]), resolution).split(chunk)
    u = torch.zeros([resolution] * 3, dtype=torch.float32)
    with torch.no_grad():
        for xi, xs in enumerate(X):
            for yi, ys in enumerate(Y):
                for zi, zs in enumerate(Z):
                    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing="ij")
                    pts = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], -1)
                    val = -sdf_value(sdf_p, pts, **kw).reshape(len(xs), len(ys), len(zs))
                    u[xi * chunk: xi * chunk + len(xs), yi * chunk: yi * chunk + len(ys),
                      zi * chunk: zi * chunk + len(zs)] = val
    return u


# ----------------------------------------------------------------------------------------
# Whole train step: rays from pose -> render -> loss -> backward
# ----------------------------------------------------------------------------------------
def train_step(sdf_p: Params, col_p: Params, variance, pose34, intr_inv, px, py, true_rgb, mask, *,
               t_rand, n_samples=64, n_importance=64, up_sample_steps=4, cos_anneal_ratio=1.0,
               igr_weight=0.1, mask_weight=0.0, background_rgb=None):
    """One reference train iteration on a given pose (exp_runner.py:497-599, 772-802).
    Tensors that require grad must be set up by the caller; returns (losses, render_out)."""
    rays_o, rays_d = gen_rays(pose34, intr_inv, px, py)
    near, far = near_far_from_sphere(rays_o, rays_d)
    out = render(sdf_p, col_p, variance, rays_o, rays_d, near, far, n_samples=n_samples,
                 n_importance=n_importance, up_sample_steps=up_sample_steps, t_rand=t_rand,
                 background_rgb=background_rgb, cos_anneal_ratio=cos_anneal_ratio)
    losses = loss_block(out, true_rgb, mask, igr_weight, mask_weight)
    return losses, out


# ----------------------------------------------------------------------------------------
# Flow / reprojection loss and unit-sphere loss  (exp_runner.py:605-688, 714-724)
# pinned by tests/golden/flow_*.npz, which oracle/gen_golden.py produces by exec'ing those very
# source lines of the reference
# ----------------------------------------------------------------------------------------
def sample_points(rays_o, rays_d, z_vals, sample_dist: float):
    """pts [B,S,3] = o + d * mid_z with mid_z = z + dists/2, last dist = sample_dist (models/renderer.py:261-272)."""
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], sample_dist)], dim=-1)
    mid = z_vals + 0.5 * dists
    return rays_o[:, None, :] + rays_d[:, None, :] * mid[:, :, None]


def reprojection_error(pts, weights, c2w34, K33, xy):
    """err [n,2] = sum_j w_j * (project(pts_j into the frame with pose c2w, intrinsics K) - xy)
    (exp_runner.py:634-654): w2c = inverse of the 4x4 pose, pixel = (K cam)[:2] / (K cam)[2]."""
    c2w = torch.cat([c2w34[:3, :4], torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=c2w34.dtype)], dim=0)
    w2c = torch.linalg.inv(c2w)[:3]
    cam = pts @ w2c[:, :3].T + w2c[:, 3]
    pix = cam @ K33[:3, :3].T
    uv = pix[..., :2] / pix[..., 2:3]
    return ((uv - xy[:, None, :]) * weights[..., None]).sum(dim=1)


def flow_loss(rays_o, rays_d, z_vals, weights, c2w_1, c2w_0, K_1, K_0, pixels_xy, pixels_xy_corr, sample_dist: float,
              flow_weight: float, maintain_shape: bool = False, detach_flow_on_sdf: bool = False):
    """exp_runner.py:605-688.  Batch layout [frame-0 rays | frame-1 rays | (additional rays)]; frame-0 points go
    into frame 1 against pixels_xy, frame-1 points into frame 0 against pixels_xy_corr; each term is
    mean(|err|) * flow_weight (F.l1_loss against zeros)."""
    B = z_vals.shape[0]
    pts = sample_points(rays_o, rays_d, z_vals, sample_dist)
    if detach_flow_on_sdf:
        weights = weights.detach()
    n = B // 4 if maintain_shape else B // 2
    second = slice(n, 2 * n) if maintain_shape else slice(n, B)
    e0 = reprojection_error(pts[:n], weights[:n], c2w_1, K_1, pixels_xy)
    e1 = reprojection_error(pts[second], weights[second], c2w_0, K_0, pixels_xy_corr)
    return (e0.abs().mean() + e1.abs().mean()) * flow_weight


def unit_sphere_loss(rays_o, rays_d, z_vals, weights, sample_dist: float, unit_sphere_weight: float):
    """exp_runner.py:714-724: mean |w| over the samples whose point lies outside the unit sphere (mask detached)."""
    pts = sample_points(rays_o, rays_d, z_vals, sample_dist)
    outside = (pts.norm(dim=-1) > 1.0).detach()
    return weights[outside].abs().mean() * unit_sphere_weight
