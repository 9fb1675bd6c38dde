"""Thin Python wrappers over the C-ABI entry points (one per exported function family)."""
import ctypes

import torch

from . import _lib as L


def sdf_query_points(qw, pts, in_scale=1.0, out_scale=1.0, precise=False):
    """SDFNetwork.sdf(pts) under no_grad (models/fields.py:106-107) -> [P,1].  `precise`: split-precision chain
    (qw built with precise=True), ~1e-5 of the fp32 network instead of ~1e-3."""
    pts = L.f32c(pts)
    P = pts.shape[0]
    out = torch.empty(P, 1, dtype=torch.float32, device=pts.device)
    if P and precise:
        # precise=True: full split (needs the residual images blob_lo); precise="act": activation-split chain (blob_lo = NULL)
        assert precise == "act" or qw.blob_lo is not None, "SdfQueryWeights(..., precise=True) is needed for the full split"
        L.check(L.lib().fmov_sdf_query_points_precise(L.ptr(pts), L.c_ll(P), L.ptr(qw.blob), L.ptr(None if precise == "act" else qw.blob_lo),
                                                      L.ptr(qw.bias), L.ptr(qw.w8), L.ptr(qw.b8), L.c_float(in_scale),
                                                      L.c_float(out_scale), L.ptr(out), L.stream()),
                "fmov_sdf_query_points_precise")
    elif P:
        L.check(L.lib().fmov_sdf_query_points(L.ptr(pts), L.c_ll(P), L.ptr(qw.blob), L.ptr(qw.bias), L.ptr(qw.w8),
                                              L.ptr(qw.b8), L.c_float(in_scale), L.c_float(out_scale), L.ptr(out),
                                              L.stream()), "fmov_sdf_query_points")
    return out


def sdf_query_rays(qw, rays_o, rays_d, z, S, z_off=0, in_scale=1.0, out_scale=1.0):
    """sdf(o + d*z[:, z_off:z_off+S]) -> [B,S] (models/renderer.py:425-428, :225-232)"""
    B = rays_o.shape[0]
    assert z.is_contiguous() and z.dtype == torch.float32
    out = torch.empty(B, S, dtype=torch.float32, device=z.device)
    if B and getattr(qw, "blob_pair", None) is not None:          # train step: CTA-pair engine, biases in the weight images
        with L.timed("sdf_query"):
            L.check(L.lib().fmov_sdf_query_rays_pair(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.c_ll(B), S, z.shape[1], z_off,
                                                     L.ptr(qw.blob_pair), L.ptr(qw.w8), L.ptr(qw.b8), L.c_float(in_scale),
                                                     L.c_float(out_scale), L.ptr(out), L.stream()), "fmov_sdf_query_rays_pair")
    elif B:
        L.check(L.lib().fmov_sdf_query_rays(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.c_ll(B), S, z.shape[1], z_off,
                                            L.ptr(qw.blob), L.ptr(qw.bias), L.ptr(qw.w8), L.ptr(qw.b8),
                                            L.c_float(in_scale), L.c_float(out_scale), L.ptr(out), L.stream()),
                "fmov_sdf_query_rays")
    return out


def sdf_query_grid(qw, bmin, bmax, res, first, count, out, in_scale=1.0, out_scale=-1.0, precise=False):
    """-sdf on points [first, first+count) of the x-major res^3 grid (models/renderer.py:9-37, :506)."""
    bm = (ctypes.c_float * 3)(*[float(v) for v in bmin])
    bx = (ctypes.c_float * 3)(*[float(v) for v in bmax])
    if precise:
        assert precise == "act" or qw.blob_lo is not None, "SdfQueryWeights(..., precise=True) is needed for the full split"
        with L.timed("sdf_query_grid_precise"):
            L.check(L.lib().fmov_sdf_query_grid_precise(bm, bx, int(res), L.c_ll(first), L.c_ll(count), L.ptr(qw.blob),
                                                        L.ptr(None if precise == "act" else qw.blob_lo), L.ptr(qw.bias), L.ptr(qw.w8), L.ptr(qw.b8),
                                                        L.c_float(in_scale), L.c_float(out_scale), L.ptr(out),
                                                        L.stream()), "fmov_sdf_query_grid_precise")
        return out
    L.check(L.lib().fmov_sdf_query_grid(bm, bx, int(res), L.c_ll(first), L.c_ll(count), L.ptr(qw.blob), L.ptr(qw.bias),
                                        L.ptr(qw.w8), L.ptr(qw.b8), L.c_float(in_scale), L.c_float(out_scale),
                                        L.ptr(out), L.stream()), "fmov_sdf_query_grid")
    return out


# ---------------------------------------------------------------------------------------------
# sampling
# ---------------------------------------------------------------------------------------------
def sample_coarse(near, far, t_rand, n_samples, z_stride):
    """coarse z [B, z_stride] with the first n_samples filled (models/renderer.py:385-405)."""
    B = near.shape[0]
    z = torch.zeros(B, z_stride, dtype=torch.float32, device=near.device)
    near_c, far_c = L.f32c(near.reshape(-1)), L.f32c(far.reshape(-1))
    tr = None if t_rand is None else L.f32c(t_rand.reshape(-1))
    L.check(L.lib().fmov_sample_coarse(L.ptr(near_c), L.ptr(far_c), L.ptr(tr), L.c_ll(B), n_samples, z_stride,
                                       L.ptr(z), L.stream()), "fmov_sample_coarse")
    return z


def sample_round(rays_o, rays_d, z, sdf, n_sorted, n_tail, with_sdf, n_new, inv_s):
    B = z.shape[0]
    with L.timed("sample_round"):
      L.check(L.lib().fmov_sample_round(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.ptr(sdf), L.c_ll(B), z.shape[1],
                                      n_sorted, n_tail, int(with_sdf), n_new, L.c_float(inv_s), L.stream()),
            "fmov_sample_round")


def sample_pdf(bins, weights, n_new):
    """sample_pdf(bins, weights, n_new, det=True) (models/renderer.py:54-86) -> [B, n_new]"""
    bins, weights = L.f32c(bins), L.f32c(weights)
    B, n = bins.shape
    assert tuple(weights.shape) == (B, n - 1)
    out = torch.empty(B, n_new, dtype=torch.float32, device=bins.device)
    L.check(L.lib().fmov_sample_pdf(L.ptr(bins), L.ptr(weights), L.c_ll(B), n, int(n_new), L.ptr(out), L.stream()),
            "fmov_sample_pdf")
    return out


def hierarchical_sample(qw, rays_o, rays_d, near, far, t_rand, n_samples, n_importance, up_sample_steps, scale=1.0):
    """z_vals [B, n_samples+n_importance] — renderer.py:385-446 (coarse z, no-grad SDF queries,
    up_sample rounds with inv_s = 64*2^i, cat_z_vals merges). All device work, no host sync."""
    S = n_samples + (n_importance if n_importance > 0 else 0)
    z = sample_coarse(near, far, t_rand, n_samples, S)
    if n_importance <= 0:
        return z
    B = rays_o.shape[0]
    m = n_importance // up_sample_steps
    sdf = torch.empty(B, S, dtype=torch.float32, device=z.device)
    lib = L.lib()

    pair = getattr(qw, "blob_pair", None)          # FineWeights: half-major images -> the CTA-pair engine

    def query(z_off, cnt):
      with L.timed("sdf_query"):
        if pair is not None:
            L.check(lib.fmov_sdf_query_rays_pair(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.c_ll(B), cnt, S, z_off,
                                                 L.ptr(pair), L.ptr(qw.w8), L.ptr(qw.b8), L.c_float(scale),
                                                 L.c_float(1.0 / scale), L.ptr(sdf_tmp), L.stream()),
                    "fmov_sdf_query_rays_pair")
        else:
            L.check(lib.fmov_sdf_query_rays(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.c_ll(B), cnt, S, z_off,
                                            L.ptr(qw.blob), L.ptr(qw.bias), L.ptr(qw.w8), L.ptr(qw.b8),
                                            L.c_float(scale), L.c_float(1.0 / scale), L.ptr(sdf_tmp), L.stream()),
                    "fmov_sdf_query_rays")

    # sdf is kept [B,S] row-aligned with z; queries write a dense [B,cnt] block that is scattered in
    sdf_tmp = torch.empty(B * max(n_samples, m), dtype=torch.float32, device=z.device)
    query(0, n_samples)
    sdf[:, :n_samples] = sdf_tmp[: B * n_samples].view(B, n_samples)
    n_cur = n_samples
    for i in range(up_sample_steps):
        last = i + 1 == up_sample_steps
        # merge the previous round's tail (if any) and draw the next m samples
        sample_round(rays_o, rays_d, z, sdf, n_cur - (m if i > 0 else 0), m if i > 0 else 0, True, m, 64.0 * 2 ** i)
        if not last:
            query(n_cur, m)
            sdf[:, n_cur:n_cur + m] = sdf_tmp[: B * m].view(B, m)
        n_cur += m
    sample_round(rays_o, rays_d, z, None, n_cur - m, m, False, 0, 0.0)   # final merge, z only
    return z


# ---------------------------------------------------------------------------------------------
# pose + rays
# ---------------------------------------------------------------------------------------------
def raygen_fwd(mode, intr_inv, px, py, c2w34=None, rot=None, trans=None, scale=None, init34=None, se3=None):
    B = px.shape[0]
    dev = px.device
    assert px.dtype == py.dtype and px.dtype in (torch.int64, torch.float32), "pixels: int64 or float32 (sub-pixel)"
    fn = L.lib().fmov_raygen_fwd if px.dtype == torch.int64 else L.lib().fmov_raygen_xy_fwd
    rays_o = torch.empty(B, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(B, 3, dtype=torch.float32, device=dev)
    near = torch.empty(B, 1, dtype=torch.float32, device=dev)
    far = torch.empty(B, 1, dtype=torch.float32, device=dev)
    c2w_out = torch.empty(3, 4, dtype=torch.float32, device=dev)
    L.check(fn(mode, L.ptr(c2w34), L.ptr(rot), L.ptr(trans), L.ptr(scale), L.ptr(init34), L.ptr(se3),
                                    L.ptr(intr_inv), intr_inv.stride(0), L.ptr(px), L.ptr(py), L.c_ll(B), L.ptr(rays_o),
                                    L.ptr(rays_d), L.ptr(near), L.ptr(far), L.ptr(c2w_out), L.stream()), "fmov_raygen_fwd")
    return rays_o, rays_d, near, far, c2w_out


def raygen_bwd(intr_inv, px, py, rays_o, rays_d, g_o, g_d, g_near, g_far):
    B = px.shape[0]
    g34 = torch.empty(3, 4, dtype=torch.float32, device=px.device)
    fn = L.lib().fmov_raygen_bwd if px.dtype == torch.int64 else L.lib().fmov_raygen_xy_bwd
    L.check(fn(L.ptr(intr_inv), intr_inv.stride(0), L.ptr(px), L.ptr(py), L.c_ll(B), L.ptr(rays_o),
                                    L.ptr(rays_d), L.ptr(g_o), L.ptr(g_d), L.ptr(g_near), L.ptr(g_far), L.ptr(g34),
                                    L.stream()), "fmov_raygen_bwd")
    return g34


def pose_fwd(mode, rot=None, trans=None, scale=None, init34=None, se3=None):
    dev = init34.device
    out = torch.empty(3, 4, dtype=torch.float32, device=dev)
    L.check(L.lib().fmov_pose_fwd(mode, L.ptr(rot), L.ptr(trans), L.ptr(scale), L.ptr(init34), L.ptr(se3), L.ptr(out),
                                  L.stream()), "fmov_pose_fwd")
    return out


def pose_bwd(mode, g34, rot=None, trans=None, scale=None, init34=None, se3=None):
    dev = init34.device
    z = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)
    g_rot, g_trans, g_scale, g_se3 = z(3), z(3), z(1), z(6)
    L.check(L.lib().fmov_pose_bwd(mode, L.ptr(rot), L.ptr(trans), L.ptr(scale), L.ptr(init34), L.ptr(se3), L.ptr(g34),
                                  L.ptr(g_rot), L.ptr(g_trans), L.ptr(g_scale if scale is not None else None),
                                  L.ptr(g_se3), L.stream()), "fmov_pose_bwd")
    return g_rot, g_trans, g_scale, g_se3


# ---------------------------------------------------------------------------------------------
# compositing
# ---------------------------------------------------------------------------------------------
def composite_fwd(rays_o, rays_d, z, sdf, nrm, rgb, inv_s, sample_dist, cos_anneal, bg=None, full=True):
    B, S = z.shape
    dev = z.device
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    out = dict(color=e(B, 3), weight_sum=e(B, 1), weight_max=e(B, 1), depth=e(B, 1), eik=e(B, 2))
    if full:
        out.update(weights=e(B, S), cdf=e(B, S), inside=e(B, S), mid_z=e(B, S), pts=e(B * S, 3))
    g = lambda k: L.ptr(out.get(k))
    with L.timed("composite_fwd"):
      L.check(L.lib().fmov_composite_fwd(L.c_ll(B), S, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.ptr(sdf), L.ptr(nrm),
                                       L.ptr(rgb), L.ptr(inv_s), L.c_float(sample_dist), L.c_float(cos_anneal),
                                       L.ptr(bg), g("color"), g("weight_sum"), g("weight_max"), g("depth"), g("weights"),
                                       g("cdf"), g("inside"), g("mid_z"), g("pts"), g("eik"), L.stream()),
            "fmov_composite_fwd")
    return out


def composite_bwd(rays_o, rays_d, z, sdf, nrm, rgb, inv_s, sample_dist, cos_anneal, bg, g_color, g_wsum, g_depth,
                  g_weights, g_eik, eik_den, g_nrm_ext=None):
    B, S = z.shape
    dev = z.device
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    out = dict(d_sdf=e(B * S), d_nrm=e(B * S, 3), d_rgb=e(B * S, 3), d_dir=e(B, 3), d_dist=e(B, S), d_mid=e(B, S),
               d_invs=e(B))
    with L.timed("composite_bwd"):
      L.check(L.lib().fmov_composite_bwd(L.c_ll(B), S, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.ptr(sdf), L.ptr(nrm),
                                       L.ptr(rgb), L.ptr(inv_s), L.c_float(sample_dist), L.c_float(cos_anneal),
                                       L.ptr(bg), L.ptr(g_color), L.ptr(g_wsum), L.ptr(g_depth), L.ptr(g_weights),
                                       L.ptr(g_eik), L.ptr(eik_den), L.ptr(g_nrm_ext), L.ptr(out["d_sdf"]),
                                       L.ptr(out["d_nrm"]), L.ptr(out["d_rgb"]), L.ptr(out["d_dir"]), L.ptr(out["d_dist"]),
                                       L.ptr(out["d_mid"]), L.ptr(out["d_invs"]), L.stream()), "fmov_composite_bwd")
    return out


def loss_fwd_bwd(color, weight_sum, true_rgb, mask, mask_sum, n_rays_global, mask_weight):
    B = color.shape[0]
    dev = color.device
    partial = torch.empty(B, 2, dtype=torch.float32, device=dev)
    g_color = torch.empty(B, 3, dtype=torch.float32, device=dev)
    g_wsum = torch.empty(B, 1, dtype=torch.float32, device=dev)
    L.check(L.lib().fmov_loss_fwd_bwd(L.ptr(color), L.ptr(weight_sum), L.ptr(true_rgb), L.ptr(mask), L.c_ll(B),
                                      L.ptr(mask_sum), L.c_ll(n_rays_global), L.c_float(mask_weight), L.ptr(partial),
                                      L.ptr(g_color), L.ptr(g_wsum), L.stream()), "fmov_loss_fwd_bwd")
    return partial, g_color, g_wsum


def ray_reduce_bwd(d_pts, d_dirs, d_dir_tc, d_dist, d_mid, rays_d, z, sample_dist, want_dz):
    B, S = z.shape
    dev = z.device
    d_o = torch.empty(B, 3, dtype=torch.float32, device=dev)
    d_d = torch.empty(B, 3, dtype=torch.float32, device=dev)
    d_z = torch.empty(B, S, dtype=torch.float32, device=dev) if want_dz else None
    L.check(L.lib().fmov_ray_reduce_bwd(L.ptr(d_pts), L.ptr(d_dirs), L.ptr(d_dir_tc), L.ptr(d_dist), L.ptr(d_mid),
                                        L.ptr(rays_d), L.ptr(z), L.c_ll(B), S, L.c_float(sample_dist), L.ptr(d_o),
                                        L.ptr(d_d), L.ptr(d_z), L.stream()), "fmov_ray_reduce_bwd")
    return d_o, d_d, d_z


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[(t.data_ptr() if t is not None else 0) for t in tensors])


def pose_gf_fwd(cid_t, b, W1, b1, W2, b2, heads, rot_k, init_all):
    """LearnPoseGF.forward in one launch -> (c2w34 [3,4], save).  heads: [(W [r,64], bias [r]), ...]"""
    dev = W1.device
    assert cid_t.dtype == torch.int64 and cid_t.is_cuda and cid_t.numel() == 1
    save = torch.empty(int(L.lib().fmov_pose_gf_save_floats()), dtype=torch.float32, device=dev)
    c2w = torch.empty(3, 4, dtype=torch.float32, device=dev)
    rows = (ctypes.c_int * len(heads))(*[int(w.shape[0]) for w, _ in heads])
    L.check(L.lib().fmov_pose_gf_fwd(L.ptr(cid_t), L.ptr(b), L.ptr(W1), L.ptr(b1), L.ptr(W2), L.ptr(b2), len(heads),
                                     _ptr_array([w for w, _ in heads]), _ptr_array([x for _, x in heads]), rows,
                                     L.c_float(rot_k), L.ptr(init_all), L.ptr(save), L.ptr(c2w), L.stream()),
            "fmov_pose_gf_fwd")
    return c2w, save


def pose_gf_bwd(cid_t, b, W1, b1, W2, b2, heads, rot_k, init_all, save, g34, need):
    """need: booleans per (W1, b1, W2, b2, then W/b of every head) -> gradients in the same order (None where not needed)"""
    dev = W1.device
    mk = lambda t, n: torch.empty_like(t) if n else None
    dW1, db1, dW2, db2 = mk(W1, need[0]), mk(b1, need[1]), mk(W2, need[2]), mk(b2, need[3])
    dWh = [mk(w, need[4 + 2 * i]) for i, (w, _) in enumerate(heads)]
    dbh = [mk(x, need[5 + 2 * i]) for i, (_, x) in enumerate(heads)]
    rows = (ctypes.c_int * len(heads))(*[int(w.shape[0]) for w, _ in heads])
    L.check(L.lib().fmov_pose_gf_bwd(L.ptr(cid_t), L.ptr(b), L.ptr(W1), L.ptr(b1), L.ptr(W2), L.ptr(b2), len(heads),
                                     _ptr_array([w for w, _ in heads]), _ptr_array([x for _, x in heads]), rows,
                                     L.c_float(rot_k), L.ptr(init_all), L.ptr(save), L.ptr(g34), L.ptr(dW1), L.ptr(db1),
                                     L.ptr(dW2), L.ptr(db2), _ptr_array(dWh), _ptr_array(dbh), L.stream()),
            "fmov_pose_gf_bwd")
    out = [dW1, db1, dW2, db2]
    for w_, b_ in zip(dWh, dbh):
        out += [w_, b_]
    return out
