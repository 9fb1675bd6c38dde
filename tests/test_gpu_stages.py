"""GPU parity of the non-MLP stages (pose/ray-gen, hierarchical sampling, compositing + losses,
ray reduction) against the CPU oracle, forward and backward."""
import numpy as np
import pytest
import torch

from oracle import neus_oracle as O
from tests._util import load_golden, params_from, t

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, atol, rtol=1e-4, msg=""):
    np.testing.assert_allclose(a.detach().cpu().double().numpy(), b.detach().cpu().double().numpy(), atol=atol,
                               rtol=rtol, err_msg=msg)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_raygen_pose_fwd_bwd(mode):
    from fmov_pose_b200 import ops
    torch.manual_seed(mode)
    B = 777
    intr = torch.tensor([[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1.0]])
    intr_inv = torch.linalg.inv(intr).contiguous()
    px = torch.randint(0, 640, [B]); py = torch.randint(0, 480, [B])
    init = torch.eye(4); init[:3, :3] = O.rodrigues_exp(torch.tensor([[0.1, -0.2, 0.05]]))[0]; init[:3, 3] = torch.tensor([0.1, -0.2, -3.0])
    rot = (torch.randn(3) * 0.3).requires_grad_(True)
    trans = (torch.randn(3) * 0.1).requires_grad_(True)
    scale = torch.tensor([1.1], requires_grad=True)
    se3 = (torch.randn(6) * 0.2).requires_grad_(True)
    if mode == 0:
        pose = init[:3].clone().requires_grad_(True)
        kw = dict(c2w34=pose.detach().to(DEV).contiguous())
    elif mode == 1:
        pose = O.pose_gf_compose(rot, trans, init[:3], scale[0])
        kw = dict(rot=rot.detach().to(DEV), trans=trans.detach().to(DEV), scale=scale.detach().to(DEV), init34=init.to(DEV))
    else:
        pose = O.compose_pair(O.se3_to_SE3(se3[None]), init[None, :3])[0]
        kw = dict(se3=se3.detach().to(DEV), init34=init.to(DEV))
    ro, rd = O.gen_rays(pose, intr_inv, px, py)
    near, far = O.near_far_from_sphere(ro, rd)
    g_o, g_d, g_n, g_f = torch.randn(B, 3), torch.randn(B, 3), torch.randn(B, 1), torch.randn(B, 1)
    leaves = {0: [pose], 1: [rot, trans, scale], 2: [se3]}[mode]
    grads = torch.autograd.grad((ro * g_o).sum() + (rd * g_d).sum() + (near * g_n).sum() + (far * g_f).sum(), leaves)
    o2, d2, n2, f2, c2w = ops.raygen_fwd(mode, intr_inv.to(DEV), px.to(DEV), py.to(DEV), **kw)
    _close(o2, ro, 1e-6); _close(d2, rd, 2e-6); _close(n2, near, 1e-5); _close(f2, far, 1e-5)
    _close(c2w, pose, 2e-6)
    g34 = ops.raygen_bwd(intr_inv.to(DEV), px.to(DEV), py.to(DEV), o2, d2, g_o.to(DEV), g_d.to(DEV), g_n.to(DEV), g_f.to(DEV))
    if mode == 0:
        _close(g34, grads[0], 1e-3 * grads[0].abs().max().item())
    else:
        g_rot, g_trans, g_scale, g_se3 = ops.pose_bwd(mode, g34, **kw)
        got = [g_rot, g_trans, g_scale] if mode == 1 else [g_se3]
        for a, b in zip(got, grads):
            _close(a, b, 2e-3 * b.abs().max().item() + 1e-6)


def test_pose_kat_small_angles():
    """Rodrigues at theta -> 0 and the Taylor se3 map against the reference's own outputs (kat.npz)."""
    from fmov_pose_b200 import ops
    d = load_golden("kat")
    eye = torch.eye(4, device=DEV)
    for r, R in zip(d["exp.r"], d["exp.R"]):
        out = ops.pose_fwd(1, rot=torch.tensor(r, device=DEV), trans=torch.zeros(3, device=DEV), init34=eye)
        np.testing.assert_allclose(out[:, :3].cpu().numpy(), R, atol=2e-6)
    for wu, Rt in zip(d["se3.wu"], d["se3.Rt"]):
        out = ops.pose_fwd(2, se3=torch.tensor(wu, device=DEV), init34=eye)
        np.testing.assert_allclose(out.cpu().numpy(), Rt, atol=2e-6)


def _rays(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([0.0, 0.0, -3.0]).repeat(B, 1) + torch.randn(B, 3, generator=g) * 0.02
    d = torch.nn.functional.normalize(torch.tensor([0.0, 0.0, 1.0]).repeat(B, 1) + torch.randn(B, 3, generator=g) * 0.12, dim=-1)
    near, far = O.near_far_from_sphere(o, d)
    return o, d, near, far, g


def test_sample_pdf_kat_through_round_kernel():
    """up_sample on the reference's KAT rays (all-outside ray, grazing ray), kat.npz 'up.*'."""
    from fmov_pose_b200 import ops
    d = load_golden("kat")
    o, dd, z, sdf = t(d, "up.o"), t(d, "up.d"), t(d, "up.z"), t(d, "up.sdf")
    for inv_s in (64, 512):
        zz = torch.zeros(3, 80); zz[:, :64] = z
        ss = torch.zeros(3, 80); ss[:, :64] = sdf
        zz, ss = zz.to(DEV), ss.to(DEV)
        ops.sample_round(o.to(DEV), dd.to(DEV), zz, ss, 64, 0, True, 16, float(inv_s))
        np.testing.assert_allclose(zz[:, 64:80].cpu().numpy(), d[f"up.new{inv_s}"], atol=2e-5)


def test_sample_pdf_edge_kats_through_the_kernel():
    """The reference's own sample_pdf outputs (kat.npz 'pdf.*': all-zero weights, dead bins, a single spike, exact ties at
    the u grid, saturated first / last bin; models/renderer.py:54-86) through the inverse-CDF device code of the round
    kernel (fmov_sample_pdf = the same warp_sample_pdf function on caller-given weights)."""
    from fmov_pose_b200 import ops
    d = load_golden("kat")
    bins, w = t(d, "pdf.bins").to(DEV), t(d, "pdf.w").to(DEV)
    for n_new in (16, 5):
        got = ops.sample_pdf(bins, w, n_new).cpu().numpy()
        ref = d[f"pdf.out{n_new}"]
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, atol=2e-5, err_msg=f"n_new={n_new}")
        assert np.all(np.diff(got, axis=1) >= -1e-6), "inverse CDF of increasing u must be monotone"
    # the oracle restatement on the same inputs (pinned bit-exactly to the reference in test_oracle_golden.py)
    got = ops.sample_pdf(bins, w, 16).cpu()
    np.testing.assert_allclose(got.numpy(), O.sample_pdf_det(bins.cpu(), w.cpu(), 16).numpy(), atol=2e-5)


@pytest.mark.parametrize("n,m,steps,perturb", [(64, 64, 4, True), (16, 32, 2, True), (32, 0, 4, False), (64, 64, 1, False)])
def test_hierarchical_sampling_vs_oracle(n, m, steps, perturb):
    """z_vals of renderer.py:385-446. The SDF used for importance sampling comes from the fp16 tensor-core
    chain, so individual samples can move slightly; compare with the oracle driven by the SAME sdf
    function (exact algorithmic parity) and with the fp32 oracle (tolerance)."""
    from fmov_pose_b200 import ops, packing
    d = load_golden("full_6464_gf")
    p = params_from(d, "sdf.")
    W = [O.eff_weight(p, "", l).to(DEV) for l in range(9)]
    b = [p[f"lin{l}.bias"].to(DEV) for l in range(9)]
    qw = packing.SdfQueryWeights(W, b)
    B = 203
    o, dd, near, far, g = _rays(B, 5)
    t_rand = torch.rand(B, 1, generator=g) if perturb else None
    z = ops.hierarchical_sample(qw, o.to(DEV), dd.to(DEV), near.to(DEV), far.to(DEV),
                                None if t_rand is None else t_rand.to(DEV), n, m, steps)
    torch.cuda.synchronize()
    zc = z.cpu()
    assert torch.all(zc[:, 1:] >= zc[:, :-1]), "z must be sorted"
    # (a) exact-algorithm check: oracle sampling with the GPU sdf as the query function
    def gpu_sdf(q):
        return ops.sdf_query_points(qw, q.to(DEV)).cpu()
    z_ref = O.coarse_z(near, far, n, t_rand)
    if m > 0:
        sdf = gpu_sdf((o[:, None] + dd[:, None] * z_ref[..., None]).reshape(-1, 3)).reshape(B, n)
        for i in range(steps):
            new_z = O.up_sample(o, dd, z_ref, sdf, m // steps, 64 * 2 ** i)
            # (a0) every round IN ISOLATION: the kernel on exactly the oracle's state of this round (same z, same sdf), so
            # that no flip of an earlier round can cascade.  The inverse CDF is discontinuous where denom < 1e-5 toggles
            # (renderer.py:81-82): a sample can then move, but only inside its own bin
            S_i = z_ref.shape[1]
            zz = torch.zeros(B, S_i + m // steps)
            zz[:, :S_i] = z_ref
            ss = torch.zeros(B, S_i + m // steps)
            ss[:, :S_i] = sdf
            zz, ss = zz.to(DEV), ss.to(DEV)
            ops.sample_round(o.to(DEV), dd.to(DEV), zz, ss, S_i, 0, True, m // steps, float(64 * 2 ** i))
            dz = (zz[:, S_i:].cpu() - new_z).abs()
            widest_bin = (z_ref[:, 1:] - z_ref[:, :-1]).max(dim=1, keepdim=True)[0]
            print(f"  round {i} in isolation (n={n} m={m}): |dz| > 2e-5 for {(dz > 2e-5).float().mean().item():.4f} of the new "
                  f"samples, max {dz.max().item():.2e}, widest bin {widest_bin.max().item():.2e}")
            assert bool((dz <= widest_bin + 1e-5).all()), (i, dz.max().item())
            assert (dz > 2e-5).float().mean().item() <= 0.05, (i, (dz > 2e-5).float().mean().item(), dz.max().item())
            z_ref, sdf = O.cat_z_vals(gpu_sdf, o, dd, z_ref, new_z, sdf, last=(i + 1 == steps))
    # The reference's inverse CDF is discontinuous where a bin's pdf sits at the `denom < 1e-5` switch
    # (renderer.py:81-82): empty bins have pdf ~ 1e-5/sum, so ulp-level differences (expf, scan order) move
    # a few samples inside their bin.  Moves are bounded by the bin width.
    diff = (zc - z_ref).abs()
    # (a moved sample also shifts the sorted positions between its old and new place, so compare as 1-D
    # transport cost: mean |diff| tiny, max bounded by a coarse bin)
    print(f"hierarchical sampling n={n} m={m} steps={steps}: vs oracle-with-GPU-sdf mean |dz| {diff.mean().item():.2e} "
          f"max {diff.max().item():.2e}, rays with any |dz| > 1e-4: {(diff.max(dim=1)[0] > 1e-4).float().mean().item():.3f}")
    # measured on B200 (round 2, three library builds): 64+64 / 4 rounds mean 2.2e-4 .. 4.0e-4, max 4.6e-2 .. 1.6e-1; 16+32 /
    # 2 rounds 1.1e-5, 2.3e-2; 64+64 / 1 round 3.1e-6, 6.0e-3; 32+0 exactly 0.  Which samples flip depends on ulps, and a
    # flip re-weights every later round (the per-round check (a0) above is the strict one; this end-to-end comparison can
    # only bound the cascade): mean below 1e-3, max below two coarse bins of n = 16
    assert diff.mean().item() < 1e-3, diff.mean().item()
    assert diff.max().item() < 0.25, diff.max().item()
    # (b) against the fp32 oracle end to end: samples are positions along the ray, tolerance 2e-3
    z32 = O.sample_z(p, o, dd, near, far, n, m, steps, t_rand)
    # (a sample that flips bins in an early round re-weights the later rounds, so only the bulk statistics are
    #  comparable against an oracle that uses a *different* (fp32) SDF)
    d32 = (zc - z32).abs()
    assert d32.median().item() < 1e-4 and d32.mean().item() < 2e-3, (d32.median().item(), d32.mean().item())


@pytest.mark.parametrize("S,bg,car", [(128, False, 1.0), (32, True, 0.4), (48, False, 0.0), (200, False, 1.0)])
def test_composite_and_loss_fwd_bwd(S, bg, car):
    from fmov_pose_b200 import ops
    B = 150
    o, dd, near, far, g = _rays(B, 7)
    z = torch.sort(near + (far - near) * torch.rand(B, S, generator=g), dim=-1)[0]
    z.requires_grad_(True); o.requires_grad_(True); dd.requires_grad_(True)
    mid = (z.detach() + 0.01)
    pts = o.detach()[:, None] + dd.detach()[:, None] * mid[..., None]
    sdf = ((pts.norm(dim=-1) - 0.6).reshape(-1, 1) + torch.randn(B * S, 1, generator=g) * 0.01).requires_grad_(True)
    nrm = (torch.nn.functional.normalize(pts.reshape(-1, 3), dim=-1) * (1 + 0.2 * torch.randn(B * S, 1, generator=g))).requires_grad_(True)
    rgb = torch.rand(B, S, 3, generator=g).requires_grad_(True)
    variance = torch.tensor(0.35, requires_grad=True)
    inv_s = O.inv_s_from_variance(variance)
    bgt = torch.tensor([[1.0, 0.5, 0.25]]) if bg else None
    sd = 2.0 / 64
    ref = O.composite(o, dd, z, sdf, nrm, rgb, inv_s, sd, background_rgb=bgt, cos_anneal_ratio=car)
    true_rgb = torch.rand(B, 3, generator=g)
    mask = (torch.rand(B, 1, generator=g) > 0.4).float()
    out = {"color_fine": ref["color"], "gradient_error": ref["gradient_error"],
           "weight_sum": ref["weights"].sum(-1, keepdim=True)}
    losses = O.loss_block(out, true_rgb, mask, 0.1, 5.0)
    depth = (ref["weights"] * ref["mid_z_vals"]).sum(-1, keepdim=True)
    gw_ext = torch.randn(B, S, generator=g) * 1e-3
    total = losses["loss"] + 0.01 * depth.sum() + (ref["weights"] * gw_ext).sum()
    grads = torch.autograd.grad(total, [sdf, nrm, rgb, variance, o, dd, z], allow_unused=True)

    D = lambda x: None if x is None else x.detach().to(DEV).contiguous()
    inv_s_d = D(inv_s.reshape(1))
    f = ops.composite_fwd(D(o), D(dd), D(z), D(sdf.reshape(-1)), D(nrm), D(rgb.reshape(-1, 3)), inv_s_d, sd, car,
                          bg=None if bgt is None else D(bgt.reshape(3)))
    _close(f["color"], ref["color"], 2e-5); _close(f["weights"], ref["weights"], 2e-5)
    _close(f["cdf"], ref["cdf"], 2e-5); _close(f["mid_z"], ref["mid_z_vals"], 1e-5)
    _close(f["inside"], ref["inside_sphere"], 0); _close(f["pts"], ref["pts"], 1e-5)
    _close(f["depth"], depth, 2e-5)
    _close(f["weight_max"], ref["weights"].max(-1, keepdim=True)[0], 2e-5)
    eik = f["eik"].sum(0)
    _close(eik[0] / (eik[1] + 1e-5), ref["gradient_error"], 1e-5)
    mm = (mask > 0.5).float()
    mask_sum = D((mm.sum() + 1e-5).reshape(1))
    partial, g_color, g_wsum = ops.loss_fwd_bwd(f["color"], f["weight_sum"], D(true_rgb), D(mask), mask_sum, B, 5.0)
    _close(partial[:, 0].sum(), losses["color_loss"], 1e-5); _close(partial[:, 1].sum(), losses["mask_loss"], 1e-5)
    g_eik = torch.tensor([0.1], device=DEV)
    bk = ops.composite_bwd(D(o), D(dd), D(z), D(sdf.reshape(-1)), D(nrm), D(rgb.reshape(-1, 3)), inv_s_d, sd, car,
                           None if bgt is None else D(bgt.reshape(3)), g_color, g_wsum,
                           torch.full((B,), 0.01, device=DEV), D(gw_ext), g_eik, eik[1:2].contiguous())
    tol = lambda x: 2e-4 * x.abs().max().item() + 1e-9
    _close(bk["d_sdf"], grads[0].reshape(-1), tol(grads[0]), rtol=2e-3, msg="d_sdf")
    _close(bk["d_nrm"], grads[1], tol(grads[1]), rtol=2e-3, msg="d_nrm")
    _close(bk["d_rgb"], grads[2].reshape(-1, 3), tol(grads[2]), rtol=2e-3, msg="d_rgb")
    d_var = bk["d_invs"].sum() * 10.0 * inv_s_d[0]
    _close(d_var, grads[3], 2e-3 * abs(grads[3].item()), msg="d_variance")
    # rays / z gradients with zero point- and dir-gradients from the (absent) MLP
    zero = torch.zeros(B * S, 3, device=DEV)
    d_o, d_d, d_z = ops.ray_reduce_bwd(zero, None, bk["d_dir"], bk["d_dist"], bk["d_mid"], D(dd), D(z), sd, True)
    # o and d also enter through pts -> relax/inside masks only (no gradient), so autograd grads are:
    _close(d_d, grads[5], tol(grads[5]) + 1e-7, rtol=2e-3, msg="d_rays_d")
    _close(d_z, grads[6], tol(grads[6]) + 1e-7, rtol=2e-3, msg="d_z")
    assert grads[4] is None and d_o.abs().max().item() == 0.0


def test_ray_reduce_bwd_points_path():
    from fmov_pose_b200 import ops
    B, S = 64, 96
    o, dd, near, far, g = _rays(B, 9)
    z = torch.sort(near + (far - near) * torch.rand(B, S, generator=g), dim=-1)[0].requires_grad_(True)
    o.requires_grad_(True); dd.requires_grad_(True)
    sd = 2.0 / 64
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((B, 1), sd)], -1)
    mid = z + dists * 0.5
    pts = o[:, None] + dd[:, None] * mid[..., None]
    dirs = dd[:, None].expand(B, S, 3)
    gp, gdirs = torch.randn(B, S, 3, generator=g), torch.randn(B, S, 3, generator=g)
    grads = torch.autograd.grad((pts * gp).sum() + (dirs * gdirs).sum(), [o, dd, z])
    D = lambda x: x.detach().to(DEV).contiguous()
    zb = torch.zeros(B, S, device=DEV)
    d_o, d_d, d_z = ops.ray_reduce_bwd(D(gp.reshape(-1, 3)), D(gdirs.reshape(-1, 3)), torch.zeros(B, 3, device=DEV), zb, zb,
                                       D(dd), D(z), sd, True)
    _close(d_o, grads[0], 1e-4); _close(d_d, grads[1], 2e-4); _close(d_z, grads[2], 1e-4)


@pytest.mark.parametrize("emphasize_rot,small_rot,cam", [(False, False, 3), (True, False, 7), (True, True, 0)])
def test_fused_pose_module_vs_oracle(emphasize_rot, small_rot, cam):
    """LearnPoseGF.forward as one launch per direction (fmov_pose_gf_fwd/_bwd) against the oracle's torch formulation
    (Fourier features -> GELU MLP -> heads -> Rodrigues tail), pose and every parameter gradient."""
    from fmov_pose_b200.models.picture_pose import LearnPoseGF
    torch.manual_seed(5)
    np.random.seed(5)
    init = torch.eye(4).repeat(9, 1, 1)
    init[:, :3, :3] = O.rodrigues_exp(torch.tensor([[0.2, -0.1, 0.3]]))[0]
    init[:, :3, 3] = torch.tensor([0.1, -0.2, -3.0])
    m = LearnPoseGF(9, init_c2w=init.clone(), emphasize_rot=emphasize_rot, small_rot=small_rot)
    with torch.no_grad():                  # off the near-zero init so that every term carries signal
        for n_, p_ in m.named_parameters():
            if n_.startswith("lin3"):
                p_.add_(torch.randn_like(p_) * 0.05)
    if emphasize_rot:
        for p_ in m.lin3_trans.parameters():
            p_.requires_grad = True        # exercised here although the reference freezes it
    p_cpu = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and k != "b" and k != "init_c2w")
             for k, v in m.state_dict().items()}
    rot, trans, scale = O.pose_gf_mlp(p_cpu, cam, emphasize_rot, small_rot)
    ref = O.pose_gf_compose(rot, trans, init[cam, :3, :], scale)
    G = torch.randn(3, 4)
    (ref * G).sum().backward()
    m = m.to(DEV)
    pose = m(cam)
    assert pose.shape == (4, 4) and float(pose[3, 3]) == 1.0
    np.testing.assert_allclose(pose[:3].detach().cpu().numpy(), ref.detach().numpy(), atol=2e-6)
    (pose[:3] * G.to(DEV)).sum().backward()
    for k, v in m.named_parameters():
        if not v.requires_grad:
            continue
        g_ref = p_cpu[k].grad
        assert v.grad is not None, k
        err = float((v.grad.cpu() - g_ref).abs().max() / (g_ref.abs().max() + 1e-12))
        assert err <= 2e-4, (k, err)
    # device-side frame index (CUDA-graph path) gives the same pose
    pose_t = m(cam, torch.tensor([cam], device=DEV))
    np.testing.assert_array_equal(pose_t.detach().cpu().numpy(), pose.detach().cpu().numpy())


def test_fused_weight_norm_vs_torch():
    """fmov_weight_norm_fwd/_bwd (all layers of both MLPs in one launch) == torch._weight_norm per layer, values and the
    gradients on weight_g / weight_v, including the 217-row and 3-row layers."""
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.models.fields import RenderingNetwork, SDFNetwork
    from fmov_pose_b200.weight_norm import effective_weights_fused
    torch.manual_seed(3)
    sdf = SDFNetwork(**synthetic.SDF_KW).to(DEV)
    col = RenderingNetwork(**synthetic.COL_KW).to(DEV)
    with torch.no_grad():
        for net in (sdf, col):
            for n_, p_ in net.named_parameters():
                if n_.endswith("weight_g"):
                    p_.mul_(1.0 + 0.3 * torch.rand_like(p_))
    (Ws, bs), (Wc, bc) = effective_weights_fused([sdf, col])
    tw = lambda net, l: torch._weight_norm(getattr(net, f"lin{l}").weight_v, getattr(net, f"lin{l}").weight_g, 0)
    Ws_ref = [tw(sdf, l) for l in range(9)]
    Wc_ref = [tw(col, l) for l in range(5)]
    assert len(Ws) == 9 and len(Wc) == 5 and bs[3].shape == (217,)
    gen = torch.Generator(device=DEV).manual_seed(1)
    Gs = [torch.randn(w.shape, device=DEV, generator=gen) for w in list(Ws) + list(Wc)]
    for a, b in zip(list(Ws) + list(Wc), list(Ws_ref) + list(Wc_ref)):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-6, atol=1e-8)
    params = [p for net in (sdf, col) for n_, p in net.named_parameters() if "weight_" in n_]
    loss = sum((w * g_).sum() for w, g_ in zip(list(Ws) + list(Wc), Gs))
    got = torch.autograd.grad(loss, params)
    loss_ref = sum((w * g_).sum() for w, g_ in zip(list(Ws_ref) + list(Wc_ref), Gs))
    ref = torch.autograd.grad(loss_ref, params)
    for a, b in zip(got, ref):
        err = float((a - b).abs().max() / (b.abs().max() + 1e-12))
        assert err <= 1e-5, err
