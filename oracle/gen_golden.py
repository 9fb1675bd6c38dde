"""Generate golden fixtures by running the IMPORTED REFERENCE (read-only /root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference ships no golden vectors, so these fixtures — outputs of the
reference's own `NeuSRenderer.render`, fields, pose modules and `camera.lie` on fixed seeds —
are what pins the oracle (`tests/test_oracle_golden.py`) and, through it, the CUDA path.
Driver lines that cannot be imported (exp_runner.py needs pyhocon/open3d/…; dataset.py needs
plyfile) are restated here citing their lines: ray-gen dataset.py:656-671, near/far :835-842,
loss block exp_runner.py:562-599,772-779, BARF pose compose :419-424.
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FMOV_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from models.renderer import NeuSRenderer, sample_pdf, extract_fields  # noqa: E402
from models.fields import SDFNetwork, RenderingNetwork, SingleVarianceNetwork, NeRF  # noqa: E402
from models.barf_fields import BarfSDFNetwork, BarfRenderingNetwork  # noqa: E402
from models.picture_pose import LearnPoseGF, SegLearnPose  # noqa: E402
from models import camera  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")

SDF_KW = dict(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5,
              scale=1.0, geometric_init=True, weight_norm=True)   # confs/ho3d_virtual.conf:79-90
COL_KW = dict(d_feature=256, mode="idr", d_in=9, d_out=3, d_hidden=256, n_layers=4,
              weight_norm=True, multires_view=4, squeeze_out=True)  # confs/ho3d_virtual.conf:96-106
INTR = np.array([[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1.0]], dtype=np.float32)


def sd_np(mod, prefix):
    return {prefix + k: v.detach().cpu().numpy() for k, v in mod.state_dict().items()}


def perturb_params(mod, std, gen):
    """Move weights off the geometric init so every term of the path is exercised
    (layer-0 PE columns and the skip PE columns are exactly zero at init)."""
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if p.dtype.is_floating_point and p.requires_grad and p.dim() > 0:
                p.add_(torch.randn(p.shape, generator=gen) * std * (p.abs().mean() + 1e-3))


def gen_rays_ref(pose, intr_inv, px, py):
    # models/dataset.py:660-671
    p = torch.stack([px, py, torch.ones_like(py)], dim=-1).float()
    p = torch.matmul(intr_inv[None, :3, :3], p[:, :, None]).squeeze()
    p_norm = torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    rays_v = p / p_norm
    rays_v = torch.matmul(pose[None, :3, :3], rays_v[:, :, None]).squeeze()
    rays_o = pose[None, :3, 3].expand(rays_v.shape)
    return rays_o, rays_v


def near_far_ref(rays_o, rays_d):
    # models/dataset.py:835-842
    a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return mid - 1.0, mid + 1.0


def loss_ref(render_out, true_rgb, mask, igr_weight, mask_weight):
    # exp_runner.py:562-599, 772-779
    if mask_weight > 0.0:
        mask = (mask > 0.5).float()
    else:
        mask = torch.ones_like(mask)
    mask_sum = mask.sum() + 1e-5
    color_fine = render_out["color_fine"]
    color_error = (color_fine - true_rgb) * mask
    color_fine_loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum
    eikonal_loss = render_out["gradient_error"]
    mask_loss = F.binary_cross_entropy(render_out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    loss = color_fine_loss + eikonal_loss * igr_weight + mask_loss * mask_weight
    return loss, color_fine_loss, eikonal_loss, mask_loss


def sub(a, n=4096):
    """strided subsample of a flattened array (keeps fixtures small)"""
    f = a.reshape(-1)
    step = max(1, f.size // n)
    return f[::step][:n].copy()


def render_case(name, *, B, n_samples, n_importance, up_steps, pose_kind, mask_weight, seed,
                barf=False, white=False, cos_anneal=1.0, perturb_std=0.0, d_hidden=256):
    torch.manual_seed(seed)
    np.random.seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    sdf_kw = dict(SDF_KW, d_hidden=d_hidden)
    col_kw = dict(COL_KW, d_hidden=d_hidden, d_feature=d_hidden)
    sdf_kw["d_out"] = d_hidden + 1
    n_img = 6
    init_c2w = torch.eye(4).repeat(n_img, 1, 1)
    init_c2w[:, :3, 3] = torch.tensor([0.0, 0.0, -3.0])
    # small per-frame deviation so R0 != I
    for i in range(n_img):
        w = torch.tensor([[0.03 * i, -0.02 * i, 0.01 * i]])
        from models.batch_lie_group_helper import Exp
        init_c2w[i, :3, :3] = Exp(w)[0]
    if barf:
        sdf_net = BarfSDFNetwork(init_c2w.clone(), n_images=n_img, **sdf_kw)
        col_net = BarfRenderingNetwork(**col_kw)
        sdf_net.progress.data.fill_(0.3)
        col_net.progress.data.fill_(0.3)
    else:
        sdf_net = SDFNetwork(**sdf_kw)
        col_net = RenderingNetwork(**col_kw)
    var_net = SingleVarianceNetwork(0.3)
    if perturb_std > 0:
        perturb_params(sdf_net, perturb_std, gen)
        perturb_params(col_net, perturb_std, gen)
    nerf = None
    renderer = NeuSRenderer(nerf, sdf_net, var_net, col_net, n_samples, n_importance, 0, up_steps, 1.0)

    img_id = 3
    extra = {}
    if pose_kind == "seg":      # ho3d_virtual.conf: SegLearnPose + emphasize_rot
        pose_net = SegLearnPose(n_img, 1, init_c2w=init_c2w.clone(), emphasize_rot=True)
        # make the pose MLP produce a non-trivial rotation
        with torch.no_grad():
            for m in pose_net.pose_mlps:
                m.lin3_rot.weight.mul_(3.0)
        pose_net.initialized_flag.data[:] = True
        pose = pose_net(torch.tensor(img_id))[:3]
        pose_params = [p for p in pose_net.pose_mlps[img_id].parameters() if p.requires_grad]
        mlp = pose_net.pose_mlps[img_id]
        extra.update({"pose." + k: v.detach().numpy() for k, v in mlp.state_dict().items()})
        extra["pose.b"] = mlp.b.detach().numpy()
    elif pose_kind == "gf":     # ho3d_barf.conf / ho3d_global_womask.conf: LearnPoseGF
        pose_net = LearnPoseGF(n_img, init_c2w=init_c2w.clone())
        with torch.no_grad():
            pose_net.lin3.weight.mul_(3.0)
        pose = pose_net(torch.tensor(img_id))[:3]
        pose_params = [p for p in pose_net.parameters() if p.requires_grad]
        extra.update({"pose." + k: v.detach().numpy() for k, v in pose_net.state_dict().items()})
        extra["pose.b"] = pose_net.b.detach().numpy()
    elif pose_kind == "se3":    # BARF se3_refine path, exp_runner.py:419-424
        with torch.no_grad():
            sdf_net.se3_refine.weight.copy_(torch.randn(n_img, 6, generator=gen) * 0.05)
        pose_refine = camera.lie.se3_to_SE3(sdf_net.se3_refine.weight)
        pose_all = camera.pose.compose([pose_refine, sdf_net.noise_poses[:, :3, :]])
        pose = pose_all[img_id, :3]
        pose_params = [sdf_net.se3_refine.weight]
        extra["pose_all"] = pose_all.detach().numpy()
    else:
        raise ValueError(pose_kind)

    intr_inv = torch.from_numpy(np.linalg.inv(INTR)).float()
    px = torch.randint(low=120, high=520, size=[B], generator=gen)
    py = torch.randint(low=40, high=440, size=[B], generator=gen)
    true_rgb = torch.rand(B, 3, generator=gen)
    mask = (((px - 320) ** 2 + (py - 240) ** 2) < 150 ** 2).float()[:, None]
    rays_o, rays_d = gen_rays_ref(pose, intr_inv, px, py)
    near, far = near_far_ref(rays_o, rays_d)

    # the jitter draw of renderer.py:404 — reproduce the global-RNG draw the reference makes
    torch.manual_seed(seed + 7)
    t_rand = torch.rand([B, 1])
    torch.manual_seed(seed + 7)
    bg = torch.ones([1, 3]) if white else None
    out = renderer.render(rays_o, rays_d, near, far, background_rgb=bg, cos_anneal_ratio=cos_anneal)
    loss, cl, el, ml = loss_ref(out, true_rgb, mask, 0.1, mask_weight)
    params = list(sdf_net.parameters()) + list(col_net.parameters()) + list(var_net.parameters())
    names = (["sdf." + n for n, _ in sdf_net.named_parameters()]
             + ["col." + n for n, _ in col_net.named_parameters()] + ["variance"])
    for p in params + pose_params:
        p.grad = None
    rays_o.retain_grad() if rays_o.requires_grad else None
    rays_d.retain_grad()
    pose.retain_grad()
    loss.backward()

    d = {}
    d.update(sd_np(sdf_net, "sdf."))
    d.update(sd_np(col_net, "col."))
    d["variance"] = var_net.variance.detach().numpy()
    d.update(extra)
    d.update(dict(init_c2w=init_c2w.numpy(), img_id=np.int64(img_id), intr_inv=intr_inv.numpy(),
                  px=px.numpy(), py=py.numpy(), true_rgb=true_rgb.numpy(), mask=mask.numpy(),
                  t_rand=t_rand.numpy(), pose=pose.detach().numpy(),
                  rays_o=rays_o.detach().numpy(), rays_d=rays_d.detach().numpy(),
                  near=near.detach().numpy(), far=far.detach().numpy(),
                  cfg=np.array([B, n_samples, n_importance, up_steps, d_hidden], dtype=np.int64),
                  mask_weight=np.float32(mask_weight), white=np.int64(white),
                  cos_anneal=np.float32(cos_anneal), pose_kind=np.array(pose_kind)))
    for k in ["color_fine", "depth_fine", "s_val", "cdf_fine", "weight_sum", "weight_max",
              "gradients", "weights", "gradient_error", "inside_sphere", "pts"]:
        d["out." + k] = out[k].detach().numpy()
    d["loss"] = np.array([loss.item(), cl.item(), el.item(), ml.item()], dtype=np.float64)
    d["grad.pose"] = pose.grad.numpy()
    d["grad.rays_d"] = rays_d.grad.numpy()
    for n, p in zip(names, params):
        if p.grad is None:
            continue
        g = p.grad.detach().numpy()
        d["gnorm." + n] = np.float64(np.linalg.norm(g.astype(np.float64)))
        if g.size <= 4096 or d_hidden < 256:
            d["grad." + n] = g
        else:
            d["gsub." + n] = sub(g)
    for i, p in enumerate(pose_params):
        d[f"grad.pose_param{i}"] = p.grad.detach().numpy() if p.grad is not None else np.zeros(1)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: loss={loss.item():.6f} color={cl.item():.5f} eik={el.item():.5f} mask={ml.item():.5f} "
          f"-> {os.path.getsize(path) / 1e6:.2f} MB")


def kat_case():
    """Small known-answer vectors for pieces with tricky edge semantics."""
    torch.manual_seed(11)
    d = {}
    # sample_pdf: ties, zero-weight bins, saturated bins (renderer.py:54-86)
    bins = torch.sort(torch.rand(7, 17) * 2 + 1, dim=-1)[0]
    w = torch.rand(7, 16)
    w[0] = 0.0                      # all-zero weights -> uniform
    w[1, 3:9] = 0.0                 # dead bins in the middle
    w[2] = 0.0; w[2, 5] = 1.0       # single spike
    w[3, :] = 1.0                   # exact ties at the u grid
    w[4, -1] = 50.0
    w[5, 0] = 50.0
    d["pdf.bins"], d["pdf.w"] = bins.numpy(), w.numpy()
    d["pdf.out16"] = sample_pdf(bins, w, 16, det=True).numpy()
    d["pdf.out5"] = sample_pdf(bins, w, 5, det=True).numpy()
    # Rodrigues near theta -> 0 and generic (batch_lie_group_helper.py:19-35)
    from models.batch_lie_group_helper import Exp
    r = torch.tensor([[0.0, 0.0, 0.0], [1e-9, 0, 0], [1e-4, -2e-4, 3e-4], [0.3, -0.2, 0.1],
                      [3.0, 0.1, -0.2], [0.0, 3.14159, 0.0]])
    d["exp.r"], d["exp.R"] = r.numpy(), Exp(r).numpy()
    # se3_to_SE3 (camera.py:89-102)
    wu = torch.randn(5, 6) * torch.tensor([[1e-6], [1e-2], [0.1], [1.0], [2.5]])
    d["se3.wu"], d["se3.Rt"] = wu.numpy(), camera.lie.se3_to_SE3(wu).numpy()
    pa, pb = camera.lie.se3_to_SE3(torch.randn(5, 6) * 0.3), camera.lie.se3_to_SE3(torch.randn(5, 6) * 0.3)
    d["compose.a"], d["compose.b"] = pa.numpy(), pb.numpy()
    d["compose.out"] = camera.pose.compose([pa, pb]).numpy()
    # Softplus threshold crossing (beta*z around 20) and embedder
    z = torch.tensor([-1.0, -0.2, -0.01, 0.0, 0.01, 0.1999, 0.2, 0.2001, 0.5, 3.0])
    d["sp.z"], d["sp.y"] = z.numpy(), torch.nn.Softplus(beta=100)(z).numpy()
    from models.embedder import get_embedder
    from models.barf_embedder import get_embedder as get_barf
    x = torch.randn(9, 3)
    e6, _ = get_embedder(6); e4, _ = get_embedder(4); b6, _ = get_barf(6)
    d["pe.x"], d["pe.e6"], d["pe.e4"] = x.numpy(), e6(x).numpy(), e4(x).numpy()
    d["pe.barf6_p03"] = b6(x, torch.tensor(0.3)).numpy()
    # up_sample with all-outside rays and a ray grazing the sphere (renderer.py:168-220)
    rend = NeuSRenderer(None, None, None, None, 64, 64, 0, 4, 1.0)
    o = torch.tensor([[0.0, 0.0, -3.0], [0.0, 5.0, -3.0], [0.9, 0.0, -3.0]])
    dd = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 0.0, 1.0]])
    z = torch.linspace(2.0, 4.0, 64)[None].repeat(3, 1)
    pts = o[:, None] + dd[:, None] * z[..., None]
    sdf = pts.norm(dim=-1) - 0.5
    d["up.o"], d["up.d"], d["up.z"], d["up.sdf"] = o.numpy(), dd.numpy(), z.numpy(), sdf.numpy()
    for inv_s in (64, 512):
        d[f"up.new{inv_s}"] = rend.up_sample(o, dd, z, sdf, 16, inv_s).numpy()
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **d)
    print("kat.npz written")


def grid_case():
    """extract_fields on a 40^3 grid (renderer.py:9-37) with a perturbed full-size SDF net."""
    torch.manual_seed(5)
    gen = torch.Generator().manual_seed(6)
    net = SDFNetwork(**SDF_KW)
    perturb_params(net, 0.05, gen)
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
    res = 40
    u = extract_fields(bmin, bmax, res, lambda pts: -net.sdf(pts))
    # state dict of the full net is 2.1 MB; ship only the seed recipe + a weight checksum
    d = {"u": u.astype(np.float32), "res": np.int64(res)}
    d.update(sd_np(net, "sdf."))
    np.savez_compressed(os.path.join(OUT, "grid40.npz"), **d)
    print("grid40.npz written", u.min(), u.max())


def _reference_block(start_marker, end_marker):
    """Source text of a block of Runner.train (exp_runner.py cannot be imported: pyhocon/open3d/... are missing), cut
    between two marker lines and dedented, so that the fixture below EXECUTES the reference's own lines."""
    import textwrap
    lines = open(os.path.join(REF, "exp_runner.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip() == start_marker)
    i1 = next(i for i in range(i0 + 1, len(lines)) if lines[i].strip() == end_marker)
    return textwrap.dedent("\n".join(lines[i0:i1])), (i0 + 1, i1)


def flow_case(name, *, B, S, maintain_shape, detach_flow_on_sdf=False, seed=0):
    """Flow / reprojection loss (exp_runner.py:604-688) and unit-sphere loss (:714-724): the reference's source lines
    are exec'd on synthetic render outputs; pts = o + d*mid_z is built as render_core does (renderer.py:261-272)."""
    from types import SimpleNamespace
    g = torch.Generator().manual_seed(seed)
    src_flow, span_flow = _reference_block("if self.flow_weight > 0.0 and use_flow:", "if self.depth_weight > 0.0:")
    src_unit, span_unit = _reference_block("if self.unit_sphere_weight > 0:", "if self.gradient_analysis:")
    n_samples = S
    sample_dist = 2.0 / n_samples

    def rand_pose(t):
        r = torch.randn(3, generator=g) * 0.2
        K_ = torch.zeros(3, 3)
        K_[0, 1], K_[0, 2], K_[1, 2] = -r[2], r[1], -r[0]
        K_ = K_ - K_.T
        R = torch.linalg.matrix_exp(K_)
        return torch.cat([R, torch.tensor(t)[:, None] + 0.05 * torch.randn(3, 1, generator=g)], dim=1)

    c2w = {0: rand_pose([0.1, 0.0, -3.0]).requires_grad_(True), 1: rand_pose([-0.2, 0.1, -2.8]).requires_grad_(True)}
    intr = torch.eye(4)[None].repeat(2, 1, 1)
    intr[:, :3, :3] = torch.tensor(INTR)
    intr[1, 0, 0], intr[1, 1, 1] = 590.0, 605.0
    # rays: origins at the camera centres of the two frames, directions towards the unit sphere
    n = B // 4 if maintain_shape else B // 2
    owner = torch.zeros(B, dtype=torch.long)
    owner[n:2 * n if maintain_shape else B] = 1
    px = torch.rand(B, generator=g) * 300 + 170
    py = torch.rand(B, generator=g) * 300 + 90
    rays_o = torch.stack([c2w[int(o)][:, 3].detach() for o in owner]).clone().requires_grad_(True)
    dirs = []
    for b in range(B):
        p_ = torch.linalg.inv(intr[int(owner[b]), :3, :3]) @ torch.tensor([px[b], py[b], 1.0])
        dirs.append(c2w[int(owner[b])][:, :3].detach() @ (p_ / p_.norm()))
    rays_d = torch.stack(dirs).clone().requires_grad_(True)
    near, far = near_far_ref(rays_o.detach(), rays_d.detach())
    z = (near + (far - near) * torch.sort(torch.rand(B, S, generator=g), dim=-1)[0]).requires_grad_(True)
    weights = (torch.rand(B, S, generator=g) ** 4 * 0.2 * (torch.rand(B, S, generator=g) > 0.1) - 0.002).requires_grad_(True)
    # models/renderer.py:261-272
    dists = z[..., 1:] - z[..., :-1]
    dists = torch.cat([dists, torch.Tensor([sample_dist]).expand(dists[..., :1].shape)], -1)
    mid_z_vals = z + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z_vals[..., :, None]).reshape(-1, 3)
    render_out = {"pts": pts, "weights": weights}
    pixels_xy = torch.stack([px[:n] + 3.0 * torch.randn(n, generator=g), py[:n] + 3.0 * torch.randn(n, generator=g)], -1)
    pixels_xy_corr = torch.stack([px[n:2 * n] + 3.0 * torch.randn(n, generator=g),
                                  py[n:2 * n] + 3.0 * torch.randn(n, generator=g)], -1)
    self_ = SimpleNamespace(flow_weight=0.1, detach_flow_on_sdf=detach_flow_on_sdf, maintain_shape=maintain_shape,
                            pose_type="gf", pose_network=lambda i: c2w[int(i)], detach_ref=False,
                            dataset=SimpleNamespace(intrinsics_all=intr), unit_sphere_weight=0.05)
    ns = dict(self=self_, use_flow=True, render_out=render_out, torch=torch, F=F, to_hom=camera.to_hom,
              img_id=torch.tensor(1), img_id_corr=torch.tensor(0), pose_all=None, pixels_xy=pixels_xy,
              pixels_xy_corr=pixels_xy_corr)
    exec(compile(src_flow, "exp_runner.py[%d:%d]" % span_flow, "exec"), ns)
    exec(compile(src_unit, "exp_runner.py[%d:%d]" % span_unit, "exec"), ns)
    flow_loss, unit_loss = ns["flow_loss"], ns["unit_sphere_loss"]
    leaves = dict(rays_o=rays_o, rays_d=rays_d, z=z, weights=weights, c2w_0=c2w[0], c2w_1=c2w[1])
    gf = torch.autograd.grad(flow_loss, list(leaves.values()), retain_graph=True, allow_unused=True)
    gu = torch.autograd.grad(unit_loss, [weights], allow_unused=True)
    d = {k: v.detach().numpy() for k, v in leaves.items()}
    d.update({"gflow_" + k: (torch.zeros_like(v) if g_ is None else g_).numpy() for (k, v), g_ in zip(leaves.items(), gf)})
    d.update(gunit_weights=gu[0].numpy(), flow_loss=flow_loss.detach().numpy(), unit_sphere_loss=unit_loss.detach().numpy(),
             pts=pts.detach().numpy(), intrinsics=intr.numpy(), pixels_xy=pixels_xy.numpy(),
             pixels_xy_corr=pixels_xy_corr.numpy(), maintain_shape=np.array(maintain_shape),
             detach_flow_on_sdf=np.array(detach_flow_on_sdf), flow_weight=np.array(0.1), unit_sphere_weight=np.array(0.05),
             sample_dist=np.array(sample_dist), ref_lines_flow=np.array(span_flow), ref_lines_unit=np.array(span_unit))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: flow_loss={flow_loss.item():.6f} unit_sphere_loss={unit_loss.item():.6f} "
          f"(exp_runner.py lines {span_flow}, {span_unit})")


def lr_case():
    """Runner.update_learning_rate (exp_runner.py:1049-1087), exec'd from the reference source on a fake Runner."""
    from types import SimpleNamespace
    src, span = _reference_block("def update_learning_rate(self, pose_mlp_index_set=None):", "def file_backup(self):")
    ns = dict(np=np, torch=torch)
    exec(compile(src, "exp_runner.py[%d:%d]" % span, "exec"), ns)
    fn = ns["update_learning_rate"]
    d = {}
    for tag, base_dir in (("global", "exp/AP13/ours"), ("wo_global", "exp/AP13/ours_wo_global_conf")):
        progress = torch.zeros(3)

        def step_progress(i):
            progress[i] += 1
            return progress[i]
        opt = SimpleNamespace(param_groups=[{"lr": 0.0}])
        pose_opts = [SimpleNamespace(param_groups=[{"lr": 0.0}]) for _ in range(3)]
        self_ = SimpleNamespace(iter_step=0, warm_up_end=50, learning_rate_alpha=0.05, end_iter=400, learning_rate=5e-4,
                                optimizer=opt, pose_type="seg", pose_network=SimpleNamespace(step_progress=step_progress),
                                base_exp_dir=base_dir, max_pro_iteration=120, pose_alpha=0.5, pose_optimizers=pose_opts,
                                pose_lr=3e-4)
        iters = list(range(0, 400, 7))
        net, pose = [], []
        for it in iters:
            self_.iter_step = it
            fn(self_, pose_mlp_index_set={it % 3})
            net.append(opt.param_groups[0]["lr"])
            pose.append([po.param_groups[0]["lr"] for po in pose_opts])
        d[tag + ".iters"] = np.array(iters)
        d[tag + ".net_lr"] = np.array(net, dtype=np.float64)
        d[tag + ".pose_lr"] = np.array(pose, dtype=np.float64)
    d["ref_lines"] = np.array(span)
    np.savez_compressed(os.path.join(OUT, "lr_schedule.npz"), **d)
    print("lr_schedule.npz written (exp_runner.py lines %d-%d)" % span)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    kat_case()
    # small-width cases pin the algorithm cheaply (oracle-vs-reference, CPU suite)
    render_case("small_6464_seg", B=24, n_samples=64, n_importance=64, up_steps=4, pose_kind="seg",
                mask_weight=5.0, seed=101, barf=True, perturb_std=0.1, d_hidden=64)
    render_case("small_3200_seg", B=32, n_samples=32, n_importance=0, up_steps=4, pose_kind="seg",
                mask_weight=5.0, seed=102, barf=True, perturb_std=0.1, d_hidden=64)
    render_case("small_6464_se3_white", B=24, n_samples=64, n_importance=64, up_steps=4,
                pose_kind="se3", mask_weight=1.0, seed=103, barf=True, white=True, cos_anneal=0.4,
                perturb_std=0.1, d_hidden=64)
    render_case("small_1632_gf_nomask", B=16, n_samples=16, n_importance=32, up_steps=2,
                pose_kind="gf", mask_weight=0.0, seed=104, barf=False, perturb_std=0.1, d_hidden=64)
    # full-size (8x256) cases: the shapes the CUDA kernels are built for
    render_case("full_6464_gf", B=48, n_samples=64, n_importance=64, up_steps=4, pose_kind="gf",
                mask_weight=1.0, seed=201, barf=True, perturb_std=0.05)
    render_case("full_3200_seg", B=64, n_samples=32, n_importance=0, up_steps=4, pose_kind="seg",
                mask_weight=5.0, seed=202, barf=True, perturb_std=0.05)
    grid_case()
    lr_case()
    flow_case("flow_half", B=32, S=48, maintain_shape=False, seed=5)
    flow_case("flow_quarter_detach", B=64, S=32, maintain_shape=True, detach_flow_on_sdf=True, seed=6)
