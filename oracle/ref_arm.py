"""Baseline arm that runs the REFERENCE ITSELF: the unmodified sources under oracle/_ref/fmov_pose (oracle/build_ref.py)
imported as the `models` package, driven by the lines of exp_runner.py that make up one train iteration.

TEST / BASELINE INFRASTRUCTURE ONLY (bench.py `--impl reference`, `cpu_baseline`, the PyTorch-on-B200 arm).  Must run
in its own process: it claims the top-level module name `models`, which fmov_pose_b200.dropin would alias.

What is the reference's code here: SDFNetwork / BarfSDFNetwork, RenderingNetwork, SingleVarianceNetwork, NeuSRenderer
(render, up_sample, sample_pdf, cat_z_vals, render_core), SegLearnPose / LearnPoseGF, Dataset.gen_random_rays_at and
near_far_from_sphere (GPU arm: called on a Dataset shell whose attributes are set without running the disk loader),
torch.optim.Adam as exp_runner.py:258-269 builds them.  What is restated (exp_runner.py needs pyhocon / trimesh /
open3d / ... to import): the loss block exp_runner.py:562-599, 772-779 and the optimiser calls :785-816;
on the CPU arm also the ray maths of models/dataset.py:656-671, because that function moves its tensors to 'cuda:0'."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref", "fmov_pose")

SDF_KW = dict(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
              geometric_init=True, weight_norm=True)          # confs/ho3d_virtual.conf:79-90
COL_KW = dict(d_feature=256, mode="idr", d_in=9, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
              multires_view=4, squeeze_out=True)              # confs/ho3d_virtual.conf:96-106
INTRINSICS = [[600.0, 0.0, 320.0], [0.0, 600.0, 240.0], [0.0, 0.0, 1.0]]


def available():
    return os.path.exists(os.path.join(REF_ROOT, "models", "renderer.py"))


def _import_reference():
    for p in (os.path.join(HERE, "_stubs"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    assert "models" not in sys.modules or REF_ROOT in getattr(sys.modules["models"], "__path__", [""])[0], \
        "the reference arm needs its own process (the module name `models` is taken)"
    import warnings
    warnings.filterwarnings("ignore")
    from models.barf_fields import BarfRenderingNetwork, BarfSDFNetwork
    from models.fields import SingleVarianceNetwork
    from models.picture_pose import SegLearnPose
    from models.renderer import NeuSRenderer
    return BarfSDFNetwork, BarfRenderingNetwork, SingleVarianceNetwork, SegLearnPose, NeuSRenderer


def build_step(B, n_samples, n_importance, up_sample_steps, device="cpu", n_images=20, two_frames=False, seed=2024):
    """-> step() running one reference train iteration on `B` rays (2 x B/2 rays of two frames with `two_frames`, the
    maintain_shape iteration of confs/ho3d_virtual.conf) and returning the loss tensor."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    gpu = str(device).startswith("cuda")
    if gpu:
        torch.set_default_tensor_type("torch.cuda.FloatTensor")          # exp_runner.py:2030
    Sdf, Col, Var, SegPose, Renderer = _import_reference()
    torch.manual_seed(seed)                                               # exp_runner.py:29-30
    np.random.seed(seed)
    H, W = 480, 640
    init = torch.eye(4).repeat(n_images, 1, 1)
    init[:, :3, 3] = torch.tensor([0.0, 0.0, -3.0])
    dev = torch.device(device)
    sdf = Sdf(noise_poses=init.clone(), n_images=n_images, barf=True, **SDF_KW).to(dev)
    col = Col(**COL_KW).to(dev)
    var = Var(0.3).to(dev)
    pose_net = SegPose(num_cams=n_images, segment_img_num=1, init_c2w=init.clone(), emphasize_rot=True).to(dev)
    pose_net.initialized_flag.data[:] = True
    sdf.se3_refine.weight.requires_grad_(False)                           # exp_runner.py:225-227
    params = list(sdf.parameters()) + list(var.parameters()) + list(col.parameters())
    opt = torch.optim.Adam(params, lr=5e-4)                               # exp_runner.py:264-269
    pose_opts = [torch.optim.Adam(m.parameters(), lr=5e-4) for m in pose_net.pose_mlps]      # :258-262
    rend = Renderer(None, sdf, var, col, n_samples=n_samples, n_importance=n_importance, n_outside=0,
                    up_sample_steps=up_sample_steps, perturb=1.0)
    g = torch.Generator(device="cpu").manual_seed(1)
    images = torch.rand(n_images, H, W, 3, generator=g, device="cpu")
    ys, xs = torch.meshgrid(torch.arange(H, device="cpu"), torch.arange(W, device="cpu"), indexing="ij")
    disc = (((xs - W // 2) ** 2 + (ys - H // 2) ** 2) < 150 ** 2).float()
    masks = disc[None, :, :, None].repeat(n_images, 1, 1, 3)
    K4 = torch.eye(4, device="cpu")
    K4[:3, :3] = torch.tensor(INTRINSICS, device="cpu")
    intr_inv = torch.inverse(K4)[None].repeat(n_images, 1, 1).to(dev)
    ds = None
    if gpu:           # the reference's own ray functions on a Dataset shell (its __init__ is the disk loader: not run)
        from models.dataset import Dataset
        ds = object.__new__(Dataset)
        ds.images, ds.masks, ds.masks_np = images.cpu(), masks.cpu(), masks.cpu().numpy()
        ds.H, ds.W, ds.intrinsics_all_inv, ds.use_mono_depth = H, W, intr_inv, False
    else:
        from models.dataset import Dataset
        images, masks = images.to(dev), masks.to(dev)
    near_far = Dataset.near_far_from_sphere                                # models/dataset.py:835-842 (does not use self)
    it = [0]

    def rays_cpu(img_id, n, pose):
        # models/dataset.py:656-671 restated (the original moves its tensors to 'cuda:0'); mask-bbox pixel range as
        # mask_guided_sampling draws it for the synthetic disc mask
        px = torch.randint(170, 470, [n], generator=g, device="cpu")
        py = torch.randint(90, 390, [n], generator=g, device="cpu")
        color, mask = images[img_id][(py, px)], masks[img_id][(py, px)]
        p = torch.stack([px, py, torch.ones_like(py)], dim=-1).float()
        p = torch.matmul(intr_inv[img_id, None, :3, :3], p[:, :, None]).squeeze()
        rays_v = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
        rays_v = torch.matmul(pose[None, :3, :3], rays_v[:, :, None]).squeeze()
        rays_o = pose[None, :3, 3].expand(rays_v.shape)
        return torch.cat([rays_o, rays_v, color, mask[:, :1]], dim=-1)

    def step():
        i = it[0]
        it[0] += 1
        frames = [(7 * i) % n_images] + ([(7 * i + 3) % n_images] if two_frames else [])
        n_per = B // len(frames)
        data = []
        for f in frames:
            img_id = torch.tensor(f).long()
            pose = pose_net(img_id)[:3]                                    # exp_runner.py:497-498
            if gpu:
                d, _ = ds.gen_random_rays_at(img_id, n_per, pose=pose, mask_guided_sampling=True, patch_size=10)
            else:
                d = rays_cpu(f, n_per, pose)
            data.append(d)
        data = torch.cat(data, dim=0)
        rays_o, rays_d, true_rgb, mask = data[:, :3], data[:, 3:6], data[:, 6:9], data[:, 9:10]
        near, far = near_far(None, rays_o, rays_d)                         # exp_runner.py:557
        mask = (mask > 0.5).float()                                        # :562-567 (mask_weight = 5)
        mask_sum = mask.sum() + 1e-5
        out = rend.render(rays_o, rays_d, near, far, background_rgb=None, cos_anneal_ratio=1.0)      # :568-575
        color_error = (out["color_fine"] - true_rgb) * mask
        color_loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum      # :585-588
        mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)                  # :599
        loss = color_loss + out["gradient_error"] * 0.1 + mask_loss * 5.0                                    # :772-779
        for f in frames:
            pose_opts[f].zero_grad()                                       # :793-795
        opt.zero_grad()                                                    # :801
        loss.backward()                                                    # :802
        opt.step()                                                         # :812
        for f in set(frames):
            pose_opts[f].step()                                            # :814-816
        return loss
    return step


def time_cpu(B, n, m, up, warm, iters, threads=None, two_frames=False):
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    fn = build_step(B, n, m, up, device="cpu", two_frames=two_frames)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return B / med, med, threads


def time_gpu(B, n, m, up, warm=3, iters=8, two_frames=False):
    """reference-PyTorch on this GPU (north_star's 50x denominator): CUDA-event timed, returns (rays/s, ms/step)"""
    import torch
    fn = build_step(B, n, m, up, device="cuda:0", two_frames=two_frames)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return B / ms * 1e3, ms


if __name__ == "__main__":       # python oracle/ref_arm.py gpu|cpu B n m [two]   -> one JSON line
    import json
    kind, B, n, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    two = len(sys.argv) > 5 and sys.argv[5] == "two"
    if kind == "gpu":
        rps, ms = time_gpu(B, n, m, 4, two_frames=two)
        print(json.dumps({"rays_per_s": rps, "ms_per_step": ms, "rays": B, "n_samples": n, "n_importance": m,
                          "two_frames": two, "impl": "reference sources under torch.set_default_tensor_type(cuda)"}))
    else:
        rps, sec, th = time_cpu(B, n, m, 4, 1, 3, two_frames=two)
        print(json.dumps({"rays_per_s": rps, "ms_per_step": sec * 1e3, "rays": B, "threads": th}))
