"""CPU tests of the host-side mirror: module construction, checkpoint-key compatibility with the reference
(state_dict keys/shapes recorded in the golden fixtures), explicit errors instead of fallbacks, and the
ray-sharded (world_size 2, gloo) loss normalisers / gradient all-reduce of TrainStep."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests._util import load_golden


def _nets():
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.models.barf_fields import BarfRenderingNetwork, BarfSDFNetwork
    from fmov_pose_b200.models.fields import NeRF, SingleVarianceNetwork
    init = synthetic.make_init_poses(6)
    return (BarfSDFNetwork(init, n_images=6, **synthetic.SDF_KW), BarfRenderingNetwork(**synthetic.COL_KW),
            SingleVarianceNetwork(0.3), NeRF(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, output_ch=4,
                                             skips=[4], use_viewdirs=True))


def test_state_dict_keys_and_shapes_match_reference_checkpoints():
    d = load_golden("full_6464_gf")
    sdf, col, var, nerf = _nets()
    for prefix, net in (("sdf.", sdf), ("col.", col)):
        ref = {k[len(prefix):]: v.shape for k, v in d.items() if k.startswith(prefix)}
        mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert set(ref) == set(mine), set(ref) ^ set(mine)
        for k in ref:
            assert tuple(ref[k]) == mine[k], k
    assert list(var.state_dict()) == ["variance"]
    assert sum(p.numel() for p in nerf.parameters()) == 606_596


def test_pose_module_state_dict_matches_reference():
    from fmov_pose_b200.models.picture_pose import LearnPoseGF, SegLearnPose
    for name, emph in (("full_6464_gf", False), ("full_3200_seg", True)):
        d = load_golden(name)
        ref = {k[5:]: v.shape for k, v in d.items() if k.startswith("pose.")}
        m = LearnPoseGF(6, init_c2w=torch.eye(4).repeat(6, 1, 1), emphasize_rot=emph)
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert set(ref) == set(mine), set(ref) ^ set(mine)
    seg = SegLearnPose(6, 1, init_c2w=torch.eye(4).repeat(6, 1, 1), emphasize_rot=True)
    assert len(seg.pose_mlps) == 6 and bool(seg.initialized_flag[0]) and not bool(seg.initialized_flag[1])
    trainable = sum(p.numel() for p in seg.pose_mlps[0].parameters() if p.requires_grad)
    assert trainable == 20_868


def test_geometric_init_statistics():
    """fields.py:47-79: last layer ~ N(sqrt(pi)/16, 1e-4), bias -0.5; lin0 only xyz columns; lin4 PE columns 0."""
    sdf, _, _, _ = _nets()
    W, b = sdf.effective_weights()
    assert abs(W[8].mean().item() - np.sqrt(np.pi) / 16) < 1e-3 and abs(b[8][0].item() + 0.5) < 1e-6
    assert W[0][:, 3:].abs().max().item() == 0.0 and W[0][:, :3].abs().max().item() > 0
    assert W[4][:, -36:].abs().max().item() == 0.0
    assert tuple(W[3].shape) == (217, 256)


def test_unsupported_configs_raise_instead_of_falling_back():
    from fmov_pose_b200 import fine
    from fmov_pose_b200.models.fields import RenderingNetwork, SDFNetwork
    from fmov_pose_b200.models.renderer import NeuSRenderer
    sdf, col, var, nerf = _nets()
    r = NeuSRenderer(nerf, sdf, var, col, 64, 64, 32, 4, 1.0)
    with pytest.raises(NotImplementedError):
        r._check()
    small = SDFNetwork(d_in=3, d_out=65, d_hidden=64, n_layers=4, skip_in=(2,), multires=6)
    Ws, _ = small.effective_weights()
    Wc, _ = col.effective_weights()
    with pytest.raises(NotImplementedError):
        fine.check_supported(Ws, Wc)
    r2 = NeuSRenderer(None, sdf, var, col, 64, 64, 0, 4, 1.0)
    o = torch.zeros(4, 3)
    with pytest.raises(RuntimeError):          # CPU tensors: no CPU fallback
        r2.render(o, o, torch.zeros(4, 1), torch.ones(4, 1))
    with pytest.raises(NotImplementedError):
        col(o, o, o, torch.zeros(4, 256))


def test_embedder_matches_reference_layout():
    from fmov_pose_b200.models.embedder import get_embedder
    d = load_golden("kat")
    e6, n6 = get_embedder(6)
    e4, n4 = get_embedder(4)
    x = torch.from_numpy(d["pe.x"])
    assert (n6, n4) == (39, 27)
    np.testing.assert_array_equal(e6(x).numpy(), d["pe.e6"])
    np.testing.assert_array_equal(e4(x).numpy(), d["pe.e4"])


def test_camera_host_mirror_matches_reference_kats():
    from fmov_pose_b200.models import camera
    from fmov_pose_b200.models.batch_lie_group_helper import Exp
    d = load_golden("kat")
    np.testing.assert_allclose(camera.lie.se3_to_SE3(torch.from_numpy(d["se3.wu"])).numpy(), d["se3.Rt"], atol=1e-7)
    np.testing.assert_allclose(camera.pose.compose([torch.from_numpy(d["compose.a"]), torch.from_numpy(d["compose.b"])]).numpy(),
                               d["compose.out"], atol=1e-7)
    np.testing.assert_array_equal(Exp(torch.from_numpy(d["exp.r"])).numpy(), d["exp.R"])


# ---- world_size-2 gloo: global normalisers + gradient all-reduce ------------------------------------------
def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fmov_pose_b200.train import TrainStep
    torch.manual_seed(0)
    B = 64
    g = torch.Generator().manual_seed(5)
    color = torch.rand(B, 3, generator=g)
    wsum = torch.rand(B, 1, generator=g)
    rgb = torch.rand(B, 3, generator=g)
    mask = (torch.rand(B, 1, generator=g) > 0.3).float()
    lin = torch.nn.Linear(3, 3)
    with torch.no_grad():
        lin.weight.copy_(torch.eye(3)); lin.bias.zero_()

    class _R:
        pass
    # a "pose MLP" of a frame that no rank renders: it must come out of the gradient all-reduce WITHOUT a gradient, so that
    # an optimiser skips it as on one GPU (exp_runner.py:785-816 steps only the pose optimisers of the rendered frames)
    idle = torch.nn.Linear(2, 2)
    scene = dict(renderer=_R(), sdf_network=lin, deviation_network=torch.nn.Module(), color_network=torch.nn.Module(),
                 pose_network=idle)
    # single-process reference on the union batch
    ts1 = TrainStep(scene, mask_weight=5.0, group=None, optimizer=False, fused_loss=False)
    out = dict(color_fine=lin(color), weight_sum=wsum, gradient_error=torch.tensor(0.25))
    l1 = ts1.losses(out, rgb, mask)
    l1["loss"].backward()
    g_full = lin.weight.grad.clone()
    lin.weight.grad = None; lin.bias.grad = None
    # sharded
    ts2 = TrainStep(scene, mask_weight=5.0, group=dist.group.WORLD, optimizer=False, fused_loss=False)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    out2 = dict(color_fine=lin(color[sl]), weight_sum=wsum[sl], gradient_error=torch.tensor(0.25))
    l2 = ts2.losses(out2, rgb[sl], mask[sl])
    # colour and bce terms are partial sums over the shard with GLOBAL denominators
    parts = torch.stack([l2["color_loss"].detach(), l2["mask_loss"].detach()])
    dist.all_reduce(parts)
    (l2["color_loss"] + 5.0 * l2["mask_loss"]).backward()
    ts2.allreduce_grads()
    ok = (abs(parts[0].item() - l1["color_loss"].item()) < 1e-6 and abs(parts[1].item() - l1["mask_loss"].item()) < 1e-6
          and torch.allclose(lin.weight.grad, g_full, atol=1e-6)
          and idle.weight.grad is None and idle.bias.grad is None and lin.bias.grad is not None)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_ray_sharded_losses_and_grad_allreduce_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1), dict(ret)


# ---- grid partitioning of the dense SDF query (config C5) ----------------------------------------------------
def test_grid_slabs_cover_the_grid_exactly():
    from fmov_pose_b200.grid import slab_of
    for res in (1, 7, 64, 100, 512):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            sizes = []
            for r in range(world):
                first, n = slab_of(res, world, r)
                assert first == nxt and n >= 0
                nxt += n
                sizes.append(n)
            assert nxt == res and max(sizes) - min(sizes) <= 1
    assert slab_of(512, 8, 3) == (192, 64)


def _grid_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fmov_pose_b200.grid import gather_slabs, slab_of
    ok = True
    for res in (6, 7):                      # even split and ragged split (7 planes over 2 ranks)
        full = torch.arange(res ** 3, dtype=torch.float32)
        first, n = slab_of(res, world, rank)
        local = full[first * res * res: (first + n) * res * res].clone()
        ok = ok and torch.equal(gather_slabs(local, res, dist.group.WORLD), full)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_grid_slab_gather_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 30100 + (os.getpid() % 500)
    mp.spawn(_grid_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1), dict(ret)


# ---- learning-rate schedule + per-pose-MLP optimiser groups -----------------------------------
@pytest.mark.parametrize("tag", ["global", "wo_global"])
def test_lr_schedule_matches_the_reference_function(tag):
    """fixture = Runner.update_learning_rate exec'd from the reference source (oracle/gen_golden.py::lr_case)"""
    from fmov_pose_b200.train import LRSchedule, TrainStep
    d = load_golden("lr_schedule")
    sch = LRSchedule(learning_rate=5e-4, learning_rate_alpha=0.05, warm_up_end=50, end_iter=400, pose_lr=3e-4,
                     pose_alpha=0.5, max_pro_iteration=120, wo_global_conf=(tag == "wo_global"))

    class _Pose(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.pose_mlps = torch.nn.ModuleList([torch.nn.Linear(2, 2) for _ in range(3)])
            self.progress = torch.zeros(3)

        def step_progress(self, i):
            self.progress[i] += 1
            return self.progress[i]

    class _R:
        process_group = None
    scene = dict(renderer=_R(), sdf_network=torch.nn.Linear(3, 3), deviation_network=torch.nn.Module(),
                 color_network=torch.nn.Module(), pose_network=_Pose())
    ts = TrainStep(scene, lr=0.0, pose_lr=0.0, optimizer="torch")          # host-side LR logic; FlatAdam needs a GPU
    assert len(ts.optimizer.param_groups) == 4 and ts.pose_group_of == {0: 1, 1: 2, 2: 3}
    net, pose = [], []
    for it in d[tag + ".iters"]:
        ts.update_learning_rate(sch, int(it), {int(it) % 3})
        net.append(ts.optimizer.param_groups[0]["lr"])
        pose.append([ts.optimizer.param_groups[1 + k]["lr"] for k in range(3)])
    np.testing.assert_allclose(net, d[tag + ".net_lr"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(pose, d[tag + ".pose_lr"], rtol=1e-12, atol=0)


def test_fused_loss_function_glue_matches_the_torch_loss_block(monkeypatch):
    """train._FusedLossFn around fmov_loss_fwd_bwd: with the kernel replaced by a torch transcription of
    csrc/composite.cu::loss_fwd_bwd_kernel the function must reproduce exp_runner.py:562-599 (values and gradients),
    including the mask_weight scaling carried by the kernel's weight_sum gradient."""
    import torch.nn.functional as F

    from fmov_pose_b200 import _lib as L
    from fmov_pose_b200 import ops, train

    def kernel(color, wsum, true_rgb, mask, mask_sum, n_rays_global, mask_weight):
        m = (mask > 0.5).float() if mask_weight > 0 else torch.ones_like(mask)
        e = (color - true_rgb) * m
        g_color = torch.sign(e) * m / mask_sum
        x = wsum.clamp(1e-3, 1 - 1e-3)
        bce = -(m * torch.log(x) + (1 - m) * torch.log(1 - x))
        g_x = (-(m / x) + (1 - m) / (1 - x)) / n_rays_global * mask_weight
        g_w = torch.where((wsum >= 1e-3) & (wsum <= 1 - 1e-3), g_x, torch.zeros_like(g_x))
        partial = torch.cat([e.abs().sum(-1, keepdim=True) / mask_sum, bce / n_rays_global], dim=1)
        return partial, g_color, g_w

    monkeypatch.setattr(ops, "loss_fwd_bwd", kernel)
    monkeypatch.setattr(L, "f32c", lambda t: t.float().contiguous())
    g = torch.Generator().manual_seed(0)
    B = 97
    for mask_weight in (5.0, 0.0):
        color = torch.rand(B, 3, generator=g).requires_grad_(True)
        wsum = (torch.rand(B, 1, generator=g) * 1.2 - 0.1).requires_grad_(True)          # some outside the clip range
        rgb = torch.rand(B, 3, generator=g)
        raw_mask = torch.rand(B, 1, generator=g)
        mask = (raw_mask > 0.5).float() if mask_weight > 0 else torch.ones_like(raw_mask)
        mask_sum = mask.sum() + 1e-5
        col, bce = train._FusedLossFn.apply(color, wsum, rgb, mask, mask_sum, B, mask_weight)
        (col + bce * mask_weight).backward()
        got = (col.item(), bce.item(), color.grad.clone(), wsum.grad.clone() if wsum.grad is not None else torch.zeros_like(wsum))
        color.grad = None
        wsum.grad = None
        col_r = ((color - rgb) * mask).abs().sum() / mask_sum
        bce_r = F.binary_cross_entropy(wsum.clip(1e-3, 1 - 1e-3), mask, reduction="sum") / B
        (col_r + bce_r * mask_weight).backward()
        assert abs(got[0] - col_r.item()) < 1e-6 and abs(got[1] - bce_r.item()) < 1e-5
        np.testing.assert_allclose(got[2].numpy(), color.grad.numpy(), atol=1e-7)
        np.testing.assert_allclose(got[3].numpy(), wsum.grad.numpy(), atol=1e-6)


def test_stash_layout_follows_the_library_directory(monkeypatch):
    """fine.Stash carves ONE buffer by the library's stash directory: the tensors the forward kernel writes come first (a
    forward-only stash is a prefix of the full one), also when a forward tensor sits at the end of the id range
    (the ReLU sign words) or tensors have zero blocks (the q tiles, which the backward rebuilds instead of storing)."""
    from fmov_pose_b200 import _lib as L
    from fmov_pose_b200 import fine

    class FakeLib:
        def __init__(self, blocks, fwd):
            self.blocks, self.fwd = blocks, fwd

        def fmov_fine_stash_count(self):
            return len(self.blocks)

        def fmov_fine_stash_blocks(self, i):
            return self.blocks[i]

        def fmov_fine_stash_is_forward(self, i):
            return 1 if i in self.fwd else 0

    default = FakeLib([1] + [4] * 8 + [4] + [4] * 8 + [1] + [4] * 4 + [4] * 4 + [4, 1] + [4] * 8 + [4] * 8 + [4] * 8, set(range(23)))
    variant = FakeLib(default.blocks[:37] + [0] * 8 + default.blocks[45:] + [1], set(range(23)) | {53})
    for lib in (default, variant):
        monkeypatch.setattr(L, "lib", lambda lib=lib: lib)
        fine.Stash._pool.clear()
        P = 128 * 5 + 3                                     # 6 tiles + the padding tile of the CTA-pair engine
        NT = 7
        full = fine.Stash(P, torch.device("cpu"), with_backward=True)
        n = len(lib.blocks)
        assert full.nt == NT and len(full.tensors) == n
        spans = []
        for i, t in enumerate(full.tensors):
            assert t is not None and t.numel() == lib.blocks[i] * NT * 16384, i
            off = t.data_ptr() - full.buf.data_ptr()
            spans.append((off, off + t.numel(), i))
        used = sorted(s for s in spans if s[1] > s[0])
        assert all(a[1] == b[0] for a, b in zip(used, used[1:])) and used[0][0] == 0          # contiguous, no overlap
        assert used[-1][1] == sum(lib.blocks) * NT * 16384 == full.buf.numel()
        n_fwd_bytes = sum(lib.blocks[i] for i in lib.fwd) * NT * 16384
        assert all((s[1] <= n_fwd_bytes) == (s[2] in lib.fwd) for s in used)                  # forward tensors = prefix
        full.buf = None                                      # do not recycle: the next stash must size itself
        fwd_only = fine.Stash(P, torch.device("cpu"), with_backward=False)
        assert fwd_only.buf.numel() == n_fwd_bytes
        assert all((t is not None) == (i in lib.fwd) for i, t in enumerate(fwd_only.tensors) if lib.blocks[i] > 0)
        fwd_only.ensure_backward(P, torch.device("cpu"))
        assert all(t is not None for t in fwd_only.tensors) and fwd_only.buf.numel() == sum(lib.blocks) * NT * 16384
        fwd_only.buf = None
    fine.Stash._pool.clear()


def test_fused_ray_path_autograd_glue(monkeypatch):
    """TrainStep(fused_rays=True): _RayGenNearFarFn (near / far from the ray-generation kernel) and _CoarseZFn (coarse z through
    fmov_sample_coarse with a closed-form backward).  With the kernels replaced by torch transcriptions of what they
    compute, values and gradients must equal the reference formulas (models/dataset.py:656-671, 835-842;
    models/renderer.py:389-390, 403-405)."""
    from fmov_pose_b200 import ops
    from fmov_pose_b200.models import dataset as D
    from fmov_pose_b200.models import renderer as R
    from oracle import neus_oracle as O

    def raygen_fwd(mode, intr_inv, px, py, c2w34=None, **kw):
        o, d = O.gen_rays(c2w34, intr_inv, px, py)
        near, far = O.near_far_from_sphere(o, d)
        return o.contiguous(), d, near, far, c2w34

    def raygen_bwd(intr_inv, px, py, rays_o, rays_d, g_o, g_d, g_near, g_far):
        # what fmov_raygen_bwd returns: d(loss)/d(pose) given the four upstream gradients (None = zero)
        with torch.enable_grad():
            q = saved["pose"].clone().requires_grad_(True)
            o, d = O.gen_rays(q, intr_inv, px, py)
            near, far = O.near_far_from_sphere(o, d)
            tot = 0
            for t_, g_ in ((o, g_o), (d, g_d), (near, g_near), (far, g_far)):
                if g_ is not None:
                    tot = tot + (t_ * g_).sum()
            return torch.autograd.grad(tot, q)[0]

    saved = {}
    monkeypatch.setattr(ops, "raygen_fwd", raygen_fwd)
    monkeypatch.setattr(ops, "raygen_bwd", raygen_bwd)
    g = torch.Generator().manual_seed(1)
    pose = torch.cat([O.rodrigues_exp(torch.randn(1, 3, generator=g) * 0.2)[0], torch.tensor([[0.1], [0.2], [-3.0]])], 1)
    pose = pose.requires_grad_(True)
    saved["pose"] = pose
    Kinv = torch.inverse(torch.tensor([[60.0, 0, 32.0], [0, 60.0, 24.0], [0, 0, 1.0]]))
    px, py = torch.randint(0, 64, [50], generator=g), torch.randint(0, 48, [50], generator=g)
    w = [torch.randn(50, 3, generator=g), torch.randn(50, 3, generator=g), torch.randn(50, 1, generator=g),
         torch.randn(50, 1, generator=g)]
    o, d, near, far = D._RayGenNearFarFn.apply(pose, Kinv, px, py)
    ((o * w[0]).sum() + (d * w[1]).sum() + (near * w[2]).sum() + (far * w[3]).sum()).backward()
    got = pose.grad.clone()
    pose.grad = None
    o2, d2 = O.gen_rays(pose, Kinv, px, py)
    n2, f2 = O.near_far_from_sphere(o2, d2)
    ((o2 * w[0]).sum() + (d2 * w[1]).sum() + (n2 * w[2]).sum() + (f2 * w[3]).sum()).backward()
    np.testing.assert_allclose(near.detach().numpy(), n2.detach().numpy(), atol=1e-6)
    np.testing.assert_allclose(got.numpy(), pose.grad.numpy(), rtol=1e-5, atol=1e-5)

    # coarse z
    def sample_coarse(near, far, t_rand, n_samples, z_stride):
        lin = torch.linspace(0.0, 1.0, n_samples)
        z = near + (far - near) * lin[None, :]
        return z if t_rand is None else z + (t_rand - 0.5) * 2.0 / n_samples

    monkeypatch.setattr(ops, "sample_coarse", sample_coarse)
    for t_rand in (None, torch.rand(50, 1, generator=g)):
        nr = (torch.rand(50, 1, generator=g) + 1.5).requires_grad_(True)
        fr = (nr.detach() + 2.0).requires_grad_(True)
        gz = torch.randn(50, 32, generator=g)
        z = R._CoarseZFn.apply(nr, fr, t_rand, 32)
        (z * gz).sum().backward()
        got = (nr.grad.clone(), fr.grad.clone())
        nr.grad = fr.grad = None
        z2 = sample_coarse(nr, fr, t_rand, 32, 32)
        (z2 * gz).sum().backward()
        np.testing.assert_allclose(z.detach().numpy(), z2.detach().numpy(), atol=1e-7)
        np.testing.assert_allclose(got[0].numpy(), nr.grad.numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(got[1].numpy(), fr.grad.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("level", [1, 2, 4])
def test_ray_dataset_full_frame_grid_at_every_resolution_level(monkeypatch, level):
    """RayDataset.gen_rays_at (models/dataset.py:547-576): pixel grid linspace(0, W-1, W//l) — integer centres at l = 1,
    sub-pixel otherwise — ray layout [H/l, W/l, 3] and the mask lookup at the truncated pixel; the kernel is replaced by the
    oracle's formula so that the host logic runs on the CPU."""
    from fmov_pose_b200.models import dataset as D
    from oracle import neus_oracle as O

    class _Fn:
        @staticmethod
        def apply(pose34, intr_inv, px, py):
            assert px.dtype == (torch.int64 if level == 1 else torch.float32) and px.is_contiguous()
            return O.gen_rays(pose34, intr_inv, px, py)

    monkeypatch.setattr(D, "_RayGenFn", _Fn)
    g = torch.Generator().manual_seed(0)
    H, W = 24, 32
    ds = D.RayDataset(torch.rand(2, H, W, 3, generator=g), (torch.rand(2, H, W, 3, generator=g) > 0.5).float(),
                      [[30.0, 0, 16.0], [0, 30.0, 12.0], [0, 0, 1.0]], device="cpu")
    pose = torch.cat([O.rodrigues_exp(torch.tensor([[0.1, -0.2, 0.05]]))[0], torch.tensor([[0.0], [0.1], [-3.0]])], 1)
    o, v, m = ds.gen_rays_at(1, resolution_level=level, pose=pose, with_mask=True)
    Hl, Wl = H // level, W // level
    assert o.shape == (Hl, Wl, 3) and v.shape == (Hl, Wl, 3) and m.shape == (Hl, Wl)
    tx, ty = torch.linspace(0, W - 1, Wl), torch.linspace(0, H - 1, Hl)
    gx, gy = torch.meshgrid(tx, ty, indexing="ij")
    o_ref, v_ref = O.gen_rays(pose, ds.intrinsics_all_inv[1], gx.reshape(-1), gy.reshape(-1))
    np.testing.assert_allclose(v.transpose(0, 1).reshape(-1, 3).numpy(), v_ref.numpy(), atol=1e-6)
    np.testing.assert_allclose(o.transpose(0, 1).reshape(-1, 3).numpy(), o_ref.numpy(), atol=1e-6)
    np.testing.assert_array_equal(m.numpy(), ds.masks[1][(gy.long(), gx.long())][..., 0].transpose(0, 1).numpy())
