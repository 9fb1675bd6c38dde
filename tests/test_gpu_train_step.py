"""Whole train iteration from pose parameters (pose module -> ray-gen -> render -> loss -> backward) on the
GPU against the reference's golden fixtures, for the three pose paths of SURVEY.md §8 a1/a2."""
import numpy as np
import pytest
import torch

from oracle import neus_oracle as O
from tests._util import load_golden, t
from tests.test_gpu_render import build_from_fixture, rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pose_module(d):
    from fmov_pose_b200.models.picture_pose import LearnPoseGF
    kind = str(d["pose_kind"])
    init = t(d, "init_c2w")
    m = LearnPoseGF(6, init_c2w=init.clone(), emphasize_rot=(kind == "seg"))
    sd = {k[5:]: torch.from_numpy(np.asarray(v)) for k, v in d.items() if k.startswith("pose.")}
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


@pytest.mark.parametrize("name", ["full_6464_gf", "full_3200_seg"])
def test_train_step_pose_gradients_vs_reference(name):
    from fmov_pose_b200.models.dataset import _RayGenFn
    d = load_golden(name)
    rend, sdf_net, col_net, var_net = build_from_fixture(d)
    pm = _pose_module(d)
    img_id = int(d["img_id"])
    pose = pm(img_id)[:3]
    np.testing.assert_allclose(pose.detach().cpu().numpy(), d["pose"], atol=3e-6)
    px, py = torch.from_numpy(d["px"]).to(DEV), torch.from_numpy(d["py"]).to(DEV)
    rays_o, rays_d = _RayGenFn.apply(pose, t(d, "intr_inv").to(DEV).contiguous(), px, py)
    np.testing.assert_allclose(rays_d.detach().cpu().numpy(), d["rays_d"], atol=3e-6)
    near, far = O.near_far_from_sphere(rays_o, rays_d)
    out = rend.render(rays_o, rays_d, near, far, cos_anneal_ratio=float(d["cos_anneal"]), t_rand=t(d, "t_rand").to(DEV))
    ls = O.loss_block(out, t(d, "true_rgb").to(DEV), t(d, "mask").to(DEV), 0.1, float(d["mask_weight"]))
    ls["loss"].backward()
    trainable = [p for p in pm.parameters() if p.requires_grad]
    assert len(trainable) == sum(1 for k in d if k.startswith("grad.pose_param"))
    for i, p in enumerate(trainable):
        ref = d[f"grad.pose_param{i}"]
        e = rel(p.grad.cpu().numpy(), ref)
        # own importance sampling + fp16 tensor-core operands vs the fp32 reference: a few per cent
        assert e <= 5e-2, (i, e)


def test_barf_se3_pose_path():
    """se3_refine -> se3_to_SE3 -> compose(noise_poses) (exp_runner.py:419-424) through fmov_pose_fwd/bwd."""
    from fmov_pose_b200.models import camera
    d = load_golden("small_6464_se3_white")
    se3 = t(d, "sdf.se3_refine.weight").to(DEV).requires_grad_(True)
    noise = t(d, "sdf.noise_poses").to(DEV)
    img_id = int(d["img_id"])
    pose = camera.barf_pose(se3[img_id], noise[img_id, :3, :])
    np.testing.assert_allclose(pose.detach().cpu().numpy(), d["pose"], atol=3e-6)
    g = t(d, "grad.pose").to(DEV)
    pose.backward(g)
    ref = d["grad.pose_param0"]
    assert rel(se3.grad.cpu().numpy(), ref) <= 2e-3
    # batched host-side torch mirror agrees too (all frames at once, as the reference computes them)
    all_p = camera.pose.compose([camera.lie.se3_to_SE3(se3.detach()), noise[:, :3, :]])
    np.testing.assert_allclose(all_p.cpu().numpy(), d["pose_all"], atol=3e-6)


@pytest.mark.parametrize("pose_type", ["seg", "se3"])
def test_graphed_step_equals_eager_step(pose_type):
    """CUDA-graph replay of the train step (train.GraphedTrainStep) == the eager launches of the same step:
    same losses and the same parameters after several Adam updates, with the frame index, pixels, jitter and
    learning rate as replay-time inputs."""
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import GraphedTrainStep, TrainStep
    B, n_steps = 256, 5
    g = torch.Generator().manual_seed(11)
    px = torch.randint(200, 440, [n_steps, B], generator=g)
    py = torch.randint(120, 360, [n_steps, B], generator=g)
    tr = torch.rand(n_steps, B, 1, generator=g)
    imgs = [3, 3, 5, 3, 5]
    lrs = [5e-4, 4e-4, 3e-4, 2e-4, 1e-4]
    res = []
    for graphed in (False, True):
        sc = synthetic.build_scene(device=DEV, n_images=8, n_samples=16, n_importance=16, up_sample_steps=2,
                                   pose_type=pose_type)
        ts = TrainStep(sc, mask_weight=5.0, capturable=graphed)
        gts = GraphedTrainStep(ts, B) if graphed else None
        losses = []
        for i in range(n_steps):
            ts.set_lr(lrs[i])
            if graphed:
                ls, _ = gts.step(imgs[i], px[i].pin_memory(), py[i].pin_memory(), tr[i].pin_memory())
            else:
                ls, _ = ts.step(imgs[i], B, pixels=(px[i].to(DEV), py[i].to(DEV)), t_rand=tr[i].to(DEV))
            losses.append(float(ls["loss"]))
        res.append((losses, [p.detach().clone() for p in ts.all_params]))
        if graphed:
            assert gts.launches_per_step > 10 and len(gts.graphs) == (2 if pose_type == "seg" else 1)
    (l0, p0), (l1, p1) = res
    np.testing.assert_allclose(l1, l0, rtol=2e-3)
    moved = 0.0
    for a, b in zip(p0, p1):
        # Adam's sign-like first steps amplify fp16 rounding noise on near-zero gradients: compare on the scale of
        # the accumulated update (5 steps x lr), not bitwise
        assert float((a - b).abs().max()) <= 3e-3, float((a - b).abs().max())
        assert float((a - b).abs().mean()) <= 1e-4, float((a - b).abs().mean())
        moved = max(moved, float(a.abs().max()))
    assert moved > 0


@pytest.mark.parametrize("kw", [
    dict(B=77, n_samples=16, n_importance=32, up_sample_steps=2, white_bkgd=True, cos_anneal_ratio=0.3, mask_weight=0.0),
    dict(B=1, n_samples=64, n_importance=64, up_sample_steps=4),
    dict(B=129, n_samples=48, n_importance=16, up_sample_steps=1, white_bkgd=True, mask_weight=1.0),
    dict(B=200, n_samples=32, n_importance=64, up_sample_steps=4, cos_anneal_ratio=0.0),
])
def test_train_iteration_vs_oracle_ragged_and_variants(kw):
    """Edge cases of the hot path against the oracle on the same z samples (north_star tolerances, asserted inside
    selfcheck.smoke): ragged ray counts (1, 77, 129: partial 128-point tiles), sample counts other than 64+64, one
    up-sample round, white background (exp_runner.py:556), cos-anneal ratios 0 / 0.3 (renderer.py:300-305),
    mask_weight 0 (mask -> ones, exp_runner.py:564-566)."""
    from fmov_pose_b200 import selfcheck
    assert selfcheck.smoke(verbose=False, **kw)


@pytest.mark.parametrize("cfg", [dict(n_samples=16, n_importance=16, B=700, mb=256, pose_type="seg"),
                                 dict(n_samples=32, n_importance=0, B=512, mb=128, pose_type="se3")])
def test_micro_batched_step_equals_one_shot_step(cfg):
    """TrainStep(micro_batch=...) — the path for ray batches whose stash does not fit HBM (config C3 at N = 2 / 4):
    fine stage + backward per micro-batch against whole-batch normalisers == the one-shot step (losses and every
    network / pose gradient), with a ragged last micro-batch and with n_importance == 0 (z keeps its link to the pose)."""
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import TrainStep
    B = cfg["B"]
    g = torch.Generator().manual_seed(4)
    px = torch.randint(150, 490, [B], generator=g).to(DEV)
    py = torch.randint(70, 410, [B], generator=g).to(DEV)
    tr = torch.rand(B, 1, generator=g).to(DEV)
    res = []
    for mb in (None, cfg["mb"]):
        sc = synthetic.build_scene(device=DEV, n_images=6, n_samples=cfg["n_samples"], n_importance=cfg["n_importance"],
                                   up_sample_steps=2, pose_type=cfg["pose_type"])
        ts = TrainStep(sc, mask_weight=5.0, optimizer=False)
        ls, out = ts.forward_backward(3, B, pixels=(px, py), t_rand=tr, micro_batch=mb)
        res.append((ls, out, [None if p.grad is None else p.grad.detach().clone() for p in ts.all_params]))
    (la, oa, ga), (lb, ob, gb) = res
    for k in ("loss", "color_loss", "eikonal_loss", "mask_loss"):
        np.testing.assert_allclose(float(lb[k].detach()), float(la[k].detach()), rtol=1e-4, err_msg=k)
    # identical samples for n_importance > 0 (both paths sample from the kernel's near / far); with n_importance == 0 the
    # micro-batches rebuild z from torch's near / far (ulps from the kernel's), hence not bitwise
    np.testing.assert_allclose(ob["color_fine"].cpu().numpy(), oa["color_fine"].detach().cpu().numpy(), atol=2e-4)
    n = 0
    for a, b in zip(ga, gb):
        assert (a is None) == (b is None)
        if a is not None and float(a.abs().max()) > 0:
            # per-micro-batch fp16 loss scales differ from the one-shot scale: fp16 rounding noise, not bias
            assert rel(b.cpu().numpy(), a.cpu().numpy()) <= 5e-3, rel(b.cpu().numpy(), a.cpu().numpy())
            n += 1
    assert n > 40


def test_training_trajectory_tracks_the_oracle_over_adam_steps():
    """30 consecutive Adam iterations (render -> losses -> backward -> Adam) of this path and of the oracle (the
    reference's PyTorch fp32 autograd formulation, run on the same GPU) from identical initial weights, with identical
    pixels / jitter and the oracle fed this path's z samples: the loss trajectories must stay together (fp16 operand
    rounding does not accumulate into drift) and the loss must go down."""
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import TrainStep
    B, n_steps, lr = 256, 30, 5e-4
    sc = synthetic.build_scene(device=DEV, n_images=4, n_samples=32, n_importance=32, up_sample_steps=2, pose_type="seg")
    for m in sc["pose_network"].pose_mlps:
        m.disable_grad()                       # fixed poses: the trajectory below is the networks' (Adam over sdf+colour+variance)
    ts = TrainStep(sc, mask_weight=5.0, lr=lr)
    sdf_p = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in sc["sdf_network"].named_parameters()}
    col_p = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in sc["color_network"].named_parameters()}
    var = sc["deviation_network"].variance.detach().clone().requires_grad_(True)
    leaves = [p for p in list(sdf_p.values()) + [var] + list(col_p.values()) if p.requires_grad]
    opt = torch.optim.Adam(leaves, lr=lr)
    ds = sc["dataset"]
    g = torch.Generator().manual_seed(21)
    mine, ref = [], []
    for it in range(n_steps):
        img = it % 4
        px = torch.randint(170, 470, [B], generator=g).to(DEV)
        py = torch.randint(90, 390, [B], generator=g).to(DEV)
        tr = torch.rand(B, 1, generator=g).to(DEV)
        ls, out = ts.step(img, B, pixels=(px, py), t_rand=tr)
        mine.append(float(ls["loss"].detach()))
        with torch.device(DEV):                # the oracle's factory calls follow the default device
            pose = ts.pose_of(img).detach()
            ro, rd = O.gen_rays(pose, ds.intrinsics_all_inv[img], px, py)
            nr, fr = O.near_far_from_sphere(ro, rd)
            r = O.render(sdf_p, col_p, var, ro, rd, nr, fr, n_samples=32, n_importance=32, up_sample_steps=2,
                         cos_anneal_ratio=1.0, z_vals=out["z_vals"].detach())
            rl = O.loss_block(r, ds.images[img][(py, px)], ds.masks[img][(py, px)][:, :1], 0.1, 5.0)
            opt.zero_grad()
            rl["loss"].backward()
            opt.step()
        ref.append(float(rl["loss"].detach()))
    mine, ref = np.array(mine), np.array(ref)
    # identical to ~1e-4 for the first iterations; Adam's normalised updates then amplify the fp16-operand rounding
    # noise while the loss falls steeply (measured: <= 5 % apart after 30 iterations, 1.5 % on average)
    relerr = np.abs(mine - ref) / np.abs(ref)
    assert relerr[:4].max() <= 2e-3 and relerr.max() <= 0.10 and relerr.mean() <= 0.03, (mine, ref)
    assert mine[-5:].mean() < 0.8 * mine[:5].mean(), mine            # it trains
    # the weights stayed together as well (relative to the distance travelled from the initial weights)
    w_mine = sc["sdf_network"].lin4.weight_v.detach()
    w_ref = sdf_p["lin4.weight_v"].detach()
    assert float((w_mine - w_ref).abs().mean()) <= 0.1 * lr * n_steps


def test_two_frame_maintain_shape_iteration_eager_and_graphed():
    """The shipped confs' maintain_shape iteration (exp_runner.py:512-548): B rays of the current frame + B rays of an
    earlier frame in one render.  (a) with the same frame twice it equals a single-frame step on the concatenated pixels;
    (b) with two frames both pose MLPs (and only those) receive gradients; (c) graph replay == eager."""
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import GraphedTrainStep, TrainStep
    B = 128
    g = torch.Generator().manual_seed(17)
    px = torch.randint(200, 440, [2 * B], generator=g).to(DEV)
    py = torch.randint(120, 360, [2 * B], generator=g).to(DEV)
    tr = torch.rand(2 * B, 1, generator=g).to(DEV)
    mk = lambda: synthetic.build_scene(device=DEV, n_images=6, n_samples=32, n_importance=0, pose_type="seg")
    # (a)
    ts1 = TrainStep(mk(), mask_weight=5.0, optimizer=False)
    l1, _ = ts1.forward_backward(2, 2 * B, pixels=(px, py), t_rand=tr)
    ts2 = TrainStep(mk(), mask_weight=5.0, optimizer=False)
    l2, o2 = ts2.forward_backward(2, B, pixels=(px[:B], py[:B]), t_rand=tr, additional_img_id=2, add_pixels=(px[B:], py[B:]))
    assert o2["color_fine"].shape[0] == 2 * B
    np.testing.assert_allclose(float(l2["loss"].detach()), float(l1["loss"].detach()), rtol=1e-5)
    for a, b in zip(ts1.all_params, ts2.all_params):
        assert (a.grad is None) == (b.grad is None)
        if a.grad is not None and float(a.grad.abs().max()) > 0:
            assert rel(b.grad.cpu().numpy(), a.grad.cpu().numpy()) <= 2e-3
    # (b)
    sc = mk()
    ts3 = TrainStep(sc, mask_weight=5.0, optimizer=False)
    ts3.forward_backward(4, B, pixels=(px[:B], py[:B]), t_rand=tr, additional_img_id=1, add_pixels=(px[B:], py[B:]))
    touched = [k for k, m in enumerate(sc["pose_network"].pose_mlps)
               if any(p.grad is not None and float(p.grad.abs().max()) > 0 for p in m.parameters())]
    assert touched == [1, 4], touched
    # (c)
    res = []
    for graphed in (False, True):
        sc = mk()
        ts = TrainStep(sc, mask_weight=5.0, capturable=graphed)
        gts = GraphedTrainStep(ts, B, two_frames=True) if graphed else None
        losses = []
        for it, (fa, fb) in enumerate([(4, 1), (3, 0), (4, 1)]):
            if graphed:
                ls, _ = gts.step(fa, px[:B], py[:B], tr, add_img_id=fb, add_px=px[B:], add_py=py[B:])
            else:
                ls, _ = ts.step(fa, B, pixels=(px[:B], py[:B]), t_rand=tr, additional_img_id=fb, add_pixels=(px[B:], py[B:]))
            losses.append(float(ls["loss"].detach()))
        res.append(losses)
        if graphed:
            assert len(gts.graphs) == 2
    np.testing.assert_allclose(res[1], res[0], rtol=2e-3)


def test_bench_configuration_8192_rays_vs_fp32_oracle_on_the_gpu():
    """Parity AT THE BENCHMARKED CONFIGURATION (bench.py defaults: 8192 rays x (64+64) samples, SegLearnPose, mask_weight 5):
    the oracle — fp32 PyTorch autograd with the double backward through sdf_network.gradient — runs on the same device in
    2048-ray chunks against whole-batch normalisers, on the z samples the CUDA path drew.  north_star tolerances: colour
    <= 2e-3, SDF <= 1e-3, every parameter / pose gradient rel <= 1e-2."""
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import TrainStep
    import torch.nn.functional as F
    B, n, m, up, img = 8192, 64, 64, 4, 3
    scene = synthetic.build_scene(device=DEV, n_samples=n, n_importance=m, up_sample_steps=up, pose_type="seg")
    ts = TrainStep(scene, igr_weight=0.1, mask_weight=5.0, optimizer=False)
    g = torch.Generator().manual_seed(1234)
    px = torch.randint(140, 500, [B], generator=g).to(DEV)
    py = torch.randint(60, 420, [B], generator=g).to(DEV)
    tr = torch.rand(B, 1, generator=g).to(DEV)
    ls, out = ts.forward_backward(img, B, pixels=(px, py), t_rand=tr)
    torch.cuda.synchronize()
    ds = scene["dataset"]
    with torch.device(DEV):          # the oracle's factory calls follow the default device
        leaf = lambda v: v.detach().clone().requires_grad_(v.requires_grad)
        sdf_p = {k: leaf(v) for k, v in scene["sdf_network"].named_parameters()}
        col_p = {k: leaf(v) for k, v in scene["color_network"].named_parameters()}
        var = leaf(scene["deviation_network"].variance)
        pose = ts.pose_of(img).detach().clone().requires_grad_(True)
        rgb = ds.images[img][(py, px)].to(DEV)
        mask = (ds.masks[img][(py, px)][:, :1].to(DEV) > 0.5).float()
        z = out["z_vals"].detach()
        S = z.shape[1]
        sd = 2.0 / n
        with torch.no_grad():
            ro, rd = O.gen_rays(pose, ds.intrinsics_all_inv[img], px, py)
            dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], sd)], -1)
            pts = ro[:, None] + rd[:, None] * (z + 0.5 * dists)[..., None]
            relax_total = (torch.linalg.norm(pts, dim=-1) < 1.2).float().sum()
        mask_sum = mask.sum() + 1e-5
        tot = torch.zeros(4, device=DEV)
        col_err = sdf_err = 0.0
        for i in range(0, B, 2048):
            sl = slice(i, i + 2048)
            ro, rd = O.gen_rays(pose, ds.intrinsics_all_inv[img], px[sl], py[sl])
            nr, fr = O.near_far_from_sphere(ro, rd)
            ref = O.render(sdf_p, col_p, var, ro, rd, nr, fr, n_samples=n, n_importance=m, up_sample_steps=up,
                           cos_anneal_ratio=1.0, z_vals=z[sl])
            relax = (torch.linalg.norm(ref["pts"].detach().reshape(-1, S, 3), dim=-1) < 1.2).float()
            eik = (relax * (torch.linalg.norm(ref["gradients"], dim=-1) - 1.0) ** 2).sum() / (relax_total + 1e-5)
            col = ((ref["color_fine"] - rgb[sl]) * mask[sl]).abs().sum() / mask_sum
            bce = F.binary_cross_entropy(ref["weight_sum"].clip(1e-3, 1 - 1e-3), mask[sl], reduction="sum") / B
            loss = col + 0.1 * eik + 5.0 * bce
            loss.backward()
            tot += torch.stack([loss.detach(), col.detach(), eik.detach(), bce.detach()])
            col_err = max(col_err, float((out["color_fine"][sl] - ref["color_fine"]).abs().max()))
            sdf_err = max(sdf_err, float((out["sdf"].reshape(B, S)[sl] - ref["sdf"].reshape(-1, S)).abs().max()))
            del ref, loss
    assert col_err <= 2e-3, col_err
    assert sdf_err <= 1e-3, sdf_err
    for k, v in zip(("loss", "color_loss", "eikonal_loss", "mask_loss"), tot.tolist()):
        assert abs(float(ls[k]) - v) <= 2e-3 * max(1.0, abs(v)), (k, float(ls[k]), v)
    worst = 0.0
    for net, pd in ((scene["sdf_network"], sdf_p), (scene["color_network"], col_p)):
        for k, p in net.named_parameters():
            if pd[k].grad is None or p.grad is None:
                assert pd[k].grad is None or float(pd[k].grad.abs().max()) == 0.0 or p.grad is not None, k
                continue
            e = float((p.grad.double() - pd[k].grad.double()).norm() / (pd[k].grad.double().norm() + 1e-30))
            worst = max(worst, e)
            assert e <= 1e-2, (k, e)
    e = float((scene["deviation_network"].variance.grad - var.grad).abs() / var.grad.abs())
    assert e <= 1e-2, ("variance", e)
    # pose: the oracle's d loss / d c2w pushed through the pose module == the step's pose-parameter gradients
    got = {k: p.grad.detach().clone() for k, p in scene["pose_network"].named_parameters() if p.grad is not None}
    assert got, "the rendered frame's pose MLP must receive gradients"
    for p in scene["pose_network"].parameters():
        p.grad = None
    ts.pose_of(img).backward(pose.grad)
    for k, p in scene["pose_network"].named_parameters():
        if k in got:
            e = float((got[k].double() - p.grad.double()).norm() / (p.grad.double().norm() + 1e-30))
            assert e <= 1e-2, ("pose", k, e)
    print(f"bench-config parity: colour {col_err:.2e} sdf {sdf_err:.2e} worst param-grad rel {worst:.2e}")
