"""Synthetic scene of SURVEY.md §8d / BASELINE.md §4: reference-shaped networks (confs/ho3d_*.conf
`model{}` kwargs), N synthetic 640x480 frames with a disc mask, K = [[600,0,320],[0,600,240],[0,0,1]],
initial poses t = (0,0,-3) looking at the origin, reference-initialised pose modules.
Used by bench.py, __graft_entry__.smoke() and the tests (there is no network for real datasets)."""
import numpy as np
import torch

SDF_KW = dict(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
              geometric_init=True, weight_norm=True)          # confs/ho3d_virtual.conf:79-90
COL_KW = dict(d_feature=256, mode="idr", d_in=9, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
              multires_view=4, squeeze_out=True)              # confs/ho3d_virtual.conf:96-106
INTRINSICS = [[600.0, 0.0, 320.0], [0.0, 600.0, 240.0], [0.0, 0.0, 1.0]]


def make_frames(n_images=20, H=480, W=640, seed=2024):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(n_images, H, W, 3, generator=g)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    disc = (((xs - W // 2) ** 2 + (ys - H // 2) ** 2) < 150 ** 2).float()
    masks = disc[None, :, :, None].repeat(n_images, 1, 1, 3)
    return images, masks


def make_init_poses(n_images=20):
    init = torch.eye(4).repeat(n_images, 1, 1)
    init[:, :3, 3] = torch.tensor([0.0, 0.0, -3.0])
    return init


def build_scene(device="cuda", n_images=20, n_samples=64, n_importance=64, up_sample_steps=4, pose_type="seg",
                seed=2024, H=480, W=640):
    """-> dict(renderer, sdf_network, color_network, deviation_network, pose_network, dataset)"""
    from .models.barf_fields import BarfRenderingNetwork, BarfSDFNetwork
    from .models.dataset import RayDataset
    from .models.fields import SingleVarianceNetwork
    from .models.picture_pose import LearnPoseGF, SegLearnPose
    from .models.renderer import NeuSRenderer
    torch.manual_seed(seed)          # exp_runner.py:29-30
    np.random.seed(seed)
    init = make_init_poses(n_images)
    sdf = BarfSDFNetwork(init.clone(), n_images=n_images, **SDF_KW)
    col = BarfRenderingNetwork(**COL_KW)
    dev_net = SingleVarianceNetwork(0.3)
    if pose_type == "seg":           # ho3d_virtual.conf: pose_type = seg, image_interval = 1, emphasize_rot
        pose = SegLearnPose(n_images, 1, init_c2w=init.clone(), emphasize_rot=True)
        pose.initialized_flag.data[:] = True
        pose._flags_host = None
    elif pose_type == "gf":          # ho3d_barf.conf / ho3d_global_womask.conf
        pose = LearnPoseGF(n_images, init_c2w=init.clone())
    elif pose_type == "se3":         # BARF se3_refine (exp_runner.py:419-424)
        pose = None
        with torch.no_grad():
            sdf.se3_refine.weight.normal_(0.0, 0.05)
    else:
        raise ValueError(pose_type)
    images, masks = make_frames(n_images, H, W, seed)
    sdf, col, dev_net = sdf.to(device), col.to(device), dev_net.to(device)
    if pose is not None:
        pose = pose.to(device)
    ds = RayDataset(images, masks, INTRINSICS, device=device)
    rend = NeuSRenderer(None, sdf, dev_net, col, n_samples, n_importance, 0, up_sample_steps, 1.0)
    return dict(renderer=rend, sdf_network=sdf, color_network=col, deviation_network=dev_net, pose_network=pose,
                dataset=ds, pose_type=pose_type, n_images=n_images)
