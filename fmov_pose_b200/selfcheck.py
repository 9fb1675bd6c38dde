"""smoke(): one small train iteration of the hot path on cuda:0, checked against the CPU oracle.
(The oracle is imported here only as the checker — see oracle/neus_oracle.py header.)"""
import numpy as np
import torch


def smoke(B=256, verbose=True, n_samples=64, n_importance=64, up_sample_steps=4, white_bkgd=False,
          cos_anneal_ratio=1.0, mask_weight=5.0):
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import TrainStep
    from oracle import neus_oracle as O
    dev = torch.device("cuda:0")
    scene = synthetic.build_scene(device=dev, n_images=4, n_samples=n_samples, n_importance=n_importance,
                                 up_sample_steps=up_sample_steps, pose_type="seg", H=120, W=160)
    # frames are 160x120 here: scale the intrinsics accordingly
    ds = scene["dataset"]
    K = torch.tensor([[150.0, 0, 80.0], [0, 150.0, 60.0], [0, 0, 1.0]])
    ds.intrinsics_all_inv = torch.linalg.inv(K)[None].repeat(4, 1, 1).contiguous().to(dev)
    ts = TrainStep(scene, igr_weight=0.1, mask_weight=mask_weight, optimizer=False)
    ts.background_rgb = torch.ones(1, 3, device=dev) if white_bkgd else None
    g = torch.Generator().manual_seed(0)
    px = torch.randint(30, 130, [B], generator=g).to(dev)
    py = torch.randint(10, 110, [B], generator=g).to(dev)
    t_rand = torch.rand(B, 1, generator=g).to(dev)
    ls, out = ts.forward_backward(1, B, pixels=(px, py), t_rand=t_rand, cos_anneal_ratio=cos_anneal_ratio)
    torch.cuda.synchronize()
    # oracle on the same z samples
    cpu = lambda x: x.detach().cpu()
    sdf_p = {k: cpu(v).clone().requires_grad_(v.requires_grad) for k, v in scene["sdf_network"].named_parameters()}
    col_p = {k: cpu(v).clone().requires_grad_(v.requires_grad) for k, v in scene["color_network"].named_parameters()}
    var = cpu(scene["deviation_network"].variance).clone().requires_grad_(True)
    pose = cpu(ts.pose_of(1))
    ro, rd = O.gen_rays(pose, cpu(ds.intrinsics_all_inv[1]), cpu(px), cpu(py))
    nr, fr = O.near_far_from_sphere(ro, rd)
    ref = O.render(sdf_p, col_p, var, ro, rd, nr, fr, n_samples=n_samples, n_importance=n_importance,
                   up_sample_steps=up_sample_steps, cos_anneal_ratio=cos_anneal_ratio, z_vals=cpu(out["z_vals"]),
                   background_rgb=torch.ones(1, 3) if white_bkgd else None)
    data_rgb = cpu(ds.images[1][(py, px)])
    mask = cpu(ds.masks[1][(py, px)])[:, :1]
    rl = O.loss_block(ref, data_rgb, mask, 0.1, mask_weight)
    rl["loss"].backward()
    col_err = (cpu(out["color_fine"]) - ref["color_fine"].detach()).abs().max().item()
    sdf_err = (cpu(out["sdf"]) - ref["sdf"].detach()).abs().max().item()
    g_gpu = cpu(scene["sdf_network"].lin4.weight_v.grad).double()
    g_ref = sdf_p["lin4.weight_v"].grad.double()
    g_rel = ((g_gpu - g_ref).norm() / g_ref.norm()).item()
    if verbose:
        print(f"smoke: loss gpu={ls['loss'].item():.6f} oracle={rl['loss'].item():.6f} colour_err={col_err:.2e} "
              f"sdf_err={sdf_err:.2e} grad_rel(lin4.weight_v)={g_rel:.2e}")
    assert col_err <= 2e-3 and sdf_err <= 1e-3 and g_rel <= 1e-2, (col_err, sdf_err, g_rel)
    assert abs(ls["loss"].item() - rl["loss"].item()) <= 2e-3 * max(1.0, abs(rl["loss"].item()))
    # flow / reprojection + unit-sphere losses on the same render, against the oracle
    if B < 2:
        return True
    from fmov_pose_b200 import flow
    sd = 2.0 / n_samples
    with torch.no_grad():
        pose_b = cpu(ts.pose_of(2))
        xy = torch.stack([cpu(px), cpu(py)], -1).float()
        Kc = cpu(torch.linalg.inv(ds.intrinsics_all_inv[1]))
        n = B // 2
        fl = flow.flow_loss({"z_vals": out["z_vals"].detach(), "weights": out["weights"].detach()}, ro.to(dev), rd.to(dev),
                            pose_b.to(dev), pose.to(dev), Kc.to(dev), Kc.to(dev), xy[:n].to(dev), xy[n:].to(dev), sd, 0.1)
        ul = flow.unit_sphere_loss({"z_vals": out["z_vals"].detach(), "weights": out["weights"].detach()}, ro.to(dev),
                                   rd.to(dev), sd, 0.05)
        w_ref, z_ref = ref["weights"].detach(), cpu(out["z_vals"])
        fl_ref = O.flow_loss(ro, rd, z_ref, w_ref, pose_b, pose, Kc, Kc, xy[:n], xy[n:], sd, 0.1)
        ul_ref = O.unit_sphere_loss(ro, rd, z_ref, w_ref, sd, 0.05)
    if verbose:
        print(f"smoke: flow loss gpu={fl.item():.5f} oracle={fl_ref.item():.5f}; unit-sphere gpu={ul.item():.6f} "
              f"oracle={ul_ref.item():.6f}")
    assert abs(fl.item() - fl_ref.item()) <= 5e-3 * abs(fl_ref.item()) + 1e-4, (fl.item(), fl_ref.item())
    if ul_ref.isfinite():        # nan when no sample lies outside the unit sphere (mean of an empty set, as the reference)
        assert abs(ul.item() - ul_ref.item()) <= 5e-3 * abs(ul_ref.item()) + 1e-5, (ul.item(), ul_ref.item())
    return True
