"""Pin the CPU oracle (oracle/neus_oracle.py) against fixtures produced by the imported
reference (oracle/gen_golden.py).  CPU-only; part of `-m "not gpu"`."""
import numpy as np
import pytest
import torch

from oracle import neus_oracle as O
from tests._util import load_golden, params_from, t

RENDER_CASES = ["small_6464_seg", "small_3200_seg", "small_6464_se3_white", "small_1632_gf_nomask",
                "full_6464_gf", "full_3200_seg"]


def test_kat_sample_pdf():
    d = load_golden("kat")
    for n in (16, 5):
        out = O.sample_pdf_det(t(d, "pdf.bins"), t(d, "pdf.w"), n)
        np.testing.assert_array_equal(out.numpy(), d[f"pdf.out{n}"])


def test_kat_pose_math():
    d = load_golden("kat")
    np.testing.assert_array_equal(O.rodrigues_exp(t(d, "exp.r")).numpy(), d["exp.R"])
    np.testing.assert_allclose(O.se3_to_SE3(t(d, "se3.wu")).numpy(), d["se3.Rt"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(O.compose_pair(t(d, "compose.a"), t(d, "compose.b")).numpy(),
                               d["compose.out"], rtol=0, atol=1e-7)


def test_kat_softplus_and_pe():
    d = load_golden("kat")
    np.testing.assert_array_equal(O.softplus100(t(d, "sp.z")).numpy(), d["sp.y"])
    x = t(d, "pe.x")
    np.testing.assert_array_equal(O.embed(x, 6).numpy(), d["pe.e6"])
    np.testing.assert_array_equal(O.embed(x, 4).numpy(), d["pe.e4"])
    # BARF embedder never applies its coarse-to-fine weights (barf_embedder.py:50-56)
    np.testing.assert_array_equal(O.embed(x, 6).numpy(), d["pe.barf6_p03"])


def test_kat_up_sample():
    d = load_golden("kat")
    for inv_s in (64, 512):
        out = O.up_sample(t(d, "up.o"), t(d, "up.d"), t(d, "up.z"), t(d, "up.sdf"), 16, inv_s)
        np.testing.assert_allclose(out.numpy(), d[f"up.new{inv_s}"], rtol=0, atol=1e-6)


def _pose_from_fixture(d):
    kind = str(d["pose_kind"])
    img_id = int(d["img_id"])
    init = t(d, "init_c2w")
    leaves = []
    if kind in ("seg", "gf"):
        p = {k[len("pose."):]: t(d, k) for k in d if k.startswith("pose.")}
        for k in p:
            if k.startswith("lin") and not k.startswith("lin3_trans"):
                p[k].requires_grad_(True)
                leaves.append((k, p[k]))
        rot, trans, scale = O.pose_gf_mlp(p, img_id, emphasize_rot=(kind == "seg"))
        pose = O.pose_gf_compose(rot, trans, init[img_id, :3, :], scale)
    else:
        se3 = t(d, "sdf.se3_refine.weight").requires_grad_(True)
        leaves.append(("se3", se3))
        pose = O.barf_poses(se3, t(d, "sdf.noise_poses"))[img_id]
    return pose, leaves


@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_step_matches_reference(name):
    d = load_golden(name)
    B, n, m, steps, dh = [int(v) for v in d["cfg"]]
    sdf_p = params_from(d, "sdf.", requires_grad=True)
    col_p = params_from(d, "col.", requires_grad=True)
    var = t(d, "variance").requires_grad_(True)
    pose, pose_leaves = _pose_from_fixture(d)
    np.testing.assert_allclose(pose.detach().numpy(), d["pose"], rtol=0, atol=2e-6)
    # render from the fixture's pose (leaf) so that 1e-7 pose differences are not amplified by
    # the inverse-CDF sampling; the pose chain itself is checked above and its backward below.
    pose_chain = pose
    pose = t(d, "pose").requires_grad_(True)
    bg = torch.ones(1, 3) if int(d["white"]) else None
    losses, out = O.train_step(sdf_p, col_p, var, pose, t(d, "intr_inv"),
                               torch.from_numpy(d["px"]), torch.from_numpy(d["py"]),
                               t(d, "true_rgb"), t(d, "mask"), t_rand=t(d, "t_rand"),
                               n_samples=n, n_importance=m, up_sample_steps=steps,
                               cos_anneal_ratio=float(d["cos_anneal"]), igr_weight=0.1,
                               mask_weight=float(d["mask_weight"]), background_rgb=bg)
    tol = 5e-4   # fp32 op-order noise of the reference itself, amplified by Softplus(beta=100)
    for k in ["color_fine", "depth_fine", "s_val", "cdf_fine", "weight_sum", "weight_max",
              "gradients", "weights", "gradient_error", "inside_sphere", "pts"]:
        np.testing.assert_allclose(out[k].detach().numpy(), d["out." + k], rtol=1e-3, atol=tol,
                                   err_msg=k)
    got = [losses[k].item() for k in ("loss", "color_loss", "eikonal_loss", "mask_loss")]
    np.testing.assert_allclose(got, d["loss"], rtol=1e-5, atol=1e-6)
    losses["loss"].backward()
    g = pose.grad.numpy()
    np.testing.assert_allclose(g, d["grad.pose"], rtol=2e-3, atol=2e-5 * np.abs(d["grad.pose"]).max())
    for k, v in d.items():
        if not k.startswith("gnorm."):
            continue
        nm = k[len("gnorm."):]
        if nm == "variance":
            gr = var.grad
        else:
            net, key = nm.split(".", 1)
            if not key.startswith("lin"):
                continue   # se3_refine etc. are covered by the pose-parameter check below
            gr =(sdf_p if net == "sdf" else col_p)[key].grad
        if gr is None:
            assert float(v) == 0.0, nm
            continue
        np.testing.assert_allclose(np.linalg.norm(gr.double().numpy()), float(v), rtol=2e-3, err_msg=nm)
        if "grad." + nm in d:
            ref = d["grad." + nm]
            np.testing.assert_allclose(gr.numpy(), ref, rtol=0, atol=3e-3 * np.abs(ref).max() + 1e-9,
                                       err_msg=nm)
    # pose-parameter gradients: push the reference's d(loss)/d(pose) through the oracle pose chain
    pose_chain.backward(t(d, "grad.pose"))
    for i, (k, leaf) in enumerate(pose_leaves):
        ref = d[f"grad.pose_param{i}"]
        assert leaf.grad is not None, k
        np.testing.assert_allclose(leaf.grad.numpy(), ref, rtol=0, atol=3e-3 * np.abs(ref).max() + 1e-12,
                                   err_msg=k)


def test_grid_query_matches_reference():
    d = load_golden("grid40")
    sdf_p = params_from(d, "sdf.")
    res = int(d["res"])
    u = O.extract_fields(sdf_p, [-1.01] * 3, [1.01] * 3, res)
    np.testing.assert_allclose(u.numpy(), d["u"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", ["flow_half", "flow_quarter_detach"])
def test_flow_and_unit_sphere_losses_match_the_reference_lines(name):
    """fixtures = the reference's own exp_runner.py lines 605-688 / 714-724 exec'd by oracle/gen_golden.py"""
    d = load_golden(name)
    lv = {k: t(d, k).requires_grad_(True) for k in ("rays_o", "rays_d", "z", "weights", "c2w_0", "c2w_1")}
    K = t(d, "intrinsics")
    sd = float(d["sample_dist"])
    np.testing.assert_allclose(O.sample_points(lv["rays_o"], lv["rays_d"], lv["z"], sd).reshape(-1, 3).detach().numpy(),
                               d["pts"], atol=1e-6)
    fl = O.flow_loss(lv["rays_o"], lv["rays_d"], lv["z"], lv["weights"], lv["c2w_1"], lv["c2w_0"], K[1], K[0],
                     t(d, "pixels_xy"), t(d, "pixels_xy_corr"), sd, float(d["flow_weight"]),
                     maintain_shape=bool(d["maintain_shape"]), detach_flow_on_sdf=bool(d["detach_flow_on_sdf"]))
    np.testing.assert_allclose(fl.item(), float(d["flow_loss"]), rtol=2e-5)
    g = torch.autograd.grad(fl, list(lv.values()), allow_unused=True)
    for (k, v), gi in zip(lv.items(), g):
        ref = d["gflow_" + k]
        got = np.zeros_like(ref) if gi is None else gi.numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-4 * np.abs(ref).max() + 1e-9, err_msg=k)
    ul = O.unit_sphere_loss(lv["rays_o"], lv["rays_d"], lv["z"], lv["weights"], sd, float(d["unit_sphere_weight"]))
    np.testing.assert_allclose(ul.item(), float(d["unit_sphere_loss"]), rtol=1e-5)
    gu, = torch.autograd.grad(ul, [lv["weights"]])
    np.testing.assert_allclose(gu.numpy(), d["gunit_weights"], rtol=1e-5, atol=1e-9)
