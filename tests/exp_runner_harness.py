"""Runs the reference's Runner (exp_runner.py, UNMODIFIED, from the oracle/_ref copy) on this package's models and checks
every train iteration against the oracle.  Executed as a subprocess by tests/test_gpu_zzzz_exp_runner.py with
tests/shims on PYTHONPATH (inert plotting stubs, pyhocon shim); test infrastructure only.

    python tests/exp_runner_harness.py <ref_root> <conf> <case> <out.json> [continue]
"""
import json
import os
import sys

import numpy as np
import torch

ref_root, conf, case, out_json = sys.argv[1:5]
is_continue = len(sys.argv) > 5 and sys.argv[5] == "continue"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, ref_root)
from fmov_pose_b200 import dropin  # noqa: E402

dropin.install(ref_root)
torch.set_default_tensor_type("torch.cuda.FloatTensor")          # exp_runner.py:2030
import exp_runner  # noqa: E402  (the reference's file, byte for byte)
from oracle import neus_oracle as O  # noqa: E402

torch.cuda.set_device(0)
scalars = {}


class RecWriter(exp_runner.SummaryWriter):
    def add_scalar(self, tag, value, step=None, *a, **k):
        scalars.setdefault(tag, []).append(float(value))
        return super().add_scalar(tag, value, step, *a, **k)


exp_runner.SummaryWriter = RecWriter
runner = exp_runner.Runner(conf, "train", case, "DTU", is_continue, -1, 0, False, has_global_conf=False)
res = {"classes": {k: type(getattr(runner, k)).__module__ for k in
                   ("renderer", "sdf_network", "color_network", "deviation_network", "pose_network", "dataset")},
       "iter_step_at_start": int(runner.iter_step)}
if is_continue:
    # checkpoint round trip (exp_runner.py:1109-1144 / 1414-1442): the continued Runner holds the saved state
    ck_dir = os.path.join(runner.base_exp_dir, "checkpoints")
    ck = torch.load(os.path.join(ck_dir, sorted(os.listdir(ck_dir))[-1]), map_location="cuda")
    worst = 0.0
    for key, net in (("sdf_network_fine", runner.sdf_network), ("color_network_fine", runner.color_network),
                     ("variance_network_fine", runner.deviation_network), ("pose_network", runner.pose_network)):
        sd = net.state_dict()
        assert set(sd.keys()) == set(ck[key].keys()), (key, set(sd.keys()) ^ set(ck[key].keys()))
        for k, v in sd.items():
            worst = max(worst, float((v.float() - ck[key][k].float()).abs().max()))
    res.update(ckpt_max_abs_diff=worst, ckpt_iter_step=int(ck["iter_step"]), ckpt_keys=sorted(ck.keys()),
               sdf_state_keys=sorted(runner.sdf_network.state_dict().keys()),
               current_pose_mlp_index=int(runner.current_pose_mlp_index))
else:
    parity = []
    orig_render = runner.renderer.render

    def render(rays_o, rays_d, near, far, **kw):
        B = rays_o.shape[0]
        t_rand = torch.rand([B, 1])                                    # models/renderer.py:404 (the draw render() makes)
        out = orig_render(rays_o, rays_d, near, far, t_rand=t_rand, **kw)
        rend = runner.renderer
        sdf_p = {k: v.detach() for k, v in runner.sdf_network.named_parameters()}
        col_p = {k: v.detach() for k, v in runner.color_network.named_parameters()}
        ref = O.render(sdf_p, col_p, runner.deviation_network.variance.detach(), rays_o.detach(), rays_d.detach(),
                       near.detach(), far.detach(), n_samples=rend.n_samples, n_importance=rend.n_importance,
                       up_sample_steps=rend.up_sample_steps, cos_anneal_ratio=kw.get("cos_anneal_ratio", 0.0),
                       background_rgb=kw.get("background_rgb"), z_vals=out["z_vals"].detach())
        parity.append(dict(rays=int(B),
                           colour=float((out["color_fine"] - ref["color_fine"]).abs().max()),
                           weight_sum=float((out["weight_sum"] - ref["weight_sum"]).abs().max()),
                           sdf=float((out["sdf"] - ref["sdf"]).abs().max()),
                           eikonal=abs(float(out["gradient_error"]) - float(ref["gradient_error"]))))
        return out

    runner.renderer.render = render
    p0 = [p.detach().clone() for p in runner.sdf_network.parameters()]
    pose0 = [p.detach().clone() for p in runner.pose_network.parameters()]
    ref_bug = None
    try:
        runner.train()
    except NameError as e:
        # the reference's own bug: the last statement of train() (exp_runner.py:976-980) calls extract_camera_poses, which
        # uses `csv` without importing it (exp_runner.py:57) — reached only after the last iteration
        if "csv" not in str(e) or runner.iter_step < runner.end_iter:
            raise
        ref_bug = str(e)
    res["reference_bug_after_last_iteration"] = ref_bug
    moved = max(float((a - b.detach()).abs().max()) for a, b in zip(p0, runner.sdf_network.parameters()))
    pose_moved = max(float((a.float() - b.detach().float()).abs().max()) for a, b in zip(pose0, runner.pose_network.parameters()))
    res.update(parity=parity, scalars=scalars, iter_step=int(runner.iter_step), sdf_param_moved=moved,
               pose_param_moved=pose_moved, current_image=int(runner.current_image),
               current_pose_mlp_index=int(runner.current_pose_mlp_index),
               checkpoints=sorted(os.listdir(os.path.join(runner.base_exp_dir, "checkpoints"))))
with open(out_json, "w") as fh:
    json.dump(res, fh)
print("HARNESS-OK")
