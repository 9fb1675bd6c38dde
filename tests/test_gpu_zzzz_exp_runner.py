"""The reference's own driver, UNMODIFIED, on this package's kernels: `exp_runner.py` from the
git-ignored copy of the reference (oracle/_ref, oracle/build_ref.py — the reference checkout does not exist on the GPU box)
runs `Runner.__init__` + `train()` on a synthetic on-disk case with the shipped confs/ho3d_virtual.conf (schedule
shortened): HOCON kwargs -> constructors, the reference's Dataset loader subclassed onto the ray-generation kernel,
SegLearnPose's progressive schedule (disable_grad / enable_grad / finish_warmup), per-pose-MLP Adams, TensorBoard
scalars, checkpoint save and --is_continue load.  Third-party packages this image lacks are replaced by tests/shims
(pyhocon: a real parser; plotting / mesh-export libraries: inert)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "fmov_pose")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "exp_runner.py")),
                               reason="oracle/_ref is built by oracle/build_ref.py in the build container")


def _env():
    # TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD: the reference (written for torch 1.9) calls torch.load on its own checkpoint, which
    # holds numpy scalars; torch >= 2.6 refuses that by default
    return dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "shims"), ROOT]), TQDM_DISABLE="1",
                TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD="1")


@needs_ref
@pytest.mark.parametrize("conf_name,rays,last_mlp", [
    ("ho3d_virtual.conf", 256, 3),      # C2: SegLearnPose, 32+0, maintain_shape: 128 + 128 rays (exp_runner.py:512-548)
    ("ho3d_barf.conf", 128, 0),         # C4: BARF networks + LearnPoseGF, 64+64, 4 up-sample rounds
])
def test_runner_train_iterations_through_the_harness_and_checkpoint_round_trip(tmp_path, conf_name, rays, last_mlp):
    from tests import _synth_case
    work = str(tmp_path)
    _synth_case.write_case(work)
    conf = _synth_case.write_conf(REF, work, end_iter=6, name=conf_name)
    out = os.path.join(work, "harness.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "exp_runner_harness.py"), REF, conf, "SYN_ori", out],
                       cwd=work, env=_env(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "HARNESS-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-5000:]
    res = json.load(open(out))
    for k in ("renderer", "sdf_network", "color_network", "deviation_network", "pose_network"):
        assert res["classes"][k].startswith("fmov_pose_b200.models"), res["classes"]
    assert res["iter_step"] == 6 and len(res["parity"]) == 6
    # every iteration of the reference's loop: this path's render vs the fp32 oracle on the live weights and poses
    for it, p in enumerate(res["parity"]):
        assert p["rays"] == rays
        assert p["colour"] <= 2e-3 and p["sdf"] <= 1e-3 and p["weight_sum"] <= 5e-3 and p["eikonal"] <= 1e-3, (it, p)
    sc = res["scalars"]
    assert len(sc["Loss/loss"]) == 6 and all(v == v and abs(v) < 1e3 for v in sc["Loss/loss"]), sc["Loss/loss"]
    for tag in ("Loss/color_loss", "Loss/eikonal_loss", "Loss/mask_loss", "Statistics/s_val", "Statistics/psnr"):
        assert len(sc[tag]) == 6, tag
    assert res["sdf_param_moved"] > 0 and res["pose_param_moved"] > 0          # both optimisers stepped
    # ho3d_virtual: the progressive schedule advanced through all frames (2 iterations per frame); ho3d_barf: one pose MLP
    assert res["current_image"] == 4 and res["current_pose_mlp_index"] == last_mlp
    assert res["checkpoints"], "save_checkpoint (exp_runner.py:1414-1442) wrote nothing"
    # --is_continue: a second Runner loads the checkpoint the first one wrote
    out2 = os.path.join(work, "harness2.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "exp_runner_harness.py"), REF, conf, "SYN_ori", out2,
                        "continue"], cwd=work, env=_env(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "HARNESS-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-5000:]
    res2 = json.load(open(out2))
    assert res2["iter_step_at_start"] == 6 and res2["ckpt_iter_step"] == 6 and res2["ckpt_max_abs_diff"] == 0.0
    # state_dict keys are the reference's (old-style weight norm, Barf buffers): SURVEY.md §5 checkpoint contract
    for k in ("lin0.weight_g", "lin0.weight_v", "lin0.bias", "lin8.weight_v", "noise_poses", "se3_refine.weight", "progress"):
        assert k in res2["sdf_state_keys"], k
    want = {"nerf", "sdf_network_fine", "variance_network_fine", "color_network_fine", "optimizer", "iter_step", "pose_network"}
    if last_mlp:
        want |= {"current_pose_mlp_index", "pro_iteration"}
    assert want <= set(res2["ckpt_keys"])


@needs_ref
def test_dropin_command_line_trains(tmp_path):
    """`python -m fmov_pose_b200.dropin exp_runner.py --mode train ...`: the command a user of the reference types.  The
    train phase must complete (iteration reports, TensorBoard events, checkpoint); what `__main__` runs
    AFTER training — render_poses / validate_mesh visualisation through matplotlib, open3d, imageio, trimesh
    (exp_runner.py:2126-2128) — needs those packages for real, so a failure there is tolerated only if it is inside
    that post-training visualisation."""
    from tests import _synth_case
    work = str(tmp_path)
    _synth_case.write_case(work)
    conf = _synth_case.write_conf(REF, work, end_iter=5)
    r = subprocess.run([sys.executable, "-m", "fmov_pose_b200.dropin", os.path.join(REF, "exp_runner.py"), "--mode", "train",
                        "--conf", conf, "--case", "SYN_ori"], cwd=work, env=_env(), capture_output=True, text=True,
                       timeout=900)
    log = r.stdout + r.stderr
    exp = os.path.join(work, "exp", "SYN_ori", "ours_wo_global_conf")
    assert "Hello FMOV" in r.stdout and "mode:  train" in r.stdout, log[-4000:]
    assert log.count("iter:") >= 5, log[-4000:]
    assert os.listdir(os.path.join(exp, "checkpoints")), log[-3000:]
    assert any(f.startswith("events.out.tfevents") for f in os.listdir(os.path.join(exp, "logs")))
    if r.returncode != 0:
        # known ways the REFERENCE's own tail fails here, all after the last training iteration: (a) its bug at
        # exp_runner.py:57 (`csv` is used but never imported) in the last statement of train(); (b) the plotting / export
        # calls of render_poses / validate_mesh against the inert stand-ins
        tail = r.stderr[-6000:]
        assert "name 'csv' is not defined" in tail or "render_poses" in tail or "validate_mesh" in tail, tail
