"""CPU checks of the test infrastructure around the reference's driver: the pyhocon stand-in (tests/shims/pyhocon) against
the ConfigTree calls exp_runner.py / models/dataset.py make, the synthetic on-disk case, and one CPU step of the
reference arm (oracle/ref_arm.py, the reference's own sources from the git-ignored oracle/_ref)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "fmov_pose")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "exp_runner.py")),
                               reason="oracle/_ref is built by oracle/build_ref.py in the build container")

CONF = """
general {
    base_exp_dir = ./exp/CASE_NAME/ours     # comment
    recording = [
        ./,
        ./models
    ]
}
train { learning_rate = 5e-4, end_iter = 300000
        use_white_bkgd = False
        "quoted_key" = 3 }
model {
    pose_type = seg
    nerf { D = 8, skips=[4], use_viewdirs=True }
    sdf_network { d_out = 257
                  scale = 1.0 }
}
"""


def _pyhocon():
    sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
    try:
        import importlib
        import pyhocon
        importlib.reload(pyhocon)
        return pyhocon
    finally:
        sys.path.pop(0)


def test_pyhocon_shim_semantics():
    ph = _pyhocon()
    c = ph.ConfigFactory.parse_string(CONF)
    assert c["general.base_exp_dir"] == "./exp/CASE_NAME/ours" and c["general.recording"] == ["./", "./models"]
    assert c.get_float("train.learning_rate") == 5e-4 and c.get_int("train.end_iter") == 300000
    assert c.get_bool("train.use_white_bkgd") is False and c["train.quoted_key"] == 3
    assert dict(**c["model.nerf"]) == {"D": 8, "skips": [4], "use_viewdirs": True}
    assert c["model"]["sdf_network"].get_int("d_out") == 257 and c.get_float("model.sdf_network.scale") == 1.0
    assert "model.barf" not in c and "model.nerf" in c
    assert c.get("train.missing", False) is False and c.get_int("train.missing", default=7) == 7
    with pytest.raises(ph.ConfigMissingException):
        c.get_int("train.missing")
    c.put("train.flow_interval", 3)
    c.put("model.barf", False)
    assert c["train.flow_interval"] == 3 and c["model.barf"] is False
    sub = c["model"]
    assert sub.get("pose_type", default="None") == "seg"


@needs_ref
def test_pyhocon_shim_parses_every_shipped_conf():
    ph = _pyhocon()
    n = 0
    for f in sorted(os.listdir(os.path.join(REF, "confs"))):
        c = ph.ConfigFactory.parse_string(open(os.path.join(REF, "confs", f)).read().replace("CASE_NAME", "X"))
        assert c.get_int("train.batch_size") == 512 and c["model.sdf_network.skip_in"] == [4]
        assert c["model.neus_renderer.n_outside"] == 0 and c.get_float("model.variance_network.init_val") == 0.3
        n += 1
    assert n >= 4


@needs_ref
def test_synthetic_case_and_conf_edit(tmp_path):
    from tests import _synth_case
    data = _synth_case.write_case(str(tmp_path))
    cams = np.load(os.path.join(data, "cameras_sphere.npz"))
    assert len(os.listdir(os.path.join(data, "image"))) == _synth_case.N_IMAGES == len(os.listdir(os.path.join(data, "mask_obj")))
    assert cams["world_mat_0000"].shape == (4, 4) and np.allclose(cams["scale_mat_0003"], np.eye(4))
    assert os.path.isdir(os.path.join(str(tmp_path), "data", "HO3Dv3", "matches", "SYN"))
    ph = _pyhocon()
    for name, n_s in (("ho3d_virtual.conf", 32), ("ho3d_barf.conf", 64)):
        conf = ph.ConfigFactory.parse_string(open(_synth_case.write_conf(REF, str(tmp_path), end_iter=6, name=name)).read())
        assert conf.get_int("train.end_iter") == 6 and conf.get_int("train.save_freq") == 6
        assert conf["general.recording"] == ["./"] and conf["model.neus_renderer.n_samples"] == n_s


@needs_ref
def test_reference_arm_one_cpu_step():
    """the baseline arm really runs the reference's sources: one tiny train iteration on the CPU, loss finite"""
    code = ("import sys; sys.path.insert(0, %r); from oracle import ref_arm; import torch; torch.set_num_threads(4); "
            "step = ref_arm.build_step(32, 16, 16, 2, device='cpu', n_images=3); a = float(step()); b = float(step()); "
            "import models.renderer as r; assert r.__file__.startswith(%r), r.__file__; "
            "assert a == a and b == b and 0 < a < 100; print('REF-ARM-OK', a, b)") % (ROOT, REF)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "REF-ARM-OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
