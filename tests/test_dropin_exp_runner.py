"""Drop-in boundary: the reference's exp_runner.py, UNMODIFIED, must import against the B200 mirror of its
`models` package.  Runs in this container only (needs /root/reference); third-party packages that are not installed here
(pyhocon, trimesh, open3d, plotly, ...) are replaced by inert stand-ins — they are plotting / config / export libraries
outside the train-step path."""
import inspect
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "exp_runner.py")),
                                reason="the reference checkout only exists in the build container")

_STUB_FINDER = textwrap.dedent('''
    import importlib.abc, importlib.machinery, sys, types
    MISSING = ("pyhocon", "trimesh", "imageio", "open3d", "plotly", "dash", "plyfile", "easydict", "mcubes", "xatlas",
               "matplotlib", "sklearn", "lpips", "kornia", "pytorch3d", "skimage", "icecream", "seaborn")
    class _Anything:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return _Anything()
        def __getattr__(self, n): return _Anything()
        def __iter__(self): return iter(())
        def __mro_entries__(self, bases): return (object,)
    class _Mod(types.ModuleType):
        __path__ = []
        def __getattr__(self, n):
            if n.startswith("__"): raise AttributeError(n)
            return _Anything()
    class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, name, path=None, target=None):
            top = name.split(".")[0]
            if top in MISSING:
                try:
                    if top not in sys.modules or isinstance(sys.modules[top], _Mod):
                        return importlib.machinery.ModuleSpec(name, self, is_package=True)
                except Exception:
                    pass
            return None
        def create_module(self, spec): return _Mod(spec.name)
        def exec_module(self, module): pass
    def install_stubs():
        import importlib.util
        really_missing = tuple(m for m in MISSING if importlib.util.find_spec(m) is None)
        globals()["MISSING"] = really_missing
        sys.meta_path.append(_Finder())
''')


def _run(code):
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", _STUB_FINDER + textwrap.dedent(code)], cwd=REF, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def test_exp_runner_imports_unmodified_against_the_mirror():
    out = _run('''
        install_stubs()
        import sys
        sys.path.insert(0, "/root/reference")
        from fmov_pose_b200 import dropin
        dropin.install("/root/reference")
        import exp_runner                                  # the reference's file, byte for byte
        for name in ("NeuSRenderer", "SDFNetwork", "RenderingNetwork", "SingleVarianceNetwork", "NeRF", "BarfSDFNetwork",
                     "BarfRenderingNetwork", "LearnPoseGF", "SegLearnPose", "SegDeepPixelPose", "Dataset"):
            cls = getattr(exp_runner, name)
            assert cls.__module__.startswith("fmov_pose_b200.models"), (name, cls.__module__)
        assert exp_runner.camera.__name__ == "fmov_pose_b200.models.camera"
        assert hasattr(exp_runner.camera, "lie") and hasattr(exp_runner.camera, "pose")
        import torch
        x = torch.arange(6.0).reshape(2, 3)
        assert exp_runner.to_hom(x).shape == (2, 4) and float(exp_runner.to_hom(x)[1, 3]) == 1.0
        # the loader is the reference's own class with the ray functions replaced
        ref_ds = sys.modules["_fmov_reference_models_dataset"].Dataset
        assert issubclass(exp_runner.Dataset, ref_ds) and exp_runner.Dataset is not ref_ds
        for fn in ("gen_random_rays_at", "gen_rays_at", "gen_random_ray_pairs_at", "near_far_from_sphere"):
            assert getattr(exp_runner.Dataset, fn) is not getattr(ref_ds, fn), fn
        assert exp_runner.Dataset.__init__ is ref_ds.__init__
        # utils/align_poses.py:4 pulls a helper out of models.dataset
        from models.dataset import load_K_Rt_from_P
        assert load_K_Rt_from_P.__module__ == "_fmov_reference_models_dataset"
        assert exp_runner.Runner.__init__.__code__.co_filename == "/root/reference/exp_runner.py"
        print("IMPORT-OK")
    ''')
    assert "IMPORT-OK" in out


def test_replaced_ray_functions_keep_the_reference_signatures():
    out = _run('''
        install_stubs()
        import inspect, sys
        sys.path.insert(0, "/root/reference")
        from fmov_pose_b200 import dropin
        dropin.install("/root/reference")
        from models.dataset import Dataset
        ref_ds = sys.modules["_fmov_reference_models_dataset"].Dataset
        for fn in ("gen_random_rays_at", "gen_rays_at", "gen_random_ray_pairs_at", "near_far_from_sphere"):
            a, b = inspect.signature(getattr(Dataset, fn)), inspect.signature(getattr(ref_ds, fn))
            assert list(a.parameters) == list(b.parameters), (fn, a, b)
            assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()], fn
        print("SIG-OK")
    ''')
    assert "SIG-OK" in out


def test_constructor_and_method_signatures_match_the_reference_classes():
    """ctor kwargs are splatted from the HOCON model{} blocks (exp_runner.py:178-216, 271-277): names must match"""
    out = _run('''
        install_stubs()
        import importlib.util, inspect, sys
        sys.path.insert(0, "/root/reference")
        def ref(path, name):
            spec = importlib.util.spec_from_file_location(name, "/root/reference/models/" + path)
            m = importlib.util.module_from_spec(spec); sys.modules[name] = m; spec.loader.exec_module(m); return m
        import models.embedder                          # reference's own helpers for its fields.py
        r_fields, r_rend, r_pose = ref("fields.py", "_r_fields"), ref("renderer.py", "_r_rend"), ref("picture_pose.py", "_r_pose")
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        from fmov_pose_b200.models import fields, renderer, picture_pose
        def params(f):
            return [p for p in inspect.signature(f).parameters if p != "self"]
        for mine, theirs in ((fields.SDFNetwork, r_fields.SDFNetwork), (fields.RenderingNetwork, r_fields.RenderingNetwork),
                             (fields.SingleVarianceNetwork, r_fields.SingleVarianceNetwork),
                             (renderer.NeuSRenderer, r_rend.NeuSRenderer), (picture_pose.LearnPoseGF, r_pose.LearnPoseGF),
                             (picture_pose.SegLearnPose, r_pose.SegLearnPose)):
            want = params(theirs.__init__)
            got = params(mine.__init__)
            assert got[:len(want)] == want, (mine.__name__, got, want)
        want = params(r_rend.NeuSRenderer.render)
        got = params(renderer.NeuSRenderer.render)
        assert got[:len(want)] == want, (got, want)
        for m in ("extract_geometry", "extract_color"):
            assert params(getattr(renderer.NeuSRenderer, m))[:3] == params(getattr(r_rend.NeuSRenderer, m))[:3], m
        for m in ("sdf", "gradient", "sdf_hidden_appearance", "forward"):
            assert hasattr(fields.SDFNetwork, m), m
        print("CTOR-OK")
    ''')
    assert "CTOR-OK" in out


def test_launcher_runs_a_script_with_the_aliases_in_place(tmp_path):
    """python -m fmov_pose_b200.dropin <script> [args]: the script sees `models.*` = the mirror, its own directory on
    sys.path and its own argv (no reference checkout needed for this one)"""
    script = tmp_path / "runner_like.py"
    (tmp_path / "utils").mkdir()
    (tmp_path / "utils" / "__init__.py").write_text("MARK = 'local utils'\n")
    script.write_text(textwrap.dedent('''
        import sys
        from models.fields import SDFNetwork, NeRF
        from models.renderer import NeuSRenderer
        import models.camera as camera
        from models.camera import to_hom
        from models.picture_pose import LearnPoseGF, SegLearnPose
        from models.pixel_pose import SegDeepPixelPose
        from utils import MARK
        assert __name__ == "__main__" and MARK == "local utils"
        assert SDFNetwork.__module__ == "fmov_pose_b200.models.fields"
        try:
            SegDeepPixelPose(3)
        except NotImplementedError:
            print("ARGS", sys.argv[1:])
    '''))
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "fmov_pose_b200.dropin", str(script), "--mode", "train"], env=env,
                       capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    assert "ARGS ['--mode', 'train']" in r.stdout
