"""Run the reference's `exp_runner.py` UNCHANGED on the B200 kernels.

The reference has no plugin interface: `exp_runner.py:13-25` simply does `from models.<module> import <class>` and calls
the classes.  `install()` makes those imports resolve to `fmov_pose_b200.models` by aliasing the package
and its modules in `sys.modules`; everything else the runner imports (`utils.*`, `confs/`, its own working directory)
stays the reference's.

    cd <reference checkout>
    python -m fmov_pose_b200.dropin exp_runner.py --mode train --conf confs/ho3d_virtual.conf --case AP13_ori ...

Image / camera loading (`Dataset.__init__`, models/dataset.py:146-545: disk I/O and one-off preprocessing) is not rebuilt:
`models.dataset.Dataset` resolves to a subclass of the reference's own loader, loaded from the checkout by file path,
whose ray functions run on the fused pose / ray-generation kernel (see models/dataset.py::make_dataset_class)."""
import importlib
import importlib.util
import os
import runpy
import sys
import types

_SUBMODULES = ("fields", "renderer", "barf_fields", "camera", "picture_pose", "dataset", "embedder", "batch_lie_group_helper")
reference_root = os.environ.get("FMOV_REFERENCE_ROOT")          # where the reference's own models/dataset.py lives


def install(root=None):
    """alias `models` and `models.<module>` to the B200 mirror; `root` = the reference checkout (for the data loader)"""
    global reference_root
    if root is not None:
        reference_root = os.path.abspath(root)
    pkg = importlib.import_module("fmov_pose_b200.models")
    sys.modules["models"] = pkg
    for name in _SUBMODULES:
        sys.modules["models." + name] = importlib.import_module("fmov_pose_b200.models." + name)
    # two module names of the reference have no file of their own in the mirror:
    #  * models.barf_embedder (models/barf_embedder.py:6-75): `get_embedder` there is the mirror's get_barf_embedder — the
    #    BARF embedder computes its coarse-to-fine weights and never applies them
    #  * models.pixel_pose (models/pixel_pose.py:28-388): exp_runner.py:25 imports SegDeepPixelPose unconditionally, but the
    #    per-pixel pose MLPs are only built with model.pixel_level = True (no shipped conf; out of scope, SURVEY.md §2 row
    #    18): the names resolve to a class that fails loudly on construction
    emb = sys.modules["models.embedder"]
    be = types.ModuleType("models.barf_embedder")
    be.Embedder, be.get_embedder = emb.Embedder, emb.get_barf_embedder
    sys.modules["models.barf_embedder"] = be
    pp = types.ModuleType("models.pixel_pose")
    for cls_name in ("PixelPose", "DeepPixelPose", "SegDeepPixelPose"):
        setattr(pp, cls_name, type(cls_name, (_UnsupportedPixelPose,), {"__module__": "fmov_pose_b200.models.pixel_pose"}))
    sys.modules["models.pixel_pose"] = pp
    pkg.barf_embedder, pkg.pixel_pose = be, pp
    return pkg


class _UnsupportedPixelPose:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__} (model.pixel_level = True) is not part of the B200 train-step path: every shipped conf "
            "uses the per-frame pose modules of models/picture_pose.py (LearnPoseGF / SegLearnPose)")


def load_reference_module(relpath, alias):
    """import one file of the reference checkout under a private module name (its `models` package is shadowed)"""
    if alias in sys.modules:
        return sys.modules[alias]
    if not reference_root:
        raise RuntimeError("the reference checkout is unknown: call fmov_pose_b200.dropin.install(root) or set "
                           "FMOV_REFERENCE_ROOT (needed for the reference's own data loader)")
    path = os.path.join(reference_root, relpath)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found")
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules[alias]
        raise
    return mod


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print("usage: python -m fmov_pose_b200.dropin <path/to/exp_runner.py> [its arguments]", file=sys.stderr)
        return 2
    script = os.path.abspath(argv[0])
    root = os.path.dirname(script)
    # `python -m fmov_pose_b200.dropin` runs this file as __main__; the models package imports the CANONICAL module
    # (fmov_pose_b200.dropin), so the aliases and the checkout location are installed there
    import fmov_pose_b200.dropin as canonical
    canonical.install(root)
    sys.path.insert(0, root)                 # `utils.*` and friends resolve as if the script had been started directly
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
