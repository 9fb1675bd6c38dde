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

_SUBMODULES = ("fields", "renderer", "barf_fields", "camera", "picture_pose", "pixel_pose", "dataset", "embedder",
               "barf_embedder", "batch_lie_group_helper")
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
    return pkg


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
