"""TEST SHIM — loaded automatically by Python when tests/shims is on PYTHONPATH.  Replaces the third-party packages that
the reference's exp_runner.py / utils import but this image does not have (plotting, mesh export, config and debugging
libraries: all outside the train-step path) with inert stand-ins, so that the reference's driver can be executed
unmodified by tests/test_gpu_zzzz_exp_runner.py.  `pyhocon` is NOT inert: tests/shims/pyhocon parses the confs for real.
Packages that ARE installed are never shadowed."""
import importlib.abc
import importlib.machinery
import importlib.util
import sys
import types

_CANDIDATES = ("trimesh", "imageio", "open3d", "plotly", "dash", "plyfile", "easydict", "mcubes", "xatlas", "matplotlib",
               "lpips", "kornia", "pytorch3d", "skimage", "icecream", "seaborn")


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, n):
        if n.startswith("__") and n.endswith("__"):
            raise AttributeError(n)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Mod(types.ModuleType):
    __path__ = []

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        if self.__name__ == "easydict" and n == "EasyDict":
            class EasyDict(dict):
                __getattr__ = dict.__getitem__
                __setattr__ = dict.__setitem__
            return EasyDict
        return _Anything()


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, names):
        self.names = names

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.names:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Mod(spec.name)

    def exec_module(self, module):
        pass


def _install():
    missing = []
    for m in _CANDIDATES:
        try:
            if importlib.util.find_spec(m) is None:
                missing.append(m)
        except (ImportError, ValueError):
            missing.append(m)
    if missing:
        sys.meta_path.append(_Finder(tuple(missing)))
    return missing


INERT = _install()
