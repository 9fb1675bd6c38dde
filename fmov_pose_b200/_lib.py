"""ctypes binding of libfmov_b200.so (the C-ABI declared in include/fmov_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised (north_star: "no CPU fallback, no multi-backend dispatch")."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FMOV_LIB", os.path.join(_HERE, "libfmov_b200.so"))   # FMOV_LIB: A/B-testing builds
_lib = None

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(fmov_pose_b200 has no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.fmov_last_error.restype = ctypes.c_char_p
        _lib.fmov_sdf_fwd_blob_bytes.restype = c_ll
        _lib.fmov_sdf_fwd_blob_offset.restype = c_ll
        _lib.fmov_sdf_pair_blob_bytes.restype = c_ll
        _lib.fmov_sdf_pair_blob_offset.restype = c_ll
        for name in ("fmov_fine_blob_bytes", "fmov_grad_offset", "fmov_grad_floats", "fmov_mc_chunk_count", "fmov_mc_group_count"):
            getattr(_lib, name).restype = c_ll
        _lib.fmov_launch_count.restype = ctypes.c_ulonglong
    return _lib


n_calls = 0      # successful C-ABI calls (each launches >= 1 kernel)


def kernel_launches():
    """kernels of libfmov_b200.so launched by this process so far (counted inside the library at every launch site; a
    launch recorded into a CUDA graph counts once, when it is captured) — bench.py's `gpu_launches`"""
    return int(lib().fmov_launch_count())


def check(status, what):
    global n_calls
    n_calls += 1
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {lib().fmov_last_error().decode()}")


def ptr(t):
    """device pointer of a tensor (None -> NULL)"""
    if t is None:
        return c_void_p(0)
    assert t.is_cuda, "fmov_pose_b200 kernels take CUDA tensors only"
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def f32c(t):
    """contiguous fp32 view/copy"""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ---- optional per-kernel CUDA-event profiling (bench.py roofline) -------------------------------------
_profile = None


def profile_reset(enable):
    global _profile
    _profile = {} if enable else None


class timed:
    """with timed("fine_fwd"): <launch>  — records CUDA events on the current stream when profiling is on"""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if _profile is not None:
            self.e1.record()
            _profile.setdefault(self.name, []).append((self.e0, self.e1))
        return False


def profile_summary():
    if _profile is None:
        return {}
    torch.cuda.synchronize()
    return {k: {"ms": sum(a.elapsed_time(b) for a, b in v), "count": len(v)} for k, v in _profile.items()}
