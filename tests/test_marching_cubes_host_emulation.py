"""The per-point device code of the marching-cubes kernels (csrc/mc_core.cuh: case build-up, vertex interpolation,
edge -> vertex-id mapping) compiled by g++ and driven like the kernels drive it, against the numpy oracle.  This is the
CPU-side check of the CUDA path's logic; the kernels themselves are checked on the GPU by test_gpu_zz_marching_cubes.py."""
import os
import subprocess

import numpy as np
import pytest

from fmov_pose_b200 import mc_tables as T
from oracle import marching_cubes as MC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    d = tmp_path_factory.mktemp("mc_emul")
    exe = str(d / "mc_host_emul")
    subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", os.path.join(ROOT, "tests", "host", "mc_host_emul.cpp"),
                    "-o", exe], check=True)
    tab = str(d / "tables.bin")
    with open(tab, "wb") as f:
        f.write(np.ascontiguousarray(T.TRI_TABLE, dtype=np.int8).tobytes())
        f.write(np.ascontiguousarray(T.N_TRIS, dtype=np.uint8).tobytes())

    def run(u, iso):
        up, vp, tp = str(d / "u.bin"), str(d / "v.bin"), str(d / "t.bin")
        np.ascontiguousarray(u, dtype=np.float32).tofile(up)
        X, Y, Z = u.shape
        subprocess.run([exe, str(X), str(Y), str(Z), repr(float(iso)), up, tab, vp, tp], check=True)
        return np.fromfile(vp, dtype=np.float32).reshape(-1, 3), np.fromfile(tp, dtype=np.int32).reshape(-1, 3)
    return run


def _field(kind, shape, seed=0):
    ax = [np.linspace(-1.0, 1.0, n) for n in shape]
    xx, yy, zz = np.meshgrid(*ax, indexing="ij")
    if kind == "sphere":
        return (0.55 - np.sqrt(xx ** 2 + yy ** 2 + zz ** 2)).astype(np.float32)
    if kind == "waves":
        return (np.sin(5 * xx) * np.cos(4 * yy) + np.sin(3 * zz + 0.3) * 0.7 + 0.1).astype(np.float32)
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)


@pytest.mark.parametrize("kind,shape,iso", [("sphere", (33, 29, 31), 0.0), ("waves", (37, 18, 50), 0.05),
                                            ("noise", (17, 16, 19), 0.0), ("noise", (2, 2, 2), 0.0),
                                            ("noise", (3, 129, 2), 0.25), ("sphere", (9, 300, 7), 0.0),
                                            # Z % 4 == 0: the count pass works on quads (mc_count_quad_kernel's code)
                                            ("sphere", (30, 31, 32), 0.0), ("waves", (21, 18, 52), 0.05),
                                            ("noise", (9, 7, 4), 0.0), ("noise", (2, 2, 8), 0.1),
                                            ("noise", (5, 67, 12), -0.2),
                                            # Y * Z % 256 == 0: the count pass marches along x (mc_count_march_kernel's code)
                                            ("noise", (5, 16, 16), 0.0), ("noise", (2, 8, 64), 0.1), ("waves", (19, 40, 32), 0.05),
                                            ("sphere", (33, 24, 32), 0.0), ("noise", (3, 2, 128), 0.0),
                                            ("noise", (4, 2, 512), 0.3)])
def test_device_logic_equals_the_oracle(emul, kind, shape, iso):
    u = _field(kind, shape)
    v, t = emul(u, iso)
    v_ref, t_ref = MC.marching_cubes(u, iso, T.TRI_TABLE, T.N_TRIS)
    assert v.shape == v_ref.shape and t.shape == t_ref.shape
    np.testing.assert_allclose(v, v_ref, rtol=0, atol=2e-5)
    np.testing.assert_array_equal(t.astype(np.int64), t_ref)          # also proves no unwritten vid3 entry was read


def test_python_glue_end_to_end_on_a_host_stand_in(tmp_path, monkeypatch):
    """fmov_pose_b200/mcubes_gpu.py (marshalling, chunk prefix sums, output sizing, world-coordinate transform) against the
    oracle, with the five C entry points served by a host build of the same per-point code and CPU tensors standing in for
    device memory (the product refuses CPU tensors; the test lifts exactly that check)."""
    import contextlib
    import ctypes

    import torch

    from fmov_pose_b200 import _lib as L
    from fmov_pose_b200 import mcubes_gpu

    so = str(tmp_path / "libmc_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                    os.path.join(ROOT, "tests", "host", "mc_host_lib.cpp"), "-o", so], check=True)
    fake = ctypes.CDLL(so)
    fake.fmov_mc_chunk_count.restype = ctypes.c_longlong
    fake.fmov_mc_group_count.restype = ctypes.c_longlong
    fake.fmov_last_error.restype = ctypes.c_char_p
    monkeypatch.setattr(L, "lib", lambda: fake)
    monkeypatch.setattr(L, "ptr", lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr()))
    monkeypatch.setattr(L, "stream", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(mcubes_gpu, "_tables_on", set())

    class _AsCuda(torch.Tensor):          # a CPU tensor that answers is_cuda like device memory would
        is_cuda = True

    for kind, shape, iso in (("sphere", (33, 29, 31), 0.0), ("noise", (17, 16, 19), 0.1), ("noise", (2, 2, 2), 0.0),
                             ("noise", (18, 17, 20), 0.1), ("noise", (6, 24, 32), 0.1), ("sphere", (104, 101, 100), 0.0)):          # the last one spans two scan groups
        u = _field(kind, shape)
        ut = torch.from_numpy(u).as_subclass(_AsCuda)
        v, t = mcubes_gpu.marching_cubes(ut, iso)
        v_ref, t_ref = MC.marching_cubes(u, iso, T.TRI_TABLE, T.N_TRIS)
        assert v.dtype == torch.float32 and t.dtype == torch.int32
        np.testing.assert_allclose(v.numpy(), v_ref, atol=2e-5)
        np.testing.assert_array_equal(t.numpy().astype(np.int64), t_ref)
        vw, tw = mcubes_gpu.extract_geometry(ut, iso, [-1.0, -2.0, 0.0], [1.0, 2.0, 4.0])
        vw_ref, _ = MC.extract_geometry(u, iso, [-1.0, -2.0, 0.0], [1.0, 2.0, 4.0], T.TRI_TABLE, T.N_TRIS)
        assert vw.dtype == np.float64 and tw.dtype == np.int64
        np.testing.assert_allclose(vw, vw_ref, atol=1e-5)
    empty = torch.full((5, 5, 5), -1.0).as_subclass(_AsCuda)
    v, t = mcubes_gpu.marching_cubes(empty, 0.0)
    assert v.shape == (0, 3) and t.shape == (0, 3)


def test_values_exactly_on_the_isovalue(emul):
    """integer-valued fields: many corners sit exactly on the iso-value (counted as "not below", PyMCubes' `<`), vertices
    then coincide with grid corners (t = 0 or 1) and triangles may degenerate, but ids, counts and topology must still agree
    with the oracle and the surface must stay closed inside the grid"""
    rng = np.random.default_rng(4)
    for shape, iso in (((9, 8, 7), 0.0), ((6, 11, 5), 1.0), ((12, 12, 12), -1.0)):
        u = rng.integers(-2, 3, size=shape).astype(np.float32)
        v, t = emul(u, iso)
        v_ref, t_ref = MC.marching_cubes(u, iso, T.TRI_TABLE, T.N_TRIS)
        np.testing.assert_allclose(v, v_ref, atol=1e-6)
        np.testing.assert_array_equal(t.astype(np.int64), t_ref)
        assert np.isfinite(v).all()
        d = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]).astype(np.int64)
        key = np.minimum(d[:, 0], d[:, 1]) * len(v) + np.maximum(d[:, 0], d[:, 1])
        uniq, cnt = np.unique(key, return_counts=True)
        assert cnt.max() <= 2
        open_edges = uniq[cnt == 1]
        a, b = v[open_edges // len(v)], v[open_edges % len(v)]
        hi = np.asarray(shape, dtype=np.float32) - 1
        on_border = lambda p: ((p == 0) | (p == hi)).any(axis=1)
        assert on_border(a).all() and on_border(b).all()
