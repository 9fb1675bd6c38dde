"""GPU marching cubes (csrc/marching_cubes.cu through the C-ABI) against the numpy oracle on identical grids, and at the
full 512^3 size of BASELINE config C5 through size-independent properties of a closed surface.  (Named to run after the
train-step suites: these kernels were written after round 1's GPU budget was spent.)"""
import numpy as np
import pytest
import torch

from fmov_pose_b200 import mc_tables as T
from oracle import marching_cubes as MC

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _field(kind, shape, seed=0):
    X, Y, Z = shape
    ax = [np.linspace(-1.0, 1.0, n) for n in shape]
    xx, yy, zz = np.meshgrid(*ax, indexing="ij")
    if kind == "sphere":
        return (0.55 - np.sqrt(xx ** 2 + yy ** 2 + zz ** 2)).astype(np.float32)
    if kind == "torus":
        return (0.2 - np.sqrt((np.sqrt(xx ** 2 + yy ** 2) - 0.6) ** 2 + zz ** 2)).astype(np.float32)
    if kind == "waves":
        return (np.sin(5 * xx) * np.cos(4 * yy) + np.sin(3 * zz + 0.3) * 0.7 + 0.1).astype(np.float32)
    rng = np.random.default_rng(seed)
    return rng.standard_normal(shape).astype(np.float32)          # ambiguous faces everywhere


@pytest.mark.parametrize("kind,shape,iso", [("sphere", (33, 29, 31), 0.0), ("torus", (40, 40, 24), 0.0),
                                            ("waves", (37, 18, 50), 0.05), ("noise", (17, 16, 19), 0.0),
                                            ("noise", (2, 2, 2), 0.0), ("sphere", (9, 300, 7), 0.0),
                                            # Z % 4 == 0: quad count kernel; the last two span 2 and 3 scan groups
                                            ("noise", (9, 7, 4), 0.0), ("noise", (2, 2, 8), 0.1), ("waves", (21, 18, 52), 0.05),
                                            ("noise", (30, 31, 32), 0.2), ("sphere", (104, 101, 100), 0.0),
                                            ("waves", (131, 127, 129), 0.05),
                                            # Y * Z % 256 == 0: marching count kernel (block ranges start mid-column in
                                            # the larger ones; (4, 2, 512): one row is two chunks).  The iso-value of the
                                            # large noise grid is exact in fp32: the C ABI takes a float, the oracle a
                                            # double, and 1.2e-8 / |f1 - f0| would exceed atol on its flattest edges
                                            ("noise", (5, 16, 16), 0.0), ("noise", (2, 8, 64), 0.1), ("noise", (4, 2, 512), 0.3),
                                            ("waves", (19, 40, 32), 0.05), ("noise", (70, 64, 64), 0.25),
                                            ("sphere", (129, 128, 128), 0.0)])
def test_mesh_equals_the_oracle(kind, shape, iso):
    from fmov_pose_b200 import mcubes_gpu
    u = _field(kind, shape)
    v_ref, t_ref = MC.marching_cubes(u, iso, T.TRI_TABLE, T.N_TRIS)
    v, t = mcubes_gpu.marching_cubes(torch.from_numpy(u).to(DEV), iso)
    assert v.dtype == torch.float32 and t.dtype == torch.int32
    assert tuple(v.shape) == v_ref.shape and tuple(t.shape) == t_ref.shape
    np.testing.assert_allclose(v.cpu().numpy(), v_ref, rtol=0, atol=2e-5)          # fp32 interpolation vs fp64
    np.testing.assert_array_equal(t.cpu().numpy().astype(np.int64), t_ref)         # indices: bit-exact


def test_count_kernels_agree_on_a_misaligned_grid():
    """the same Z % 4 == 0 grid at a 16-byte aligned and at a 4-byte aligned address: quad and per-point count kernels"""
    from fmov_pose_b200 import mcubes_gpu
    u = _field("noise", (24, 20, 36))
    buf = torch.empty(u.size + 1, dtype=torch.float32, device=DEV)
    shifted = buf[1:].view(u.shape)
    shifted.copy_(torch.from_numpy(u))
    assert shifted.data_ptr() % 16 == 4
    v0, t0 = mcubes_gpu.marching_cubes(torch.from_numpy(u).to(DEV), 0.1)
    v1, t1 = mcubes_gpu.marching_cubes(shifted, 0.1)
    assert torch.equal(v0, v1) and torch.equal(t0, t1) and t0.shape[0] > 1000


def test_empty_and_world_coordinates():
    from fmov_pose_b200 import mcubes_gpu
    u = torch.full((8, 8, 8), -1.0, device=DEV)
    v, t = mcubes_gpu.marching_cubes(u, 0.0)
    assert v.shape == (0, 3) and t.shape == (0, 3)
    f = _field("sphere", (24, 24, 24))
    vw, tw = mcubes_gpu.extract_geometry(torch.from_numpy(f).to(DEV), 0.0, [-1.0, -2.0, 0.0], [1.0, 2.0, 4.0])
    v_ref, t_ref = MC.extract_geometry(f, 0.0, [-1.0, -2.0, 0.0], [1.0, 2.0, 4.0], T.TRI_TABLE, T.N_TRIS)
    assert vw.dtype == np.float64 and tw.dtype == np.int64
    np.testing.assert_allclose(vw, v_ref, atol=1e-5)
    np.testing.assert_array_equal(tw, t_ref)


def test_renderer_extract_geometry_on_the_network_grid():
    """NeuSRenderer.extract_geometry (models/renderer.py:500-507): grid query + marching cubes on the device == the oracle's
    marching cubes on the very same grid; geometric init => a closed surface near the sphere of radius 0.5"""
    from fmov_pose_b200 import synthetic
    scene = synthetic.build_scene(device=torch.device(DEV), n_images=2, H=48, W=64)
    rend = scene["renderer"]
    lo, hi = torch.tensor([-1.01] * 3, device=DEV), torch.tensor([1.01] * 3, device=DEV)
    res = 64
    verts, tris = rend.extract_geometry(lo, hi, res, threshold=0.0)
    u = rend.extract_fields(lo, hi, res).reshape(res, res, res).cpu().numpy()
    v_ref, t_ref = MC.extract_geometry(u, 0.0, [-1.01] * 3, [1.01] * 3, T.TRI_TABLE, T.N_TRIS)
    np.testing.assert_allclose(verts, v_ref, atol=1e-5)
    np.testing.assert_array_equal(tris, t_ref)
    boundary, nonmanifold, consistent = MC.edge_manifold_report(tris)
    assert boundary == 0 and nonmanifold == 0 and consistent
    r = np.linalg.norm(verts, axis=1)
    assert 0.3 < r.min() and r.max() < 0.7


def test_full_size_grid_properties():
    """512^3 (config C5): analytic sphere, closed + oriented + Euler characteristic 2 + radius within h^2"""
    from fmov_pose_b200 import mcubes_gpu
    n = 512
    ax = torch.linspace(-1.01, 1.01, n, device=DEV)
    u = 0.5 - torch.sqrt(ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2)
    h = 2.02 / (n - 1)
    v, t = mcubes_gpu.marching_cubes(u, 0.0, scale=(h, h, h), offset=(-1.01, -1.01, -1.01))
    del u
    assert int(t.min()) == 0 and int(t.max()) == v.shape[0] - 1
    r = torch.linalg.norm(v.double(), dim=1)
    assert float((r - 0.5).abs().max()) < 2 * h * h + 1e-6
    vn, tn = v.cpu().numpy(), t.cpu().numpy().astype(np.int64)
    boundary, nonmanifold, consistent = MC.edge_manifold_report(tn)
    assert boundary == 0 and nonmanifold == 0 and consistent
    assert MC.euler_characteristic(len(vn), tn) == 2
    vol = MC.signed_volume(vn, tn)
    assert abs(vol / (4 / 3 * np.pi * 0.125) - 1) < 1e-3
