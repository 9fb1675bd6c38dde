"""Marching cubes (reference call: models/renderer.py:43 -> PyMCubes, absent here): the derived case
table (fmov_pose_b200/mc_tables.py) and the numpy oracle (oracle/marching_cubes.py) against the published algorithm's
known rows and against analytic properties of closed surfaces.  CPU only."""
import numpy as np
import pytest

from fmov_pose_b200 import mc_tables as T
from oracle import marching_cubes as MC


def _grid(n, lo=-1.0, hi=1.0):
    ax = np.linspace(lo, hi, n)
    return np.meshgrid(ax, ax, ax, indexing="ij")


def test_case_table_uses_exactly_the_crossed_edges_and_closes_every_loop():
    for case in range(256):
        marked = [(case >> n) & 1 for n in range(8)]
        crossed = {e for e, (a, b) in enumerate(T.EDGES) if marked[a] != marked[b]}
        tri = T.TRI_TABLE[case][:3 * T.N_TRIS[case]].reshape(-1, 3)
        assert set(tri.ravel().tolist()) == crossed, case
        assert T.EDGE_MASK[case] == sum(1 << e for e in crossed), case
        assert (T.TRI_TABLE[case][3 * T.N_TRIS[case]:] == -1).all()
        # a polygon with k corners gives k-2 triangles; every cell-interior triangle edge is used twice, in opposite
        # directions; the remaining (boundary) edges are the face segments: one per crossed edge
        d = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]) if len(tri) else np.zeros((0, 2), int)
        seen = {}
        for a, b in d.tolist():
            seen[(a, b)] = seen.get((a, b), 0) + 1
        assert all(v == 1 for v in seen.values()), case
        boundary = [(a, b) for (a, b) in seen if (b, a) not in seen]
        assert len(boundary) == len(crossed), case
        assert sorted(a for a, _ in boundary) == sorted(crossed) and sorted(b for _, b in boundary) == sorted(crossed), case
    assert T.N_TRIS.max() == T.MAX_TRIS == 5
    assert T.N_TRIS[0] == T.N_TRIS[255] == 0


def test_case_table_reproduces_the_classic_rows_where_they_are_unambiguous():
    """rows of the classic (Lorensen-Cline / PyMCubes) table, compared as oriented polygons (cyclic order)"""
    def cyc(tri):
        tri = list(tri)
        i = tri.index(min(tri))
        return tuple(tri[i:] + tri[:i])
    classic = {1: [(0, 8, 3)], 2: [(0, 1, 9)], 4: [(1, 2, 10)], 8: [(3, 11, 2)], 16: [(4, 7, 8)], 32: [(9, 5, 4)],
               64: [(10, 6, 5)], 128: [(7, 6, 11)], 5: [(0, 8, 3), (1, 2, 10)], 254: [(0, 3, 8)], 253: [(0, 9, 1)]}
    for case, tris in classic.items():
        mine = T.TRI_TABLE[case][:3 * T.N_TRIS[case]].reshape(-1, 3)
        assert sorted(cyc(t) for t in mine.tolist()) == sorted(cyc(t) for t in tris), case
    # quads of the classic table (cases 3, 6, 9, 15): same oriented loop, possibly split along the other diagonal
    for case, loop in {3: (1, 9, 8, 3), 6: (0, 2, 10, 9), 15: (8, 11, 10, 9)}.items():
        mine = T.TRI_TABLE[case][:6].reshape(2, 3)
        edges = {(a, b) for t in mine.tolist() for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0]))}
        outer = {(a, b) for (a, b) in edges if (b, a) not in edges}
        k = len(loop)
        assert outer == {(loop[i], loop[(i + 1) % k]) for i in range(k)}, case


def test_single_corner_cases_face_the_marked_corner():
    for n in range(8):
        tri = T.TRI_TABLE[1 << n][:3]
        p = [0.5 * (T.CORNERS[T.EDGES[e][0]] + T.CORNERS[T.EDGES[e][1]]) for e in tri]
        normal = np.cross(p[1] - p[0], p[2] - p[0])
        assert np.dot(normal, T.CORNERS[n] - p[0]) > 0


def test_interpolation_and_vertex_count():
    u = np.zeros((2, 2, 2))
    u[0] = -1.0
    u[1] = 3.0
    v, t = MC.marching_cubes(u, 0.0, T.TRI_TABLE, T.N_TRIS)
    assert v.shape == (4, 3) and np.allclose(v[:, 0], 0.25) and len(t) == 2
    rng = np.random.default_rng(0)
    u = rng.standard_normal((9, 8, 7))
    v, t = MC.marching_cubes(u, 0.1, T.TRI_TABLE, T.N_TRIS)
    b = u < 0.1
    n_cross = (b[:-1] != b[1:]).sum() + (b[:, :-1] != b[:, 1:]).sum() + (b[:, :, :-1] != b[:, :, 1:]).sum()
    assert len(v) == n_cross
    assert t.min() >= 0 and t.max() < len(v)


def test_plane_is_reproduced_exactly():
    xx, yy, zz = _grid(12)
    u = 0.3 * xx - 0.5 * yy + 0.8 * zz + 0.05
    v, t = MC.extract_geometry(u, 0.0, [-1, -1, -1], [1, 1, 1], T.TRI_TABLE, T.N_TRIS)
    assert len(t) > 0
    np.testing.assert_allclose(0.3 * v[:, 0] - 0.5 * v[:, 1] + 0.8 * v[:, 2] + 0.05, 0.0, atol=1e-12)


@pytest.mark.parametrize("n", [24, 41])
def test_sphere_is_closed_oriented_and_has_the_right_size(n):
    xx, yy, zz = _grid(n, -1.01, 1.01)
    r = 0.5
    u = r - np.sqrt(xx ** 2 + yy ** 2 + zz ** 2)          # u = -sdf, as renderer.py:506
    v, t = MC.extract_geometry(u, 0.0, [-1.01] * 3, [1.01] * 3, T.TRI_TABLE, T.N_TRIS)
    boundary, nonmanifold, consistent = MC.edge_manifold_report(t)
    assert boundary == 0 and nonmanifold == 0 and consistent
    assert MC.euler_characteristic(len(v), t) == 2
    h = 2.02 / (n - 1)
    np.testing.assert_allclose(np.linalg.norm(v, axis=1), r, atol=h * h)          # linear interpolation of a smooth field
    vol = MC.signed_volume(v, t)                                                    # > 0: normals point outwards (u < 0 side)
    assert 0.97 * 4 / 3 * np.pi * r ** 3 < vol < 4 / 3 * np.pi * r ** 3


def test_torus_has_genus_one():
    xx, yy, zz = _grid(48)
    u = 0.2 - np.sqrt((np.sqrt(xx ** 2 + yy ** 2) - 0.6) ** 2 + zz ** 2)
    v, t = MC.marching_cubes(u, 0.0, T.TRI_TABLE, T.N_TRIS)
    boundary, nonmanifold, consistent = MC.edge_manifold_report(t)
    assert boundary == 0 and nonmanifold == 0 and consistent
    assert MC.euler_characteristic(len(v), t) == 0


def test_noise_field_has_no_cracks():
    """white noise makes ambiguous faces everywhere: the surface must still be closed (open edges only on the grid's
    outer faces) and consistently oriented, which is exactly what the face rule of mc_tables.py guarantees"""
    rng = np.random.default_rng(1)
    n = 14
    u = rng.standard_normal((n, n, n))
    v, t = MC.marching_cubes(u, 0.0, T.TRI_TABLE, T.N_TRIS)
    d = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    key = np.minimum(d[:, 0], d[:, 1]) * len(v) + np.maximum(d[:, 0], d[:, 1])
    uniq, cnt = np.unique(key, return_counts=True)
    assert cnt.max() == 2
    open_edges = uniq[cnt == 1]
    a, b = open_edges // len(v), open_edges % len(v)
    on_border = lambda p: ((p == 0) | (p == n - 1)).any(axis=1)
    assert on_border(v[a]).all() and on_border(v[b]).all()
    assert MC.edge_manifold_report(t)[2]
