"""CPU restatement (numpy) of the marching-cubes step of validate_mesh — TEST INFRASTRUCTURE ONLY.

Reference call site: `vertices, triangles = mcubes.marching_cubes(u, threshold)` (models/renderer.py:43) followed by the
rescale of models/renderer.py:47-50.  The algorithm lives in a third-party dependency that is absent from
/root/reference and from this image: PyMCubes 0.1.4 (requirements.txt:12).  Its published algorithm (Lorensen & Cline,
"Marching Cubes", SIGGRAPH 1987, as implemented by PyMCubes' marching_cubes for arrays) is restated here:

  * a grid corner is "marked" when u < isovalue; every grid edge whose end points differ gets ONE vertex at
    x1 + (x2 - x1) * (iso - f1) / (f2 - f1) in index coordinates (shared by the cells around the edge: the mesh is indexed)
  * every cell looks its 8-bit case up in a triangle table over its 12 edges.

PARITY UNPINNED: neither PyMCubes nor any golden mesh of the reference is available, so this oracle is anchored on the
published algorithm and on analytic properties (tests/test_marching_cubes_oracle.py): the vertex set is table-independent
and exact; the triangulation uses the case table passed in (the product's derived table, fmov_pose_b200/mc_tables.py,
whose differences from the classic table are documented there).  Output order is canonical, not PyMCubes':
vertices by (grid point x-major, axis), triangles by (cell x-major, table order)."""
import numpy as np

# cell edge e -> (offset of the grid point the edge starts at, axis); same numbering as fmov_pose_b200/mc_tables.py
_CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
_EDGES = np.array([[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]])


def marching_cubes(u, isovalue, tri_table, n_tris):
    """u [X,Y,Z] float -> (vertices float64 [V,3] in index coordinates, triangles int64 [T,3])"""
    u = np.asarray(u, dtype=np.float64)
    X, Y, Z = u.shape
    below = u < isovalue
    # ---- vertices: one per crossed grid edge, ordered by (point, axis) -------------------------------------------
    flag = np.zeros((X, Y, Z, 3), dtype=bool)
    flag[:-1, :, :, 0] = below[:-1] != below[1:]
    flag[:, :-1, :, 1] = below[:, :-1] != below[:, 1:]
    flag[:, :, :-1, 2] = below[:, :, :-1] != below[:, :, 1:]
    vid = np.cumsum(flag.ravel()).reshape(flag.shape) - 1          # id of the vertex on (point, axis) where flagged
    pi, pj, pk, ax = np.nonzero(flag)
    f1 = u[pi, pj, pk]
    f2 = u[pi + (ax == 0), pj + (ax == 1), pk + (ax == 2)]
    t = (isovalue - f1) / (f2 - f1)
    verts = np.stack([pi, pj, pk], axis=1).astype(np.float64)
    verts[np.arange(len(ax)), ax] += t
    # ---- triangles -----------------------------------------------------------------------------------------------
    case = np.zeros((X - 1, Y - 1, Z - 1), dtype=np.int64)
    for n, (dx, dy, dz) in enumerate(_CORNERS):
        case |= below[dx:X - 1 + dx, dy:Y - 1 + dy, dz:Z - 1 + dz].astype(np.int64) << n
    ci, cj, ck = np.nonzero((case != 0) & (case != 255))
    cc = case[ci, cj, ck]
    tris = []
    e_org = np.minimum(_CORNERS[_EDGES[:, 0]], _CORNERS[_EDGES[:, 1]])
    e_ax = np.argmax(np.abs(_CORNERS[_EDGES[:, 0]] - _CORNERS[_EDGES[:, 1]]), axis=1)
    nt = np.asarray(n_tris)[cc].astype(np.int64)
    for t_i in range(int(nt.max()) if len(nt) else 0):
        sel = nt > t_i
        e = np.asarray(tri_table)[cc[sel], 3 * t_i:3 * t_i + 3].astype(np.int64)          # [n,3] cell edges
        ids = vid[ci[sel, None] + e_org[e, 0], cj[sel, None] + e_org[e, 1], ck[sel, None] + e_org[e, 2], e_ax[e]]
        # keep (cell, triangle) order: remember the cell rank and the triangle slot
        tris.append((np.nonzero(sel)[0], np.full(sel.sum(), t_i), ids))
    if not tris:
        return verts, np.zeros((0, 3), dtype=np.int64)
    rank = np.concatenate([a for a, _, _ in tris])
    slot = np.concatenate([b for _, b, _ in tris])
    ids = np.concatenate([c for _, _, c in tris])
    order = np.lexsort((slot, rank))
    return verts, ids[order]


def extract_geometry(u, threshold, bound_min, bound_max, tri_table, n_tris):
    """models/renderer.py:40-51 on a given grid: marching cubes + rescale from index to world coordinates."""
    res = np.asarray(u.shape, dtype=np.float64)          # the reference's grids are cubic (one `resolution`)
    v, t = marching_cubes(u, threshold, tri_table, n_tris)
    b_min, b_max = np.asarray(bound_min, dtype=np.float64), np.asarray(bound_max, dtype=np.float64)
    return v / (res - 1.0)[None, :] * (b_max - b_min)[None, :] + b_min[None, :], t


# ---- mesh properties used by the tests -----------------------------------------------------------------------------
def edge_manifold_report(triangles):
    """-> (n_boundary_edges, n_nonmanifold_edges, consistent): every undirected edge of a closed oriented surface is
    used by exactly two triangles, once in each direction."""
    t = np.asarray(triangles, dtype=np.int64)
    d = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    key = np.minimum(d[:, 0], d[:, 1]) * (t.max() + 1 if len(t) else 1) + np.maximum(d[:, 0], d[:, 1])
    sign = np.where(d[:, 0] < d[:, 1], 1, -1)
    uniq, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    bal = np.zeros(len(uniq), dtype=np.int64)
    np.add.at(bal, inv, sign)
    return int((cnt == 1).sum()), int((cnt > 2).sum()), bool(((bal == 0) | (cnt != 2)).all())


def euler_characteristic(n_vertices, triangles):
    t = np.asarray(triangles, dtype=np.int64)
    d = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    key = np.minimum(d[:, 0], d[:, 1]) * (n_vertices + 1) + np.maximum(d[:, 0], d[:, 1])
    return n_vertices - len(np.unique(key)) + len(t)


def signed_volume(vertices, triangles):
    v = np.asarray(vertices, dtype=np.float64)
    a, b, c = v[triangles[:, 0]], v[triangles[:, 1]], v[triangles[:, 2]]
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)
