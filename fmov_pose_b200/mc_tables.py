"""Marching-cubes case table, derived (not transcribed) from the cube's geometry.

The reference hands the `u = -sdf` grid to PyMCubes (`mcubes.marching_cubes(u, threshold)`, models/renderer.py:43;
PyMCubes 0.1.4, requirements.txt:12 — third-party, not vendored, not installed here).  PyMCubes implements the classic
Lorensen & Cline marching cubes with the corner / edge numbering below and marks a corner when its value is BELOW the
iso-value.  This module rebuilds a case table for that numbering from first principles:

  * corners   v0=(0,0,0) v1=(1,0,0) v2=(1,1,0) v3=(0,1,0) v4=(0,0,1) v5=(1,0,1) v6=(1,1,1) v7=(0,1,1)   (x,y,z)
  * edges     e0=v0v1 e1=v1v2 e2=v2v3 e3=v3v0 e4=v4v5 e5=v5v6 e6=v6v7 e7=v7v4 e8=v0v4 e9=v1v5 e10=v2v6 e11=v3v7
  * case      bit n set  <=>  value(v_n) < iso
  * on every cube face the crossed edges are joined by segments; a face with four crossings (marked corners on a
    diagonal) is resolved by cutting off each MARKED corner separately.  The rule depends only on the face's own corner
    values, so both cells sharing the face agree and the surface has no cracks (the classic table resolves some of those
    faces differently on the two sides and can leave holes there; everywhere else the two tables describe the same
    polygons, possibly split along a different diagonal).
  * the segments close into loops; each loop is a polygon, oriented so that its normal points to the marked (lower
    value) side, and triangulated without diagonals that lie in a cube face (see _pick_triangulation).

The mesh VERTICES (one per crossed grid edge) do not depend on the table at all."""
import numpy as np

CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int64)
EDGES = np.array([[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]], dtype=np.int64)
# edge e of the cell at grid point (i,j,k) = the grid edge that starts at (i,j,k) + EDGE_ORIGIN[e] along axis EDGE_AXIS[e]
EDGE_ORIGIN = np.minimum(CORNERS[EDGES[:, 0]], CORNERS[EDGES[:, 1]])
EDGE_AXIS = np.argmax(np.abs(CORNERS[EDGES[:, 0]] - CORNERS[EDGES[:, 1]]), axis=1)
# faces: (axis, side) -> the four corners in cyclic order
_FACES = []
for _axis in range(3):
    for _side in (0, 1):
        _idx = [n for n in range(8) if CORNERS[n][_axis] == _side]
        _a, _b = [a for a in range(3) if a != _axis]
        _c0 = [n for n in _idx if CORNERS[n][_a] == 0 and CORNERS[n][_b] == 0][0]
        _c1 = [n for n in _idx if CORNERS[n][_a] == 1 and CORNERS[n][_b] == 0][0]
        _c2 = [n for n in _idx if CORNERS[n][_a] == 1 and CORNERS[n][_b] == 1][0]
        _c3 = [n for n in _idx if CORNERS[n][_a] == 0 and CORNERS[n][_b] == 1][0]
        _normal = np.zeros(3)
        _normal[_axis] = 1.0 if _side else -1.0
        _FACES.append(([_c0, _c1, _c2, _c3], _normal))
MAX_TRIS = 5          # checked by build_tables(): no case needs more (as in the classic table)


def _edge_between(a, b):
    for e, (p, q) in enumerate(EDGES):
        if (p == a and q == b) or (p == b and q == a):
            return e
    raise KeyError((a, b))


def _mid(e):
    return 0.5 * (CORNERS[EDGES[e][0]] + CORNERS[EDGES[e][1]]).astype(np.float64)


def _case_polygons(case):
    marked = [(case >> n) & 1 for n in range(8)]
    nxt = {}                                     # directed segments: edge -> next edge of its loop
    for corners, normal in _FACES:
        ring = [_edge_between(corners[i], corners[(i + 1) % 4]) for i in range(4)]       # ring[i] joins corner i, i+1
        crossed = [marked[corners[i]] != marked[corners[(i + 1) % 4]] for i in range(4)]
        n_cross = sum(crossed)
        segs = []
        if n_cross == 2:
            a, b = [ring[i] for i in range(4) if crossed[i]]
            toward = sum((CORNERS[c] for c in corners if marked[c]), np.zeros(3)) / max(1, sum(marked[c] for c in corners)) \
                - sum((CORNERS[c] for c in corners if not marked[c]), np.zeros(3)) / max(1, sum(1 - marked[c] for c in corners))
            segs.append((a, b, toward))
        elif n_cross == 4:                       # ambiguous face: cut off every marked corner on its own
            for i in range(4):
                if marked[corners[i]]:
                    a, b = ring[(i - 1) % 4], ring[i]                 # the two face edges that meet at corner i
                    toward = CORNERS[corners[i]] - 0.5 * (_mid(a) + _mid(b))
                    segs.append((a, b, toward))
        for a, b, toward in segs:
            # polygon normal points to the marked side and the polygon lies inside the cube, so along this face the
            # boundary runs in direction  toward x face_normal
            d = np.cross(toward, normal)
            if np.dot(_mid(b) - _mid(a), d) < 0:
                a, b = b, a
            assert a not in nxt, "edge leaves twice"
            nxt[a] = b
    loops, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, e = [], start
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == start, "open loop"
        loops.append(loop)
    return loops


def _share_face(a, b):
    """do cell edges a and b lie on a common cube face?"""
    for corners, _ in _FACES:
        on = set(corners)
        if set(EDGES[a]) <= on and set(EDGES[b]) <= on:
            return True
    return False


def _triangulations(loop):
    """all triangulations of the oriented polygon `loop` (Catalan many; polygons here have <= 8 corners)"""
    if len(loop) < 3:
        return [[]]
    if len(loop) == 3:
        return [[tuple(loop)]]
    out = []
    a, b = loop[0], loop[1]                      # the triangle on side (a, b) has its apex at loop[i]
    for i in range(2, len(loop)):
        left = _triangulations(loop[1:i + 1])    # polygon b .. loop[i]
        right = _triangulations([loop[0]] + loop[i:])
        for lt in left:
            for rt in right:
                out.append([(a, b, loop[i])] + lt + rt)
    return out


def _pick_triangulation(loop):
    """A diagonal between two vertices that sit on the same cube face lies IN that face and can coincide with a
    diagonal or segment of the neighbouring cell (an edge shared by four triangles).  Take the first triangulation (in
    the fan-first enumeration order) with the fewest such diagonals; for every polygon of the table that is zero."""
    boundary = {(loop[i], loop[(i + 1) % len(loop)]) for i in range(len(loop))}
    best, best_bad = None, None
    for tris in _triangulations(loop):
        bad = 0
        for t in tris:
            for x, y in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                if (x, y) not in boundary and (y, x) not in boundary and _share_face(x, y):
                    bad += 1
        if best is None or bad < best_bad:
            best, best_bad = tris, bad
        if bad == 0:
            break
    return best, best_bad // 2


def build_tables():
    """-> (tri_table int8 [256, 3*MAX_TRIS] of cell edge numbers, -1 padded; n_tris uint8 [256]; edge_mask uint16 [256])"""
    tri = -np.ones((256, 3 * MAX_TRIS), dtype=np.int8)
    ntri = np.zeros(256, dtype=np.uint8)
    emask = np.zeros(256, dtype=np.uint16)
    in_face = 0
    for case in range(256):
        out = []
        for loop in _case_polygons(case):
            tris, bad = _pick_triangulation(loop)
            in_face += bad
            for t in tris:
                out += list(t)
            for e in loop:
                emask[case] |= 1 << e
        assert len(out) <= 3 * MAX_TRIS, (case, len(out))
        tri[case, :len(out)] = out
        ntri[case] = len(out) // 3
    return tri, ntri, emask, in_face


TRI_TABLE, N_TRIS, EDGE_MASK, IN_FACE_DIAGONALS = build_tables()
