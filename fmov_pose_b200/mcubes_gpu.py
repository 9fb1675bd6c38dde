"""GPU marching cubes on the dense `u = -sdf` grid: the step after `extract_fields` in validate_mesh
(`mcubes.marching_cubes(u, threshold)` + rescale, models/renderer.py:43-50; PyMCubes runs it on the CPU after a 512 MiB
device->host copy).  Kernels: csrc/marching_cubes.cu; case table: mc_tables.py."""
import ctypes

import numpy as np
import torch

from . import _lib as L
from . import mc_tables

_tables_on = set()


def _ensure_tables(device):
    key = str(device)
    if key not in _tables_on:
        tri = np.ascontiguousarray(mc_tables.TRI_TABLE, dtype=np.int8)
        nt = np.ascontiguousarray(mc_tables.N_TRIS, dtype=np.uint8)
        with torch.cuda.device(device):
            L.check(L.lib().fmov_mc_set_tables(tri.ctypes.data_as(ctypes.c_void_p), nt.ctypes.data_as(ctypes.c_void_p)),
                    "fmov_mc_set_tables")
        _tables_on.add(key)


def marching_cubes(u, isovalue=0.0, scale=(1.0, 1.0, 1.0), offset=(0.0, 0.0, 0.0)):
    """u: [X,Y,Z] fp32 CUDA tensor -> (vertices fp32 [V,3] = index coordinates * scale + offset, triangles int32 [T,3]),
    both on the device.  Vertex = crossed grid edge (u < isovalue on exactly one end), shared between cells."""
    if not (torch.is_tensor(u) and u.is_cuda and u.dim() == 3):
        raise ValueError("marching_cubes takes a 3-d CUDA tensor (there is no CPU fallback)")
    u = L.f32c(u)
    X, Y, Z = (int(v) for v in u.shape)
    dev = u.device
    lib = L.lib()
    _ensure_tables(dev)
    nch = int(lib.fmov_mc_chunk_count(X, Y, Z))
    counts = torch.empty(2, nch, dtype=torch.int32, device=dev)
    work = torch.empty(nch + 1, dtype=torch.int32, device=dev)          # [n_list | list of non-empty chunks]
    n_list, lst = work[:1], work[1:]
    # int64: [scan scratch | totals (V, T) | vertex offsets | triangle offsets]; offsets are exclusive prefix sums with the
    # totals appended (chunk c emits [off[c], off[c + 1]), the emit kernels skip chunks whose range is empty)
    ng = int(lib.fmov_mc_group_count(X, Y, Z))
    w64 = torch.empty(ng + 2 + 2 * (nch + 1), dtype=torch.int64, device=dev)
    gsum, totals, excl = w64[:ng], w64[ng:ng + 2], w64[ng + 2:].view(2, nch + 1)
    L.check(lib.fmov_mc_count(L.ptr(u), X, Y, Z, L.c_float(isovalue), L.ptr(counts[0]), L.ptr(counts[1]), L.ptr(lst),
                              L.ptr(n_list), L.stream()), "fmov_mc_count")
    L.check(lib.fmov_mc_scan(L.ptr(counts[0]), L.ptr(counts[1]), L.c_ll(nch), L.ptr(gsum), L.ptr(excl[0]),
                             L.ptr(excl[1]), L.ptr(totals), L.stream()), "fmov_mc_scan")
    n_v, n_t = (int(v) for v in totals.tolist())          # the one host sync: output sizes are data dependent
    verts = torch.empty(n_v, 3, dtype=torch.float32, device=dev)
    tris = torch.empty(n_t, 3, dtype=torch.int32, device=dev)
    if n_v == 0:
        return verts, tris
    vid3 = torch.empty(X * Y * Z * 3, dtype=torch.int32, device=dev)          # written only where an edge is crossed
    s, o = [float(v) for v in scale], [float(v) for v in offset]
    L.check(lib.fmov_mc_vertices(L.ptr(u), X, Y, Z, L.c_float(isovalue), L.ptr(excl[0]), L.ptr(lst), L.ptr(n_list),
                                 L.c_float(s[0]), L.c_float(s[1]),
                                 L.c_float(s[2]), L.c_float(o[0]), L.c_float(o[1]), L.c_float(o[2]), L.ptr(verts),
                                 L.ptr(vid3), L.stream()), "fmov_mc_vertices")
    if n_t:
        L.check(lib.fmov_mc_triangles(L.ptr(u), X, Y, Z, L.c_float(isovalue), L.ptr(excl[1]), L.ptr(lst), L.ptr(n_list),
                                      L.ptr(vid3), L.ptr(tris), L.stream()), "fmov_mc_triangles")
    return verts, tris


def extract_geometry(u, threshold, bound_min, bound_max):
    """models/renderer.py:40-51 on a device grid -> (vertices float64 numpy [V,3] in world coordinates, triangles int64
    numpy [T,3]) like the reference's return value (the only device->host copies are the mesh itself)."""
    res = [int(v) for v in u.shape]
    b_min = [float(v) for v in bound_min]
    b_max = [float(v) for v in bound_max]
    scale = [(b_max[a] - b_min[a]) / (res[a] - 1.0) for a in range(3)]
    v, t = marching_cubes(u, float(threshold), scale=scale, offset=b_min)
    return v.double().cpu().numpy(), t.long().cpu().numpy()
