"""The C-ABI shared library must build (nvcc cross-compiles without a GPU), load, and export every
function include/fmov_b200.h declares.  No compute calls here (CPU-only container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fmov_pose_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def declared_functions():
    src = open(os.path.join(ROOT, "include", "fmov_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fmov_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ["fmov_raygen_fwd", "fmov_raygen_bwd", "fmov_pose_fwd", "fmov_pose_bwd", "fmov_sample_coarse",
                 "fmov_sample_round", "fmov_sdf_query_points", "fmov_sdf_query_rays", "fmov_sdf_query_grid",
                 "fmov_fine_fwd", "fmov_fine_bwd", "fmov_dw", "fmov_composite_fwd", "fmov_composite_bwd",
                 "fmov_loss_fwd_bwd", "fmov_ray_reduce_bwd", "fmov_last_error", "fmov_mc_set_tables", "fmov_mc_count",
                 "fmov_mc_scan", "fmov_mc_vertices", "fmov_mc_triangles", "fmov_weight_norm_fwd", "fmov_pose_gf_fwd", "fmov_flow_fwd"]:
        assert must in names, must


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_static_queries_without_a_gpu(lib):
    lib.fmov_fine_blob_bytes.restype = ctypes.c_longlong
    lib.fmov_grad_floats.restype = ctypes.c_longlong
    lib.fmov_grad_offset.restype = ctypes.c_longlong
    lib.fmov_sdf_fwd_blob_bytes.restype = ctypes.c_longlong
    assert lib.fmov_version() >= 100
    assert lib.fmov_fine_stash_count() == 54          # 53 tile tensors + the ReLU sign words of C1..C4
    assert lib.fmov_fine_image_count() == 52
    # flat gradient buffer = effective weights + biases of both MLPs (reference shapes)
    sdf_w = 256 * 39 + 2 * 256 * 256 + 217 * 256 + 4 * 256 * 256 + 257 * 256
    sdf_b = 7 * 256 + 217 + 257
    col_w = 256 * 289 + 3 * 256 * 256 + 3 * 256
    col_b = 4 * 256 + 3
    assert lib.fmov_grad_floats() == sdf_w + sdf_b + col_w + col_b
    assert lib.fmov_grad_offset(1, 0) == sdf_w
    assert lib.fmov_sdf_fwd_blob_bytes() == 128 * (256 * 1 + 256 * 4 * 2 + 224 * 4 + 256 * 5 + 256 * 4 * 3)
    # CTA-pair engine: the same eight images half-major, each followed by its [N x 16] bias slice; FP0 is image 44
    lib.fmov_sdf_pair_blob_bytes.restype = ctypes.c_longlong
    assert lib.fmov_sdf_pair_blob_bytes() == lib.fmov_sdf_fwd_blob_bytes() + 32 * (7 * 256 + 224)
    lib.fmov_sdf_pair_blob_offset.restype = ctypes.c_longlong
    off, npad, kb = ctypes.c_longlong(), ctypes.c_int(), ctypes.c_int()
    assert lib.fmov_fine_image_info(44, ctypes.byref(off), ctypes.byref(npad), ctypes.byref(kb)) == 0
    assert off.value == lib.fmov_sdf_pair_blob_offset() and (npad.value, kb.value) == (256, 1)
    assert lib.fmov_fine_image_info(52, ctypes.byref(off), ctypes.byref(npad), ctypes.byref(kb)) != 0       # no such image
    assert lib.fmov_sdf_pair_blob_offset() + lib.fmov_sdf_pair_blob_bytes() == lib.fmov_fine_blob_bytes()    # FP0..FP7 close the blob


def test_errors_are_status_codes_not_exceptions(lib):
    lib.fmov_last_error.restype = ctypes.c_char_p
    # invalid arguments are rejected before any CUDA call
    st = lib.fmov_sample_coarse(None, None, None, ctypes.c_longlong(4), 0, 0, None, None)
    assert st == -1 and b"bad shape" in lib.fmov_last_error()
    st = lib.fmov_pose_fwd(7, None, None, None, None, None, None, None)
    assert st == -1
    # marching cubes: null grid / degenerate sizes / tables not uploaded are argument errors, not crashes
    assert lib.fmov_mc_count(None, 8, 8, 8, ctypes.c_float(0.0), None, None, None, None, None) == -1
    buf = (ctypes.c_float * 8)()
    assert lib.fmov_mc_count(buf, 1, 2, 2, ctypes.c_float(0.0), None, None, None, None, None) == -1 and b"bad grid" in lib.fmov_last_error()
    assert lib.fmov_mc_count(buf, 2, 2, 2, ctypes.c_float(0.0), None, None, None, None, None) == -1 and b"fmov_mc_set_tables" in lib.fmov_last_error()
    assert lib.fmov_mc_set_tables(None, None) == -1
    lib.fmov_mc_chunk_count.restype = ctypes.c_longlong
    assert lib.fmov_mc_chunk_count(512, 512, 512) == 512 ** 3 // 256 and lib.fmov_mc_chunk_count(3, 3, 3) == 1
    lib.fmov_mc_group_count.restype = ctypes.c_longlong
    assert lib.fmov_mc_group_count(512, 512, 512) == 128 and lib.fmov_mc_group_count(3, 3, 3) == 1
    assert lib.fmov_mc_scan(None, None, ctypes.c_longlong(1), None, None, None, None, None) == -1


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under fmov_pose_b200/ may import it (bench.py's CPU legs and
    selfcheck.smoke(), which __graft_entry__.smoke() calls, are the allowed exceptions)."""
    pkg = os.path.join(ROOT, "fmov_pose_b200")
    offenders = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py") and f != "selfcheck.py":
                if re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(dp, f)).read(), flags=re.M):
                    offenders.append(os.path.join(dp, f))
    assert not offenders, offenders
