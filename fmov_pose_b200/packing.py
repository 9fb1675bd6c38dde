"""Host glue that turns the fp32 *effective* weights (W = g*v/|v|, weight-norm applied in torch
so its backward stays in autograd — SURVEY.md §8b "Ownership") into the fp16/bf16 operand
images the tcgen05 kernels consume."""
import ctypes
import math

import torch

from . import _lib as L

SQ2 = math.sqrt(2.0)


def _ints(v):
    return (ctypes.c_int * len(v))(*v)


def pack_image(src, dst_u8, byte_off, npad, kblocks, segs, scale=1.0, bf16=False, n_valid=None, row_off=0,
               transpose=False):
    """src: 2-D fp32 CUDA tensor [rows, cols]. Image rows n <- src rows (or cols if transpose);
    segs = [(dst_k0, src_k0, len), ...] maps image K columns to source columns (rows if transpose)."""
    assert src.is_cuda and src.dtype == torch.float32 and src.dim() == 2
    sn, sk = src.stride(0), src.stride(1)
    rows = src.shape[0]
    if transpose:
        sn, sk = sk, sn
        rows = src.shape[1]
    if n_valid is None:
        n_valid = min(rows - row_off, npad)
    nbytes = npad * 128 * kblocks
    assert byte_off % 1024 == 0 and byte_off + nbytes <= dst_u8.numel()
    d = [s[0] for s in segs]
    s_ = [s[1] for s in segs]
    ln = [s[2] for s in segs]
    st = L.lib().fmov_pack_image(L.ptr(src), L.c_ll(sn), L.c_ll(sk), n_valid, row_off, len(segs), _ints(d), _ints(s_),
                                 _ints(ln), L.c_float(scale), int(bf16), L.c_void_p(dst_u8.data_ptr() + byte_off),
                                 npad, kblocks, L.stream())
    L.check(st, "fmov_pack_image")
    return nbytes


def ti_from_rowmajor(x, kblocks, bf16=False):
    """[P, cols] fp32 -> tile image (uint8 tensor)."""
    x = L.f32c(x)
    P, cols = x.shape
    nt = (P + 127) // 128
    out = torch.empty(nt * kblocks * 16384, dtype=torch.uint8, device=x.device)
    L.check(L.lib().fmov_ti_from_rowmajor(L.ptr(x), L.c_ll(P), cols, x.stride(0), kblocks, int(bf16), L.ptr(out),
                                          L.stream()), "fmov_ti_from_rowmajor")
    return out


def ti_to_rowmajor(img, P, cols, kblocks, bf16=False):
    out = torch.zeros(P, cols, dtype=torch.float32, device=img.device)
    L.check(L.lib().fmov_ti_to_rowmajor(L.ptr(img), L.c_ll(P), cols, cols, kblocks, int(bf16), L.ptr(out), L.stream()),
            "fmov_ti_to_rowmajor")
    return out


class SdfQueryWeights:
    """Forward images of lin0..lin7 + fp32 side arrays for fmov_sdf_query_* (value-only chain)."""

    @classmethod
    def from_views(cls, blob, bias, w8, b8):
        self = cls.__new__(cls)
        self.blob, self.bias, self.w8, self.b8 = blob, bias, w8, b8
        self.blob_lo = None
        self.blob_pair = None
        return self

    def __init__(self, W, b, precise=False):
        """W, b: lists of the 9 effective fp32 weights / biases of the SDF net (CUDA).  `precise`: also pack the fp16
        residual images (`blob_lo`) that the split-precision chain (fmov_sdf_query_*_precise) adds in."""
        dev = W[0].device
        lib = L.lib()
        assert len(W) == 9 and W[0].shape == (256, 39) and W[3].shape == (217, 256) and W[8].shape[1] == 256, \
            "kernels are built for the 8x256, multires=6, skip_in=(4,) SDF network of the shipped confs"
        self.blob = torch.zeros(int(lib.fmov_sdf_fwd_blob_bytes()), dtype=torch.uint8, device=dev)
        self.blob_lo = torch.zeros_like(self.blob) if precise else None
        self.blob_pair = None          # only FineWeights carries the half-major images of the CTA-pair engine
        for fmt, blob in ((0, self.blob), (2, self.blob_lo)):
            if blob is None:
                continue
            for l in range(8):
                off = int(lib.fmov_sdf_fwd_blob_offset(l))
                Wl = W[l].detach().float()
                if l == 0:
                    pack_image(Wl, blob, off, 256, 1, [(0, 0, 39), (39, 0, 3)], bf16=fmt)          # 39..41: residuals of x, y, z
                elif l == 3:
                    pack_image(Wl, blob, off, 224, 4, [(0, 0, 256)], bf16=fmt)
                elif l == 4:
                    pack_image(Wl, blob, off, 256, 5, [(0, 0, 217), (256, 217, 39), (295, 217, 3)], scale=1.0 / SQ2, bf16=fmt)
                else:
                    pack_image(Wl, blob, off, 256, 4, [(0, 0, 256)], bf16=fmt)
        self.bias = torch.zeros(8, 256, dtype=torch.float32, device=dev)
        for l in range(8):
            self.bias[l, : b[l].numel()] = b[l].detach().float()
        self.w8 = W[8][0].detach().float().contiguous()
        self.b8 = b[8].detach().float().contiguous()
