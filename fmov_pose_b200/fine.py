"""Host side of the fine stage: weight-image packing, stash allocation and the autograd Function
that strings the CUDA kernels together (fine_fwd -> composite_fwd | composite_bwd -> fine_bwd -> dw ->
ray_reduce).  Mirrors the reference's render_core (models/renderer.py:244-372) + loss.backward()."""
import ctypes
import math

import torch

from . import _lib as L
from . import ops, packing

SQ2 = math.sqrt(2.0)

# image ids (csrc/mlp_fine.cu ImgId)
IMG_F0, IMG_T0, IMG_C0 = 0, 9, 17
IMG_CT0A, IMG_CT0B, IMG_CT1, IMG_CT2, IMG_CT3 = 22, 23, 24, 25, 26
IMG_FB0, IMG_TB0 = 27, 35


def _img_info(i):
    off, npad, kb = ctypes.c_longlong(), ctypes.c_int(), ctypes.c_int()
    L.check(L.lib().fmov_fine_image_info(i, ctypes.byref(off), ctypes.byref(npad), ctypes.byref(kb)), "fmov_fine_image_info")
    return off.value, npad.value, kb.value


def check_supported(W_sdf, W_col):
    ok = (len(W_sdf) == 9 and len(W_col) == 5 and tuple(W_sdf[0].shape) == (256, 39)
          and tuple(W_sdf[3].shape) == (217, 256) and tuple(W_sdf[8].shape) == (257, 256)
          and tuple(W_col[0].shape) == (256, 289) and tuple(W_col[4].shape) == (3, 256)
          and all(tuple(W_sdf[l].shape) == (256, 256) for l in (1, 2, 4, 5, 6, 7))
          and all(tuple(W_col[l].shape) == (256, 256) for l in (1, 2, 3)))
    if not ok:
        raise NotImplementedError(
            "fmov_pose_b200 kernels are built for the network shapes of the shipped confs: SDFNetwork "
            "d_hidden=256, n_layers=8, skip_in=(4,), multires=6, d_out=257; RenderingNetwork mode='idr', "
            "d_feature=256, d_hidden=256, n_layers=4, multires_view=4, d_out=3 (confs/ho3d_*.conf)")


class FineWeights:
    """All operand images + fp32 side arrays for the MLP kernels, packed from the effective weights by ONE
    launch (fmov_pack_all).  The first fmov_sdf_fwd_blob_bytes() bytes are exactly the value-chain blob of
    fmov_sdf_query_*, so hierarchical sampling shares it (`.query` is a packing.SdfQueryWeights view)."""

    def __init__(self, W_sdf, b_sdf, W_col, b_col, need_backward=True):
        check_supported(W_sdf, W_col)
        lib = L.lib()
        dev = W_sdf[0].device
        srcs = [t.detach().float().contiguous() for t in (list(W_sdf) + list(b_sdf) + list(W_col) + list(b_col))]
        self._keep = srcs
        self.blob = torch.empty(int(lib.fmov_fine_blob_bytes()), dtype=torch.uint8, device=dev)
        self.side = torch.empty(int(lib.fmov_side_floats()), dtype=torch.float32, device=dev)
        ptrs = (ctypes.c_void_p * 28)(*[t.data_ptr() for t in srcs])
        L.check(lib.fmov_pack_all(ptrs, L.ptr(self.blob), L.ptr(self.side), int(bool(need_backward)), L.stream()),
                "fmov_pack_all")
        o = [int(lib.fmov_side_offset(i)) for i in range(6)]
        self.bias_sdf = self.side[o[0]: o[0] + 2048]
        self.b8 = self.side[o[1]: o[1] + 257]
        self.w8row = self.side[o[2]: o[2] + 256]
        self.bias_col = self.side[o[3]: o[3] + 1024]
        self.bc4 = self.side[o[4]: o[4] + 3]
        self.wc4 = self.side[o[5]: o[5] + 768]
        self.query = packing.SdfQueryWeights.from_views(self.blob, self.bias_sdf, self.w8row, self.b8)
        nb = int(lib.fmov_sdf_pair_blob_bytes())
        if nb > 0:          # half-major copies FP0..FP7 (+ bias slices): the sampling queries run on the CTA-pair engine
            off = int(lib.fmov_sdf_pair_blob_offset())
            self.query.blob_pair = self.blob[off: off + nb]


class Stash:
    """Tile-image tensors in HBM (activations of the forward, gradient tiles of the backward) carved out of ONE
    buffer.  Buffers are recycled through a small free-list so that a train loop does not allocate/free tens of
    GB per step (a Stash is returned to the pool at the end of backward or when its autograd node dies)."""
    _pool = {}

    def __init__(self, P, device, with_backward=True):
        lib = L.lib()
        n = lib.fmov_fine_stash_count()
        # + 1: the padding tile that the odd CTA of a pair writes when it has one tile less than its partner (csrc/mlp_fine.cu)
        self.nt = (P + 127) // 128 + 1
        self.device = device
        self.blocks = [lib.fmov_fine_stash_blocks(i) for i in range(n)]
        # buffer order: the tensors the forward kernel writes first, so that a forward-only stash is a prefix
        self.order = sorted(range(n), key=lambda i: (0 if lib.fmov_fine_stash_is_forward(i) else 1, i))
        n_fwd = sum(self.blocks[i] for i in range(n) if lib.fmov_fine_stash_is_forward(i))
        n_all = sum(self.blocks)
        self.key = (str(device), self.nt)
        self.buf = None
        free = Stash._pool.get(self.key, [])
        while free:
            cand = free.pop()
            if cand.numel() >= (n_all if with_backward else n_fwd) * self.nt * 16384:
                self.buf = cand
                break
        if self.buf is None:
            self.buf = torch.empty((n_all if with_backward else n_fwd) * self.nt * 16384, dtype=torch.uint8, device=device)
        self._carve()

    def _carve(self):
        have = self.buf.numel() // (self.nt * 16384)
        self.tensors, off = [None] * len(self.blocks), 0
        for i in self.order:
            nb = self.blocks[i]
            if off + nb <= have:
                self.tensors[i] = self.buf[off * self.nt * 16384: (off + nb) * self.nt * 16384]
            off += nb
        self.ptrs = (ctypes.c_void_p * len(self.blocks))(*[(t.data_ptr() if t is not None else 0) for t in self.tensors])

    def ensure_backward(self, P, device):
        if any(t is None for t in self.tensors):
            # forward-only stash asked to run backward: grow (keeps the forward part)
            big = torch.empty(sum(self.blocks) * self.nt * 16384, dtype=torch.uint8, device=device)
            big[: self.buf.numel()].copy_(self.buf)
            self.buf = big
            self._carve()

    def release(self):
        if self.buf is not None:
            lst = Stash._pool.setdefault(self.key, [])
            if len(lst) < 2:
                lst.append(self.buf)
            self.buf, self.tensors = None, []

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def fine_forward(fw, stash, rays_o, rays_d, z, sample_dist):
    B, S = z.shape
    dev = z.device
    P = B * S
    sdf = torch.empty(P, dtype=torch.float32, device=dev)
    nrm = torch.empty(P, 3, dtype=torch.float32, device=dev)
    rgb = torch.empty(P, 3, dtype=torch.float32, device=dev)
    ge = torch.empty(-(-P // 128) * 128 * 40, dtype=torch.float32, device=dev)      # opaque: [tile][40][128]
    with L.timed("fine_fwd"):
      L.check(L.lib().fmov_fine_fwd(L.c_ll(B), S, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.c_float(sample_dist),
                                  L.ptr(fw.blob), stash.ptrs, L.ptr(fw.bias_sdf), L.ptr(fw.b8), L.ptr(fw.w8row),
                                  L.ptr(fw.bias_col), L.ptr(fw.bc4), L.ptr(sdf), L.ptr(nrm), L.ptr(rgb), L.ptr(ge),
                                  L.stream()), "fmov_fine_fwd")
    return sdf, nrm, rgb, ge


def fine_backward(fw, stash, rays_o, rays_d, z, sample_dist, rgb, ge, d_sdf, d_nrm, d_rgb, amax):
    B, S = z.shape
    dev = z.device
    P = B * S
    d_pts = torch.empty(P, 3, dtype=torch.float32, device=dev)
    d_dirs = torch.empty(P, 3, dtype=torch.float32, device=dev)
    zc4 = torch.empty(P, 4, dtype=torch.float32, device=dev)
    eb = torch.empty(-(-P // 128) * 128 * 40, dtype=torch.float32, device=dev)      # scratch: [tile][40][128]
    with L.timed("fine_bwd"):
      L.check(L.lib().fmov_fine_bwd(L.c_ll(B), S, L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.c_float(sample_dist),
                                  L.ptr(fw.blob), stash.ptrs, L.ptr(fw.bias_sdf), L.ptr(fw.b8), L.ptr(fw.w8row),
                                  L.ptr(fw.bias_col), L.ptr(fw.bc4), L.ptr(fw.wc4), L.ptr(rgb), L.ptr(ge), L.ptr(d_sdf),
                                  L.ptr(d_nrm), L.ptr(d_rgb), L.ptr(amax), L.ptr(d_pts), L.ptr(d_dirs), L.ptr(zc4), L.ptr(eb), L.stream()),
            "fmov_fine_bwd")
    return d_pts, d_dirs, zc4


_SDF_SHAPES = [(256, 39), (256, 256), (256, 256), (217, 256), (256, 256), (256, 256), (256, 256), (256, 256), (257, 256)]
_COL_SHAPES = [(256, 289), (256, 256), (256, 256), (256, 256), (3, 256)]


def weight_grads(stash, P, d_sdf, zc4, amax):
    """-> (dW_sdf[9], db_sdf[9], dW_col[5], db_col[5]) as views of one flat fp32 buffer"""
    lib = L.lib()
    dev = d_sdf.device
    flat = torch.empty(int(lib.fmov_grad_floats()), dtype=torch.float32, device=dev)
    with L.timed("dw"):
        L.check(lib.fmov_dw(L.c_ll(P), stash.ptrs, L.ptr(d_sdf), L.ptr(zc4), L.ptr(amax), L.ptr(flat), L.stream()), "fmov_dw")

    def view(kind, l, shape):
        off = int(lib.fmov_grad_offset(kind, l))
        n = 1
        for s_ in shape:
            n *= s_
        return flat[off: off + n].view(*shape)

    dW_s = [view(0, l, _SDF_SHAPES[l]) for l in range(9)]
    db_s = [view(1, l, (_SDF_SHAPES[l][0],)) for l in range(9)]
    dW_c = [view(2, l, _COL_SHAPES[l]) for l in range(5)]
    db_c = [view(3, l, (_COL_SHAPES[l][0],)) for l in range(5)]
    return dW_s, db_s, dW_c, db_c, flat


class RenderCoreFunction(torch.autograd.Function):
    """render_core + render() reductions (models/renderer.py:244-372, 477-498) as one autograd node.

    inputs : rays_o [B,3], rays_d [B,3], z_vals [B,S], inv_s (0-d, already clipped), cfg dict,
             then 9 W_sdf, 9 b_sdf, 5 W_col, 5 b_col (effective fp32 weights)
    outputs: color, weight_sum, weight_max, depth, weights, cdf, inside_sphere, mid_z, pts, sdf, gradients,
             gradient_error
    """

    @staticmethod
    def forward(ctx, rays_o, rays_d, z_vals, inv_s, cfg, *params):
        W_sdf, b_sdf, W_col, b_col = params[0:9], params[9:18], params[18:23], params[23:28]
        rays_o, rays_d, z = L.f32c(rays_o), L.f32c(rays_d), L.f32c(z_vals)
        B, S = z.shape
        dev = z.device
        need_bwd = cfg.get("need_backward", True)
        fw = cfg.get("fine_weights") or FineWeights(W_sdf, b_sdf, W_col, b_col, need_backward=need_bwd)
        stash = Stash(B * S, dev, with_backward=need_bwd)
        sd, car = float(cfg["sample_dist"]), float(cfg["cos_anneal_ratio"])
        bg = cfg.get("background_rgb")
        bg = None if bg is None else L.f32c(bg.reshape(-1))
        inv_s_d = L.f32c(inv_s.detach().reshape(1))
        sdf, nrm, rgb, ge = fine_forward(fw, stash, rays_o, rays_d, z, sd)
        f = ops.composite_fwd(rays_o, rays_d, z, sdf, nrm, rgb, inv_s_d, sd, car, bg=bg, full=True)
        eik = f["eik"].sum(0)
        group = cfg.get("group")
        if cfg.get("eik_den") is not None:      # micro-batch of a larger step: partial numerator / whole-batch normaliser
            eik = torch.stack([eik[0], cfg["eik_den"].reshape(()).to(eik.dtype)])
        elif group is not None:
            # global eikonal normaliser; the caller's other whole-batch sums (sum(mask), ray count) ride on
            # the same collective: one all-reduce per forward instead of two
            extra = cfg.get("reduce_extra")
            if extra is not None:
                pack = torch.cat([eik, extra.reshape(-1)])
                torch.distributed.all_reduce(pack, group=group)
                eik = pack[:2]
                extra.copy_(pack[2:].view_as(extra))
            else:
                torch.distributed.all_reduce(eik, group=group)
        gradient_error = eik[0] / (eik[1] + 1e-5)
        ctx.fw, ctx.stash, ctx.cfg = fw, stash, cfg
        ctx.bg = bg
        ctx.eik_den = eik[1:2].contiguous()
        ctx.save_for_backward(rays_o, rays_d, z, inv_s_d, sdf, nrm, rgb, ge)
        outs = (f["color"], f["weight_sum"], f["weight_max"], f["depth"], f["weights"], f["cdf"], f["inside"],
                f["mid_z"], f["pts"], sdf.view(-1, 1), nrm.view(B, S, 3), gradient_error)
        ctx.set_materialize_grads(False)      # unused outputs arrive as None (NULL pointers), not as zero-filled tensors
        ctx.mark_non_differentiable(f["weight_max"], f["cdf"], f["inside"], f["mid_z"], f["pts"], outs[9])
        return outs

    @staticmethod
    def backward(ctx, g_color, g_wsum, g_wmax, g_depth, g_weights, g_cdf, g_inside, g_mid, g_pts, g_sdf, g_nrm, g_eik):
        rays_o, rays_d, z, inv_s_d, sdf, nrm, rgb, ge = ctx.saved_tensors
        cfg, fw, stash = ctx.cfg, ctx.fw, ctx.stash
        B, S = z.shape
        dev = z.device
        sd, car = float(cfg["sample_dist"]), float(cfg["cos_anneal_ratio"])
        stash.ensure_backward(B * S, dev)
        c = lambda g: None if g is None else L.f32c(g)
        g_eik_d = None if g_eik is None else L.f32c(g_eik.reshape(1))
        bk = ops.composite_bwd(rays_o, rays_d, z, sdf, nrm, rgb, inv_s_d, sd, car, ctx.bg, c(g_color), c(g_wsum),
                               c(g_depth), c(g_weights), g_eik_d, ctx.eik_den,
                               None if g_nrm is None else L.f32c(g_nrm.reshape(-1, 3)))
        amax = torch.empty(1, dtype=torch.float32, device=dev)
        L.check(L.lib().fmov_grad_amax(L.ptr(bk["d_sdf"]), L.ptr(bk["d_nrm"]), L.ptr(bk["d_rgb"]), L.c_ll(B * S), L.ptr(amax),
                                       L.stream()), "fmov_grad_amax")
        d_pts, d_dirs, zc4 = fine_backward(fw, stash, rays_o, rays_d, z, sd, rgb, ge, bk["d_sdf"], bk["d_nrm"], bk["d_rgb"], amax)
        dW_s, db_s, dW_c, db_c, flat = weight_grads(stash, B * S, bk["d_sdf"], zc4, amax)
        want_dz = ctx.needs_input_grad[2]
        d_o, d_d, d_z = ops.ray_reduce_bwd(d_pts, d_dirs, bk["d_dir"], bk["d_dist"], bk["d_mid"], rays_d, z, sd, want_dz)
        d_inv_s = bk["d_invs"].sum().reshape(())
        stash.release()           # recycle the stash as soon as the gradients exist
        ctx.stash = None
        return (d_o, d_d, d_z, d_inv_s, None) + tuple(dW_s) + tuple(db_s) + tuple(dW_c) + tuple(db_c)
