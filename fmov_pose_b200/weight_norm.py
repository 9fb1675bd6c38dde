"""Effective weights W = g * v / ||v|| (nn.utils.weight_norm, dim 0) of every layer of several MLPs in ONE launch, with a
one-launch backward (csrc/weight_norm.cu) — the reference recomputes them per layer inside every forward
(models/fields.py:81-82, 160-161): 14 + 14 torch launches per iteration for the SDF and colour networks."""
import ctypes

import torch

from . import _lib as L


def _parr(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _iarr(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


class FusedWeightNormFn(torch.autograd.Function):
    """(v_0, g_0, v_1, g_1, ...) -> (W_0, W_1, ...);  v_i [rows, cols], g_i [rows, 1]"""

    @staticmethod
    def forward(ctx, *vg):
        vs = [L.f32c(t.detach()) for t in vg[0::2]]
        gs = [L.f32c(t.detach()).reshape(-1) for t in vg[1::2]]
        dev = vs[0].device
        rows, cols = [v.shape[0] for v in vs], [v.shape[1] for v in vs]
        flat = torch.empty(sum(r * c for r, c in zip(rows, cols)), dtype=torch.float32, device=dev)
        norms = torch.empty(sum(rows), dtype=torch.float32, device=dev)
        Ws, ns, o, q = [], [], 0, 0
        for r, c in zip(rows, cols):
            Ws.append(flat[o: o + r * c].view(r, c))
            ns.append(norms[q: q + r])
            o, q = o + r * c, q + r
        L.check(L.lib().fmov_weight_norm_fwd(len(vs), _parr(vs), _parr(gs), _iarr(rows), _iarr(cols), _parr(Ws), _parr(ns),
                                             L.stream()), "fmov_weight_norm_fwd")
        ctx.save_for_backward(norms, *vs, *gs)
        ctx.dims = (rows, cols)
        return tuple(Ws)

    @staticmethod
    def backward(ctx, *dWs):
        rows, cols = ctx.dims
        n = len(rows)
        sv = ctx.saved_tensors
        norms, vs, gs = sv[0], list(sv[1:1 + n]), list(sv[1 + n:1 + 2 * n])
        dev = norms.device
        dWs = [torch.zeros(r, c, dtype=torch.float32, device=dev) if d is None else L.f32c(d)
               for d, r, c in zip(dWs, rows, cols)]
        flat = torch.empty(sum(r * c for r, c in zip(rows, cols)), dtype=torch.float32, device=dev)
        dgf = torch.empty(sum(rows), dtype=torch.float32, device=dev)
        dvs, dgs, ns, o, q = [], [], [], 0, 0
        for r, c in zip(rows, cols):
            dvs.append(flat[o: o + r * c].view(r, c))
            dgs.append(dgf[q: q + r])
            ns.append(norms[q: q + r])
            o, q = o + r * c, q + r
        L.check(L.lib().fmov_weight_norm_bwd(n, _parr(vs), _parr(gs), _iarr(rows), _iarr(cols), _parr(ns), _parr(dWs),
                                             _parr(dvs), _parr(dgs), L.stream()), "fmov_weight_norm_bwd")
        out = []
        for dv, dg in zip(dvs, dgs):
            out += [dv, dg.view(-1, 1)]
        return tuple(out)


def effective_weights_fused(nets):
    """[(Ws, bs) per network] like `net.effective_weights()`, with every weight-normalised layer of every network
    evaluated by one fused launch (layers without weight norm pass through)."""
    slots, vg = [], []
    for net in nets:
        for l in range(net.num_layers - 1):
            lin = getattr(net, "lin" + str(l))
            if hasattr(lin, "weight_v"):
                slots.append(None)
                vg += [lin.weight_v, lin.weight_g]
            else:
                slots.append(lin.weight)
    fused = list(FusedWeightNormFn.apply(*vg)) if vg else []
    out, k = [], 0
    for net in nets:
        Ws, bs = [], []
        for l in range(net.num_layers - 1):
            W = slots[k]
            k += 1
            Ws.append(fused.pop(0) if W is None else W)
            bs.append(getattr(net, "lin" + str(l)).bias)
        out.append((Ws, bs))
    return out
