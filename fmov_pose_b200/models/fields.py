"""SDFNetwork / RenderingNetwork / SingleVarianceNetwork / NeRF — drop-in for models/fields.py.

Same constructor kwargs, parameter names and shapes (old-style weight-norm: `linK.weight_g`,
`linK.weight_v`, `linK.bias`; `variance`), same initialisation (geometric init, fields.py:47-79), so
reference checkpoints load unchanged (exp_runner.py:1109-1144).  The maths runs in the fused CUDA kernels:
the modules only own parameters and expose `effective_weights()`; `NeuSRenderer.render` is the
differentiable entry point.  The direct calls the reference makes outside render() — `.sdf(x)`,
`.gradient(x)`, `forward(x)` on raw points (exp_runner.py:1660, utils/textured_mesh.py:171) — are served
by the same kernels without autograd."""
import numpy as np
import torch
import torch.nn as nn

from .. import fine as _fine
from .. import ops as _ops
from .. import packing as _packing


class _WeightNormMLP(nn.Module):
    def _eff(self, l):
        lin = getattr(self, "lin" + str(l))
        if hasattr(lin, "weight_v"):
            return torch._weight_norm(lin.weight_v, lin.weight_g, 0), lin.bias
        return lin.weight, lin.bias

    def effective_weights(self):
        """([W_l], [b_l]) with W_l = g * v / |v| (nn.utils.weight_norm, dim=0) — differentiable.  On the GPU every layer
        comes from the one fused launch (`weight_norm.effective_weights_fused`, which `NeuSRenderer.render` uses for both
        networks at once), so all call sites see bit-identical weights; the per-layer torch expression below only serves
        parameter inspection on the CPU (shape checks in the host-logic tests)."""
        if next(self.parameters()).is_cuda:
            from ..weight_norm import effective_weights_fused
            return effective_weights_fused([self])[0]
        Ws, bs = [], []
        for l in range(self.num_layers - 1):
            W, b = self._eff(l)
            Ws.append(W)
            bs.append(b)
        return Ws, bs


class SDFNetwork(_WeightNormMLP):
    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False):
        super().__init__()
        dims = [d_in] + [d_hidden for _ in range(n_layers)] + [d_out]
        self.multires = multires
        if multires > 0:
            dims[0] = d_in * (1 + 2 * multires)
        self.num_layers = len(dims)
        self.skip_in = tuple(skip_in)
        self.scale = scale
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:      # models/fields.py:47-79
                if l == self.num_layers - 2:
                    sgn = -1.0 if inside_outside else 1.0
                    nn.init.normal_(lin.weight, mean=sgn * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                    nn.init.constant_(lin.bias, -sgn * bias)
                elif multires > 0 and l == 0:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.constant_(lin.weight[:, 3:], 0.0)
                    nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                elif multires > 0 and l in self.skip_in:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
                else:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)

    # ---- kernel-backed, non-differentiable direct calls ------------------------------------------------
    def _query_weights(self, precise=False):
        W, b = self.effective_weights()
        return _packing.SdfQueryWeights([w.detach() for w in W], [x.detach() for x in b], precise=precise is True)

    def sdf(self, x, precise="act"):
        """[N,3] -> [N,1] (models/fields.py:106-107). No autograd: use NeuSRenderer.render for training.
        `precise`: "act" (default) = activation-split chain, activations and encoded inputs as fp16 hi + lo: <= 5e-4 of the
        fp32 network up to |x| = 2 (north_star: SDF <= 1e-3); True = full split-precision chain (weights split too): ~1e-5,
        1.33x the time; False = plain fp16 chain: 1.9x faster than "act", 7e-4 on the +-1.01 box, 1e-3 at |x| ~ 2."""
        with torch.no_grad():
            return _ops.sdf_query_points(self._query_weights(precise), x.reshape(-1, 3), in_scale=float(self.scale),
                                         out_scale=1.0 / float(self.scale), precise=precise)

    def _value_normal_feat(self, x, col=None):
        if float(self.scale) != 1.0:
            raise NotImplementedError("fine-stage kernels support scale == 1.0 (every shipped conf)")
        with torch.no_grad():
            W, b = self.effective_weights()
            if col is None:
                Wc = [torch.zeros(256, 289, device=x.device)] + [torch.zeros(256, 256, device=x.device)] * 3 + \
                     [torch.zeros(3, 256, device=x.device)]
                bc = [torch.zeros(256, device=x.device)] * 4 + [torch.zeros(3, device=x.device)]
            else:
                Wc, bc = col.effective_weights()
            fw = _fine.FineWeights(W, b, Wc, bc, need_backward=False)
            x = x.reshape(-1, 3).float().contiguous()
            N = x.shape[0]
            st = _fine.Stash(N, x.device, with_backward=False)
            # points as degenerate rays: o = x, d = 0, one sample each
            sdf, nrm, rgb, _ = _fine.fine_forward(fw, st, x, torch.zeros_like(x), torch.zeros(N, 1, device=x.device), 0.0)
            feat = _packing.ti_to_rowmajor(st.tensors[9], N, 256, 4)
            st.release()
        return sdf.view(-1, 1), nrm, feat, rgb

    def forward(self, inputs):
        """[N,3] -> [N,257] = [sdf, feature] (models/fields.py:88-104); features carry fp16 rounding."""
        sdf, _, feat, _ = self._value_normal_feat(inputs)
        return torch.cat([sdf, feat], dim=-1)

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def gradient(self, x):
        """d sdf / dx, [N,1,3] (models/fields.py:112-124) — analytic reverse sweep in the fine kernel."""
        _, nrm, _, _ = self._value_normal_feat(x)
        return nrm.unsqueeze(1)


class RenderingNetwork(_WeightNormMLP):
    def __init__(self, d_feature, mode, d_in, d_out, d_hidden, n_layers, weight_norm=True, multires_view=0,
                 squeeze_out=True):
        super().__init__()
        self.mode = mode
        self.squeeze_out = squeeze_out
        self.multires_view = multires_view
        dims = [d_in + d_feature] + [d_hidden for _ in range(n_layers)] + [d_out]
        if multires_view > 0:
            dims[0] += 3 * (1 + 2 * multires_view) - 3
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)

    def forward(self, points, normals, view_dirs, feature_vectors):
        raise NotImplementedError(
            "RenderingNetwork is evaluated inside the fused fine-stage kernel (NeuSRenderer.render / "
            ".extract_color); a stand-alone forward has no kernel — there is no CPU/eager fallback by design")


class SingleVarianceNetwork(nn.Module):
    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        # models/fields.py:293-294
        return torch.ones([len(x), 1], device=self.variance.device) * torch.exp(self.variance * 10.0)


class NeRF(nn.Module):
    """Background NeRF++ model (models/fields.py:197-285): parameters only.  Every shipped conf has
    n_outside = 0, so it never executes; it must exist for exp_runner.py:178 and for
    checkpoint keys."""

    def __init__(self, D=8, W=256, d_in=3, d_in_view=3, multires=0, multires_view=0, output_ch=4, skips=[4],
                 use_viewdirs=False):
        super().__init__()
        self.D, self.W, self.d_in, self.d_in_view = D, W, d_in, d_in_view
        self.input_ch = d_in * (1 + 2 * multires) if multires > 0 else 3
        self.input_ch_view = d_in_view * (1 + 2 * multires_view) if multires_view > 0 else 3
        self.skips = skips
        self.use_viewdirs = use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(self.input_ch, W)]
            + [nn.Linear(W, W) if i not in skips else nn.Linear(W + self.input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(self.input_ch_view + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)

    def forward(self, input_pts, input_views):
        raise NotImplementedError("n_outside > 0 (NeRF++ background) is out of scope: no shipped conf enables it")
