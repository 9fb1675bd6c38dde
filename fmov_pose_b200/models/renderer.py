"""NeuSRenderer — drop-in for models/renderer.py:89-532 on the B200-native kernels.

Same constructor, attributes, `render()` signature and output dict (keys and shapes of
models/renderer.py:486-498), differentiable where the reference's outputs are: `color_fine`,
`weight_sum`, `depth_fine`, `weights`, `gradients`, `gradient_error`, `s_val` carry gradients to the SDF /
colour / variance parameters and to `rays_o`, `rays_d` (→ pose parameters); with `n_importance == 0`
also through `near` / `far` (models/renderer.py:390).  `pts` carries its gradient to `rays_o`, `rays_d` (and z).  Not supported (explicit errors, no fallback):
`n_outside > 0`, network shapes other than the shipped confs'.

The jitter draw `torch.rand([B,1])` (renderer.py:404) stays in host PyTorch so RNG parity with the
reference is preserved; pass `t_rand=` to inject a given draw."""
import numpy as np
import torch

from .. import fine as _fine
from .. import flow as _flow
from .. import mcubes_gpu as _mcubes
from .. import ops as _ops
from .. import packing as _packing
from .. import weight_norm as _wn


def extract_fields(bound_min, bound_max, resolution, query_func):
    """Generic chunked grid evaluation with the reference's traversal (models/renderer.py:9-37).
    `NeuSRenderer.extract_fields` is the fused single-launch path; this one serves arbitrary callables."""
    N = 64
    dev = bound_min.device if torch.is_tensor(bound_min) else None
    X = torch.linspace(float(bound_min[0]), float(bound_max[0]), resolution, device=dev).split(N)
    Y = torch.linspace(float(bound_min[1]), float(bound_max[1]), resolution, device=dev).split(N)
    Z = torch.linspace(float(bound_min[2]), float(bound_max[2]), resolution, device=dev).split(N)
    u = np.zeros([resolution, resolution, resolution], dtype=np.float32)
    with torch.no_grad():
        for xi, xs in enumerate(X):
            for yi, ys in enumerate(Y):
                for zi, zs in enumerate(Z):
                    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing="ij")
                    pts = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
                    val = query_func(pts).reshape(len(xs), len(ys), len(zs)).detach().cpu().numpy()
                    u[xi * N: xi * N + len(xs), yi * N: yi * N + len(ys), zi * N: zi * N + len(zs)] = val
    return u


def extract_geometry(bound_min, bound_max, resolution, threshold, query_func=None, u=None):
    """models/renderer.py:40-51: grid -> marching cubes -> vertices in world coordinates.  The reference calls PyMCubes on
    the host (`mcubes.marching_cubes`, :43); here the grid stays on the GPU and `mcubes_gpu` extracts the indexed mesh
    there (same vertices; triangulation per mc_tables.py).  Returns numpy (vertices float64 [V,3], triangles int64 [T,3])."""
    if u is None:
        u = extract_fields(bound_min, bound_max, resolution, query_func)
    if not torch.is_tensor(u):
        dev = bound_min.device if torch.is_tensor(bound_min) and bound_min.is_cuda else torch.device("cuda")
        u = torch.as_tensor(np.ascontiguousarray(u), dtype=torch.float32).to(dev)
    return _mcubes.extract_geometry(u, threshold, [float(v) for v in bound_min], [float(v) for v in bound_max])


class _CoarseZFn(torch.autograd.Function):
    """z = near + (far - near) * linspace(0, 1, n) + jitter (models/renderer.py:389-390, 403-405) through fmov_sample_coarse,
    keeping the autograd link to near / far that the reference has when n_importance == 0:
    dz_j/dnear = 1 - t_j, dz_j/dfar = t_j."""

    @staticmethod
    def forward(ctx, near, far, t_rand, n_samples):
        z = _ops.sample_coarse(near.detach(), far.detach(), t_rand, n_samples, n_samples)
        ctx.n = n_samples
        return z

    @staticmethod
    def backward(ctx, g_z):
        lin = torch.linspace(0.0, 1.0, ctx.n, device=g_z.device, dtype=g_z.dtype)
        g_far = g_z @ lin[:, None]
        g_near = g_z.sum(dim=1, keepdim=True) - g_far
        return g_near, g_far, None, None


class NeuSRenderer:
    def __init__(self, nerf, sdf_network, deviation_network, color_network, n_samples, n_importance, n_outside,
                 up_sample_steps, perturb):
        self.nerf = nerf
        self.sdf_network = sdf_network
        self.deviation_network = deviation_network
        self.color_network = color_network
        self.n_samples = n_samples
        self.n_importance = n_importance
        self.n_outside = n_outside
        self.up_sample_steps = up_sample_steps
        self.perturb = perturb

    # ------------------------------------------------------------------------------------------------
    def _check(self):
        if self.n_outside > 0:
            raise NotImplementedError("n_outside > 0 (NeRF++ background, render_core_outside) is out of scope: "
                                      "every shipped conf has n_outside = 0")
        if getattr(self.color_network, "mode", "idr") != "idr":
            raise NotImplementedError("RenderingNetwork mode must be 'idr' (every shipped conf)")
        if float(getattr(self.sdf_network, "scale", 1.0)) != 1.0:
            raise NotImplementedError("SDFNetwork.scale must be 1.0 (every shipped conf)")

    def sample_z(self, rays_o, rays_d, near, far, t_rand, qw=None):
        """z_vals [B, n_samples + n_importance] (models/renderer.py:385-446)."""
        if self.n_importance > 0:
            with torch.no_grad():
                if qw is None:
                    W, b = self.sdf_network.effective_weights()
                    qw = _packing.SdfQueryWeights(W, b)
                return _ops.hierarchical_sample(qw, rays_o.detach().float().contiguous(),
                                                rays_d.detach().float().contiguous(), near.detach(), far.detach(),
                                                t_rand, self.n_samples, self.n_importance, self.up_sample_steps)
        # n_importance == 0: z keeps its autograd link to near/far (renderer.py:389-390, 403-405)
        if getattr(self, "fused_coarse", False):          # opt-in (train.TrainStep(fused_rays=True)): one launch + 3 in backward
            return _CoarseZFn.apply(near, far, t_rand, self.n_samples)
        lin = torch.linspace(0.0, 1.0, self.n_samples, device=near.device)
        z = near + (far - near) * lin[None, :]
        if t_rand is not None:
            z = z + (t_rand - 0.5) * 2.0 / self.n_samples
        return z

    def render(self, rays_o, rays_d, near, far, perturb_overwrite=-1, background_rgb=None, cos_anneal_ratio=0.0,
               eval=False, t_rand=None, z_vals=None, eik_den=None, group=None, reduce_extra=None):
        """models/renderer.py:374-498.  Extras for micro-batched steps (train.TrainStep, `micro_batch`): `z_vals`
        injects samples drawn earlier by `sample_z`; `eik_den` (device scalar) is the whole-batch sum(relax) so that
        `gradient_error` is this call's numerator over the global normaliser (the calls' values then add up).
        Ray-sharded train steps pass `group` (torch.distributed process group): the eikonal numerator / normaliser are
        then summed over the ranks — a COLLECTIVE, issued only for calls that ask for it and run with autograd (never
        for eval / no-grad renders such as validate_image on one rank) — and `reduce_extra` (a small fp32 device
        tensor, e.g. [sum(mask), ray count]) is summed in place by the same all-reduce."""
        self._check()
        if not rays_o.is_cuda:
            raise RuntimeError("fmov_pose_b200 renders on CUDA tensors only (no CPU fallback)")
        batch_size = len(rays_o)
        sample_dist = 2.0 / self.n_samples
        perturb = self.perturb
        if perturb_overwrite >= 0:
            perturb = perturb_overwrite
        if perturb > 0:
            if t_rand is None:
                t_rand = torch.rand([batch_size, 1], device=rays_o.device)
        else:
            t_rand = None
        (W_s, b_s), (W_c, b_c) = _wn.effective_weights_fused([self.sdf_network, self.color_network])
        need_bwd = torch.is_grad_enabled() and not eval
        with torch.no_grad():      # one packing launch serves the sampling queries and the fine stage
            fw = _fine.FineWeights(W_s, b_s, W_c, b_c, need_backward=need_bwd)
        if z_vals is None:
            z_vals = self.sample_z(rays_o, rays_d, near, far, t_rand, qw=fw.query)
        n_samples = z_vals.shape[1]
        inv_s = torch.exp(self.deviation_network.variance * 10.0).clip(1e-6, 1e6)   # fields.py:294, renderer.py:290
        cfg = dict(sample_dist=sample_dist, cos_anneal_ratio=cos_anneal_ratio, background_rgb=background_rgb,
                   need_backward=need_bwd, group=group if need_bwd else None, fine_weights=fw, eik_den=eik_den,
                   reduce_extra=reduce_extra if need_bwd else None)
        if not need_bwd:
            with torch.no_grad():
                outs = _fine.RenderCoreFunction.apply(rays_o, rays_d, z_vals, inv_s, cfg, *W_s, *b_s, *W_c, *b_c)
        else:
            outs = _fine.RenderCoreFunction.apply(rays_o, rays_d, z_vals, inv_s, cfg, *W_s, *b_s, *W_c, *b_c)
        (color, weight_sum, weight_max, depth, weights, cdf, inside, mid_z, pts, sdf, gradients, grad_err) = outs
        if need_bwd:      # autograd link pts = o + d*mid_z (renderer.py:269-272); costs nothing unless pts is differentiated
            pts = _flow.PtsFunction.apply(rays_o, rays_d, z_vals, pts, sample_dist)
        s_val = (1.0 / inv_s).reshape(1, 1).expand(batch_size, 1)     # mean over samples of a constant
        return {
            "color_fine": color,
            "depth_fine": depth,
            "s_val": s_val,
            "cdf_fine": cdf,
            "weight_sum": weight_sum,
            "weight_max": weight_max,
            "gradients": gradients,
            "weights": weights,
            "gradient_error": grad_err,
            "inside_sphere": inside,
            "pts": pts,
            # extras (not in the reference dict)
            "z_vals": z_vals, "mid_z_vals": mid_z, "sdf": sdf,
        }

    # ------------------------------------------------------------------------------------------------
    def render_image(self, rays_o, rays_d, near_far_fn=None, chunk_rays=8192, cos_anneal_ratio=1.0,
                     background_rgb=None):
        """Forward-only render of a whole frame: what validate_image / render_novel_image /
        render_poses do with 600 sequential 512-ray render() calls under autograd (exp_runner.py:1444-1562,
        1874-1926), as a few large no-grad launches.  rays_o, rays_d: [..., 3] (e.g. [H,W,3] from gen_rays_at).
        Returns device tensors shaped like the input: color_fine [...,3], normals [...,3] = sum_j w_j * grad_j *
        inside_j (exp_runner.py:1494-1501, world frame), depth_fine [...,1], weight_sum [...,1]."""
        self._check()
        shape = rays_o.shape[:-1]
        ro = rays_o.reshape(-1, 3).float().contiguous()
        rd = rays_d.reshape(-1, 3).float().contiguous()
        N = ro.shape[0]
        dev = ro.device
        color = torch.empty(N, 3, device=dev)
        normals = torch.empty(N, 3, device=dev)
        depth = torch.empty(N, 1, device=dev)
        wsum = torch.empty(N, 1, device=dev)
        if near_far_fn is None:
            def near_far_fn(o, d):          # Dataset.near_far_from_sphere (models/dataset.py:835-842)
                a = torch.sum(d ** 2, dim=-1, keepdim=True)
                b = 2.0 * torch.sum(o * d, dim=-1, keepdim=True)
                mid = 0.5 * (-b) / a
                return mid - 1.0, mid + 1.0
        with torch.no_grad():
            for i in range(0, N, chunk_rays):
                o, d = ro[i:i + chunk_rays], rd[i:i + chunk_rays]
                near, far = near_far_fn(o, d)
                out = self.render(o, d, near, far, perturb_overwrite=0, background_rgb=background_rgb,
                                  cos_anneal_ratio=cos_anneal_ratio, eval=True)
                color[i:i + chunk_rays] = out["color_fine"]
                normals[i:i + chunk_rays] = (out["gradients"] * (out["weights"] * out["inside_sphere"])[..., None]).sum(1)
                depth[i:i + chunk_rays] = out["depth_fine"]
                wsum[i:i + chunk_rays] = out["weight_sum"]
        return {"color_fine": color.reshape(*shape, 3), "normals": normals.reshape(*shape, 3),
                "depth_fine": depth.reshape(*shape, 1), "weight_sum": wsum.reshape(*shape, 1)}

    # ------------------------------------------------------------------------------------------------
    def extract_fields(self, bound_min, bound_max, resolution, first=0, count=None, out=None, precise="act"):
        """u = -sdf on the res^3 grid (models/renderer.py:9-37 with query_func of :506) in one launch.
        `first`/`count` select a contiguous x-major range (grid partitioning across ranks).
        `precise`: "act" (default) = activation-split chain, 3.6e-4 of the reference's own grid on the whole +-1.01 box
        (north_star: SDF <= 1e-3) at 0.40 G queries/s; True = full split-precision chain, 1.3e-5 at 0.30 G/s; False = plain
        fp16 chain, 6.7e-4 at 0.74 G/s (profiles/r2b_grid_modes.txt)."""
        with torch.no_grad():
            W, b = self.sdf_network.effective_weights()
            qw = _packing.SdfQueryWeights(W, b, precise=precise is True)
            total = resolution ** 3
            if count is None:
                count = total - first
            if out is None:
                out = torch.empty(count, dtype=torch.float32, device=W[0].device)
            sc = float(self.sdf_network.scale)
            _ops.sdf_query_grid(qw, [float(v) for v in bound_min], [float(v) for v in bound_max], resolution, first,
                                count, out, in_scale=sc, out_scale=-1.0 / sc, precise=precise)
        return out

    def extract_geometry(self, bound_min, bound_max, resolution, threshold=0.0, precise="act"):
        """models/renderer.py:500-507: the 512^3 grid query and the marching cubes both run on the device; only the
        mesh crosses to the host (the reference copies 512 chunks of the grid to the host and runs PyMCubes there)."""
        u = self.extract_fields(bound_min, bound_max, resolution, precise=precise).reshape(resolution, resolution, resolution)
        return extract_geometry(bound_min, bound_max, resolution, threshold, u=u)

    def extract_color(self, vertices):
        """models/renderer.py:509-532: colour at mesh vertices with view dir = -normal."""
        pts = torch.as_tensor(vertices, dtype=torch.float32, device=self.deviation_network.variance.device)
        out = []
        with torch.no_grad():
            W, b = self.sdf_network.effective_weights()
            Wc, bc = self.color_network.effective_weights()
            fw = _fine.FineWeights(W, b, Wc, bc, need_backward=False)
            for chunk in pts.split(1 << 16):
                N = chunk.shape[0]
                st = _fine.Stash(N, chunk.device, with_backward=False)
                z0 = torch.zeros(N, 1, device=chunk.device)
                # pass 1: normals at the vertices (points as degenerate rays o = x, d = 0)
                _, nrm, _, _ = _fine.fine_forward(fw, st, chunk.contiguous(), torch.zeros_like(chunk), z0, 0.0)
                # pass 2: view direction = -normal; keep the point fixed by sampling at mid_z = 0
                #         (z = 0 and sample_dist = 0  =>  pts = o + d*0)
                _, _, rgb, _ = _fine.fine_forward(fw, st, chunk.contiguous(), (-nrm).contiguous(), z0, 0.0)
                out.append(rgb.cpu().numpy())
        return np.concatenate(out, axis=0)
