"""One train iteration of the reference's Runner.train (exp_runner.py:497-599, 772-816) on the B200
kernels, optionally ray-sharded across ranks: every rank renders its own ray shard with
replicated weights; the three whole-batch normalisers (sum(mask), sum(relax), ray count) and the
parameter gradients are all-reduced so that N-GPU results equal the single-GPU step on the union batch."""
import math

import torch
import torch.nn.functional as F

from . import _lib as _L
from . import flow as _flow
from . import ops as _ops
from .models import camera as _camera


class _FusedLossFn(torch.autograd.Function):
    """colour + mask loss terms of exp_runner.py:562-599 in one launch (fmov_loss_fwd_bwd computes the per-ray terms and
    their gradients together): (color [B,3], weight_sum [B,1]) -> (color_loss, mask_loss) scalars."""

    @staticmethod
    def forward(ctx, color, weight_sum, true_rgb, mask, mask_sum, n_rays_global, mask_weight):
        partial, g_color, g_wsum = _ops.loss_fwd_bwd(_L.f32c(color.detach()), _L.f32c(weight_sum.detach()),
                                                     _L.f32c(true_rgb), _L.f32c(mask), _L.f32c(mask_sum.reshape(1)),
                                                     int(n_rays_global), float(mask_weight))
        ctx.save_for_backward(g_color, g_wsum)
        ctx.mask_weight = float(mask_weight)
        sums = partial.sum(0)
        return sums[0], sums[1]

    @staticmethod
    def backward(ctx, g_col, g_bce):
        g_color, g_wsum = ctx.saved_tensors
        gc = None if g_col is None else g_color * g_col
        # the kernel's weight_sum gradient already carries mask_weight (loss = colour + mask_weight * bce)
        gw = None if (g_bce is None or ctx.mask_weight == 0.0) else g_wsum * (g_bce / ctx.mask_weight)
        return gc, gw, None, None, None, None, None


class LRSchedule:
    """Runner.update_learning_rate (exp_runner.py:1049-1087) as a pure host function: linear warm-up then cosine decay to
    `learning_rate_alpha` for the networks; for pose_type "seg" each pose MLP follows its own cosine over its own
    progress counter (`step / max_pro_iteration` with floor `pose_alpha`, or the global `step / end_iter` with floor
    `learning_rate_alpha` for "_wo_global_conf" experiments)."""

    def __init__(self, learning_rate=5e-4, learning_rate_alpha=0.05, warm_up_end=5000, end_iter=300000, pose_lr=5e-4,
                 pose_alpha=0.5, max_pro_iteration=1000, wo_global_conf=False):
        self.learning_rate, self.learning_rate_alpha = learning_rate, learning_rate_alpha
        self.warm_up_end, self.end_iter = warm_up_end, end_iter
        self.pose_lr, self.pose_alpha, self.max_pro_iteration = pose_lr, pose_alpha, max_pro_iteration
        self.wo_global_conf = wo_global_conf

    @staticmethod
    def _cosine(progress, alpha):
        return (math.cos(math.pi * progress) + 1.0) * 0.5 * (1 - alpha) + alpha

    def net_lr(self, iter_step):
        if iter_step < self.warm_up_end:
            return self.learning_rate * (iter_step / self.warm_up_end)
        return self.learning_rate * self._cosine((iter_step - self.warm_up_end) / (self.end_iter - self.warm_up_end),
                                                 self.learning_rate_alpha)

    def pose_mlp_lr(self, step):
        """`step` = the pose MLP's own progress counter after this iteration's increment (SegLearnPose.step_progress)"""
        if self.wo_global_conf:
            return self.pose_lr * self._cosine(step / self.end_iter, self.learning_rate_alpha)
        return self.pose_lr * self._cosine(step / self.max_pro_iteration, self.pose_alpha)


class TrainStep:
    def __init__(self, scene, igr_weight=0.1, mask_weight=5.0, lr=5e-4, pose_lr=5e-4, group=None, optimizer=True,
                 capturable=False, flow_weight=0.0, unit_sphere_weight=0.0, maintain_shape=False,
                 detach_flow_on_sdf=False, detach_ref=False, fused_loss=True, fused_rays=None):
        """`optimizer`: True = this library's FlatAdam (gradient gather + Adam in two launches, only the parameter groups
        of the rendered frames are stepped — also across ranks), "torch" = torch.optim.Adam(fused=True) with one param
        group per pose MLP, False = none (the caller steps).  `fused_loss`: colour + mask terms through
        fmov_loss_fwd_bwd (one launch instead of ~40 torch ops).  `fused_rays`: near / far from the ray-generation kernel
        and, for n_importance == 0, the coarse z through fmov_sample_coarse with its closed-form backward (~40 torch
        launches less); None = on when the dataset is this package's RayDataset (the call needs its `with_near_far`)."""
        from .models.dataset import RayDataset
        self.s = scene
        self.fused_loss = bool(fused_loss)
        if fused_rays is None:
            fused_rays = isinstance(scene.get("dataset"), RayDataset)
        self.fused_rays = bool(fused_rays)
        scene["renderer"].fused_coarse = self.fused_rays
        self.igr_weight, self.mask_weight = igr_weight, mask_weight
        # exp_runner.py:150-160, 327-336 (train.flow_weight, unit_sphere_weight, maintain_shape, detach_*)
        self.flow_weight, self.unit_sphere_weight = float(flow_weight), float(unit_sphere_weight)
        self.maintain_shape, self.detach_flow_on_sdf, self.detach_ref = maintain_shape, detach_flow_on_sdf, detach_ref
        self.group = group
        self.background_rgb = None       # torch.ones([1,3]) when use_white_bkgd (exp_runner.py:556)
        self.world = torch.distributed.get_world_size(group) if group is not None else 1
        nets = [scene["sdf_network"], scene["deviation_network"], scene["color_network"]]
        self.params = [p for n in nets for p in n.parameters() if p.requires_grad]
        if scene["pose_network"] is not None:
            self.pose_params = [p for p in scene["pose_network"].parameters() if p.requires_grad]
        else:
            self.pose_params = []
        self.all_params = self.params + self.pose_params
        self.optimizer = None
        self.own_optimizer = False
        self.pose_group_of = {}
        if optimizer:       # exp_runner.py:264-269 (nets) and :258-262 (pose MLPs)
            dev = self.params[0].device
            own = optimizer != "torch"
            if capturable and not own:        # learning rates live on the device so a captured step follows the LR schedule
                lr = torch.tensor(float(lr), device=dev)
                pose_lr = torch.tensor(float(pose_lr), device=dev)
            # one optimiser over [networks | pose MLP 0 | pose MLP 1 | ...]: the reference keeps one Adam per pose MLP
            # (exp_runner.py:258-262, 812-816) and steps only those of the rendered frames
            groups = [dict(params=self.params, lr=lr)]
            mlps = getattr(scene["pose_network"], "pose_mlps", None)
            if mlps is not None:
                for k, mlp in enumerate(mlps):
                    ps = [p for p in mlp.parameters() if p.requires_grad]
                    if ps:
                        self.pose_group_of[k] = len(groups)
                        groups.append(dict(params=ps, lr=pose_lr.clone() if torch.is_tensor(pose_lr) else pose_lr))
            elif self.pose_params:
                groups.append(dict(params=self.pose_params, lr=pose_lr))
            if own:
                from .optim import FlatAdam
                self.optimizer = FlatAdam(groups, process_group=group)
                self.own_optimizer = True
                capturable = True
            else:
                # torch keeps its step count per parameter and skips parameters without a gradient, so param groups give
                # the reference's per-pose-MLP updates on one GPU
                self.optimizer = torch.optim.Adam(groups, fused=True, capturable=capturable)
                if capturable:        # state is created up front: lazy creation inside a capture would be replayed
                    for grp in self.optimizer.param_groups:
                        for p in grp["params"]:
                            self.optimizer.state[p] = dict(step=torch.zeros((), dtype=torch.float32, device=p.device),
                                                           exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
        self.capturable = capturable

    def active_groups(self, *img_ids):
        """optimiser groups a step over these frames updates: the networks + the pose MLP of every rendered frame
        (pose_mlp_index_set, exp_runner.py:785-791); gf / se3 poses live in one always-active group"""
        act = {0}
        pn = self.s["pose_network"]
        seg = getattr(pn, "segment_img_num", None)
        if self.pose_group_of and seg:
            for i in img_ids:
                if i is not None and int(i) // seg in self.pose_group_of:
                    act.add(self.pose_group_of[int(i) // seg])
        elif self.pose_params:
            act.add(1)
        return sorted(act)

    @staticmethod
    def _write_lr(grp, v):
        if torch.is_tensor(grp["lr"]):
            grp["lr"].fill_(float(v))
        else:
            grp["lr"] = float(v)

    def set_lr(self, lr, pose_lr=None, pose_mlp_lrs=None):
        """exp_runner.py:1049-1087 writes param_group['lr'] every iteration; device-side when capturable.
        `pose_lr` applies to every pose group, `pose_mlp_lrs` = {pose MLP index: lr} to single pose MLPs."""
        groups = self.optimizer.param_groups
        self._write_lr(groups[0], lr)
        for grp in groups[1:]:
            self._write_lr(grp, lr if pose_lr is None else pose_lr)
        for k, v in (pose_mlp_lrs or {}).items():
            if k in getattr(self, "pose_group_of", {}):
                self._write_lr(groups[self.pose_group_of[k]], v)

    def update_learning_rate(self, schedule, iter_step, pose_mlp_index_set=None):
        """Runner.update_learning_rate (exp_runner.py:1049-1087): networks from the global iteration; for SegLearnPose
        every pose MLP in `pose_mlp_index_set` advances its progress counter and gets its own rate."""
        groups = self.optimizer.param_groups
        self._write_lr(groups[0], schedule.net_lr(iter_step))
        pn = self.s["pose_network"]
        if hasattr(pn, "pose_mlps"):
            for k in (pose_mlp_index_set or ()):
                step = float(pn.step_progress(k))
                if k in self.pose_group_of:
                    self._write_lr(groups[self.pose_group_of[k]], schedule.pose_mlp_lr(step))
        else:        # pose_type gf / se3: the reference has no separate pose schedule (only self.optimizer's groups)
            for grp in groups[1:]:
                self._write_lr(grp, schedule.net_lr(iter_step))

    def pose_of(self, img_id, img_t=None):
        s = self.s
        if s["pose_network"] is not None:
            return s["pose_network"](img_id, img_t)[:3] if img_t is not None else s["pose_network"](img_id)[:3]
        sdf = s["sdf_network"]
        if img_t is not None:
            it = img_t.reshape(1)
            return _camera.barf_pose(sdf.se3_refine.weight.index_select(0, it)[0],
                                     sdf.noise_poses.index_select(0, it)[0, :3, :])
        return _camera.barf_pose(sdf.se3_refine.weight[int(img_id)], sdf.noise_poses[int(img_id), :3, :])

    def graph_key(self, img_id):
        """frames that share trainable pose parameters share one captured graph"""
        pn = self.s["pose_network"]
        seg = getattr(pn, "segment_img_num", None)
        return int(img_id) // seg if seg else 0

    def mask_stats(self, mask):
        """-> (mask in {0,1}, stats = [sum(mask), ray count] on this rank).  Under ray sharding `stats` rides on the
        eikonal normaliser's all-reduce inside render() (one collective for all three whole-batch normalisers)."""
        if self.mask_weight > 0.0:
            mask = (mask > 0.5).float()
        else:
            mask = torch.ones_like(mask)
        stats = torch.stack([mask.sum(), torch.full((), float(mask.shape[0]), device=mask.device)])
        return mask, stats

    def losses(self, out, true_rgb, mask, stats=None):
        """exp_runner.py:562-599, 772-779 with global normalisers when ray-sharded.  `stats`: what mask_stats returned,
        already summed over the ranks by render(); None = compute (and all-reduce) here."""
        if stats is None:
            mask, stats = self.mask_stats(mask)
            if self.group is not None:
                torch.distributed.all_reduce(stats, group=self.group)
        msum, n_rays = stats[0], stats[1]
        mask_sum = msum + 1e-5
        if self.fused_loss:
            # every rank holds the same number of rays (fixed split), so the global ray count is known on the host
            color_loss, bce = _FusedLossFn.apply(out["color_fine"], out["weight_sum"], true_rgb, mask, mask_sum,
                                                 mask.shape[0] * self.world, self.mask_weight)
            eik = out["gradient_error"]
            loss = color_loss + eik * self.igr_weight + bce * self.mask_weight
            return dict(loss=loss, color_loss=color_loss, eikonal_loss=eik, mask_loss=bce)
        color_error = (out["color_fine"] - true_rgb) * mask
        color_loss = color_error.abs().sum() / mask_sum
        eik = out["gradient_error"]                       # already globally normalised by the renderer
        bce = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask, reduction="sum") / n_rays
        loss = color_loss + eik * self.igr_weight + bce * self.mask_weight
        return dict(loss=loss, color_loss=color_loss, eikonal_loss=eik, mask_loss=bce)

    def forward_backward(self, img_id, batch_size, pixels=None, t_rand=None, cos_anneal_ratio=1.0, img_t=None,
                         micro_batch=None, additional_img_id=None, add_pixels=None, add_img_t=None, reduce=True):
        """One ordinary iteration.  `additional_img_id` (with `maintain_shape`, exp_runner.py:512-548): `batch_size` more
        rays of a second, already registered frame through ITS pose are appended, so the render sees 2*batch_size rays
        (t_rand then has 2*batch_size rows) and two pose MLPs receive gradients."""
        if micro_batch is not None and batch_size > micro_batch:
            assert additional_img_id is None, "micro-batching takes one frame per step"
            return self.forward_backward_chunked(img_id, batch_size, micro_batch, pixels, t_rand, cos_anneal_ratio, img_t,
                                                 reduce=reduce)
        s = self.s
        ds, rend = s["dataset"], s["renderer"]
        pose = self.pose_of(img_id, img_t)
        nf = self.fused_rays
        nf_kw = dict(with_near_far=True) if nf else {}          # default call unchanged (any dataset with the reference's API)
        r = ds.gen_random_rays_at(img_id, batch_size, pose, pixels=pixels, img_idx_t=img_t, **nf_kw)
        data, near, far = r[0], (r[2] if nf else None), (r[3] if nf else None)
        if additional_img_id is not None:
            add_pose = self.pose_of(additional_img_id, add_img_t)
            r = ds.gen_random_rays_at(additional_img_id, batch_size, add_pose, pixels=add_pixels, img_idx_t=add_img_t,
                                      **nf_kw)
            data = torch.cat([data, r[0]], dim=0)
            if nf:
                near, far = torch.cat([near, r[2]], dim=0), torch.cat([far, r[3]], dim=0)
        rays_o, rays_d, true_rgb, mask = data[:, :3], data[:, 3:6], data[:, 6:9], data[:, 9:10]
        if not nf:
            near, far = ds.near_far_from_sphere(rays_o, rays_d)
        mask, stats = self.mask_stats(mask)
        out = rend.render(rays_o, rays_d, near, far, cos_anneal_ratio=cos_anneal_ratio, t_rand=t_rand,
                          background_rgb=self.background_rgb, group=self.group, reduce_extra=stats)
        ls = self.losses(out, true_rgb, mask, stats)
        if self.unit_sphere_weight > 0:      # exp_runner.py:714-724
            ls["unit_sphere_loss"] = _flow.unit_sphere_loss(out, rays_o, rays_d, 2.0 / rend.n_samples,
                                                            self.unit_sphere_weight, group=self.group)
            ls["loss"] = ls["loss"] + ls["unit_sphere_loss"]
        for p in self.all_params:
            p.grad = None
        ls["loss"].backward()
        if self.group is not None and reduce:
            self.allreduce_grads()
        return ls, out

    def forward_backward_chunked(self, img_id, batch_size, micro_batch, pixels=None, t_rand=None, cos_anneal_ratio=1.0,
                                 img_t=None, reduce=True):
        """The same iteration for ray batches whose activation stash does not fit HBM at once (27 GB per 8192 rays at
        64+64: BASELINE config C3 puts 32 K / 16 K rays on a GPU at N = 2 / 4).  Rays and the no-grad hierarchical
        samples are produced for the whole batch; the fine stage + losses + backward then run per micro-batch against
        the WHOLE-batch normalisers (sum(mask), ray count, and sum(relax), which depends only on the sample positions
        and is therefore known before the fine stage), gradients accumulate, and the pose / ray-generation graph is
        back-propagated once at the end.  Result == the one-shot step up to fp32 summation order."""
        s = self.s
        ds, rend = s["dataset"], s["renderer"]
        dev = self.params[0].device
        pose = self.pose_of(img_id, img_t)
        nf_k = None
        if self.fused_rays:      # the one-shot step samples from the ray-generation kernel's near / far: so does this one
            r = ds.gen_random_rays_at(img_id, batch_size, pose, pixels=pixels, img_idx_t=img_t, with_near_far=True)
            data, nf_k = r[0], (r[2].detach(), r[3].detach())
        else:
            data, _ = ds.gen_random_rays_at(img_id, batch_size, pose, pixels=pixels, img_idx_t=img_t)
        rays_o, rays_d, true_rgb, mask = data[:, :3], data[:, 3:6], data[:, 6:9], data[:, 9:10]
        true_rgb, mask = true_rgb.detach(), mask.detach()      # slices of the same cat() as the rays
        ro = rays_o.detach().contiguous().requires_grad_(rays_o.requires_grad)      # cut: per-chunk backward stops here
        rd = rays_d.detach().contiguous().requires_grad_(rays_d.requires_grad)
        if t_rand is None and rend.perturb > 0:
            t_rand = torch.rand([batch_size, 1], device=dev)
        sd = 2.0 / rend.n_samples
        with torch.no_grad():
            near, far = nf_k if nf_k is not None else ds.near_far_from_sphere(ro, rd)
            z_all = rend.sample_z(ro, rd, near, far, t_rand)
            dists = torch.cat([z_all[:, 1:] - z_all[:, :-1], torch.full_like(z_all[:, :1], sd)], dim=-1)
            pts = ro[:, None, :] + rd[:, None, :] * (z_all + 0.5 * dists)[:, :, None]
            relax_sum = (torch.linalg.norm(pts, ord=2, dim=-1) < 1.2).sum().float()        # renderer.py:343-346
            del pts, dists
            m = (mask > 0.5).float() if self.mask_weight > 0.0 else torch.ones_like(mask)
            pack = torch.stack([m.sum(), torch.full((), float(batch_size), device=dev), relax_sum])
            if self.group is not None:
                torch.distributed.all_reduce(pack, group=self.group)
            mask_sum, n_rays, eik_den = pack[0] + 1e-5, pack[1], pack[2]
        for p in self.all_params:
            p.grad = None
        tot = None
        keep = {}
        for i in range(0, batch_size, micro_batch):
            sl = slice(i, min(i + micro_batch, batch_size))
            o_c, d_c = ro[sl], rd[sl]
            nf = ds.near_far_from_sphere(o_c, d_c)
            tr_c = None if t_rand is None else t_rand[sl]
            # n_importance == 0: z keeps its link to near/far (renderer.py:390), so it is rebuilt inside render()
            out = rend.render(o_c, d_c, nf[0], nf[1], cos_anneal_ratio=cos_anneal_ratio, t_rand=tr_c,
                              background_rgb=self.background_rgb, z_vals=z_all[sl] if rend.n_importance > 0 else None,
                              eik_den=eik_den)
            col = ((out["color_fine"] - true_rgb[sl]) * m[sl]).abs().sum() / mask_sum
            bce = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), m[sl], reduction="sum") / n_rays
            eik = out["gradient_error"]
            loss = col + eik * self.igr_weight + bce * self.mask_weight
            loss.backward()
            part = torch.stack([loss.detach(), col.detach(), eik.detach(), bce.detach()])
            tot = part if tot is None else tot + part
            for k in ("color_fine", "weight_sum", "depth_fine", "weight_max"):
                keep.setdefault(k, []).append(out[k].detach())
            del out, loss
        roots, grads = [], []
        for t_, g_ in ((rays_o, ro.grad), (rays_d, rd.grad)):
            if t_.requires_grad and g_ is not None:
                roots.append(t_)
                grads.append(g_)
        if roots:
            torch.autograd.backward(roots, grads)
        if self.group is not None and reduce:
            self.allreduce_grads()
        ls = dict(loss=tot[0], color_loss=tot[1], eikonal_loss=tot[2], mask_loss=tot[3])
        return ls, {k: torch.cat(v, 0) for k, v in keep.items()}

    def flow_forward_backward(self, img_id_corr, batch_size, current_img_num, interval=1, additional_img_id=None,
                              img_id=None, indexs=None, add_pixels=None, t_rand=None, cos_anneal_ratio=1.0, reduce=True):
        """A `use_flow` iteration (exp_runner.py:441-456, 512-548, 562-599, 605-688, 714-724): batch_size//2 matched
        pixel pairs between frame img_id_corr and a matched frame (+ batch_size rays of `additional_img_id` with
        maintain_shape), one render of the concatenated rays, colour/eikonal/mask losses plus the reprojection loss
        of every sample point into the other frame of the pair.  Returns (losses, render dict, img_id) or None when the
        frame has no usable match (the reference then falls back to an ordinary iteration)."""
        s = self.s
        ds, rend, pn = s["dataset"], s["renderer"], s["pose_network"]
        if pn is None:
            raise NotImplementedError("flow iterations need a pose network (pose_type gf / seg), as the reference's "
                                      "gen_random_ray_pairs_at (models/dataset.py:728, 749)")
        data, pixels_xy, pixels_xy_corr, img_id, _ = ds.gen_random_ray_pairs_at(
            img_id_corr, batch_size // 2, pn, current_img_num, interval, img_id=img_id, indexs=indexs)
        if data is None:
            return None
        if self.maintain_shape:
            add_pose = pn(additional_img_id)[:3]
            add_data, _ = ds.gen_random_rays_at(additional_img_id, batch_size, add_pose, pixels=add_pixels)
            data = torch.cat([data, add_data], dim=0)
        rays_o, rays_d, true_rgb, mask = data[:, :3], data[:, 3:6], data[:, 6:9], data[:, 9:10]
        near, far = ds.near_far_from_sphere(rays_o, rays_d)
        mask, stats = self.mask_stats(mask)
        out = rend.render(rays_o, rays_d, near, far, cos_anneal_ratio=cos_anneal_ratio, t_rand=t_rand,
                          background_rgb=self.background_rgb, group=self.group, reduce_extra=stats)
        ls = self.losses(out, true_rgb, mask, stats)
        sd = 2.0 / rend.n_samples
        n_terms = None if self.group is None else 2 * (data.shape[0] // (4 if self.maintain_shape else 2)) * self.world
        # the pose modules are evaluated again for the projection, as exp_runner.py:628-631, 660-663
        fl = _flow.flow_loss(out, rays_o, rays_d, pn(int(img_id))[:3], pn(int(img_id_corr))[:3],
                             ds.intrinsics_all[int(img_id)], ds.intrinsics_all[int(img_id_corr)], pixels_xy,
                             pixels_xy_corr, sd, self.flow_weight, maintain_shape=self.maintain_shape,
                             detach_flow_on_sdf=self.detach_flow_on_sdf, detach_ref=self.detach_ref, n_terms=n_terms)
        ls["flow_loss"] = fl
        ls["loss"] = ls["loss"] + fl
        if self.unit_sphere_weight > 0:
            ls["unit_sphere_loss"] = _flow.unit_sphere_loss(out, rays_o, rays_d, sd, self.unit_sphere_weight,
                                                            group=self.group)
            ls["loss"] = ls["loss"] + ls["unit_sphere_loss"]
        for p in self.all_params:
            p.grad = None
        ls["loss"].backward()
        if self.group is not None and reduce:
            self.allreduce_grads()
        return ls, out, img_id

    def step_flow(self, img_id_corr, batch_size, current_img_num, interval=1, additional_img_id=None, **kw):
        r = self.flow_forward_backward(img_id_corr, batch_size, current_img_num, interval, additional_img_id,
                                       reduce=not self.own_optimizer, **kw)
        if r is not None and self.optimizer is not None:
            if self.own_optimizer:
                self.optimizer.step(self.active_groups(img_id_corr, r[2], additional_img_id))
            else:
                self.optimizer.step()
        return r

    def allreduce_grads(self):
        """one SUM all-reduce of all MLP + pose gradients (~3.2 MB fp32) over NCCL/NVLink, for callers that step the
        parameters themselves (`optimizer=False` / "torch"; FlatAdam reduces its own flat buffer).  A per-parameter
        has-gradient flag rides along: parameters that NO rank produced a gradient for (pose MLPs of frames nobody
        rendered) keep `grad = None`, so an optimiser skips them exactly as on one GPU (exp_runner.py:785-816).  Reading
        the flags is a host sync; inside a CUDA-graph capture it is skipped and such gradients stay zero tensors."""
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.all_params]
        flags = torch.tensor([0.0 if p.grad is None else 1.0 for p in self.all_params], device=grads[0].device)
        flat = torch._utils._flatten_dense_tensors(grads + [flags])
        torch.distributed.all_reduce(flat, group=self.group)
        parts = torch._utils._unflatten_dense_tensors(flat, grads + [flags])
        capturing = grads[0].is_cuda and torch.cuda.is_current_stream_capturing()
        have = None if capturing else (parts[-1] > 0).tolist()
        for i, (p, g) in enumerate(zip(self.all_params, parts[:-1])):
            p.grad = g if (have is None or have[i]) else None

    def step(self, img_id, batch_size, pixels=None, t_rand=None, cos_anneal_ratio=1.0, img_t=None, micro_batch=None,
             additional_img_id=None, add_pixels=None, add_img_t=None):
        ls, out = self.forward_backward(img_id, batch_size, pixels, t_rand, cos_anneal_ratio, img_t, micro_batch,
                                        additional_img_id, add_pixels, add_img_t, reduce=not self.own_optimizer)
        if self.optimizer is not None:
            if self.own_optimizer:      # gathers the local gradients, all-reduces them when sharded, steps the union
                self.optimizer.step(self.active_groups(img_id, additional_img_id))
            else:
                self.optimizer.step()
        return ls, out


class GraphedTrainStep:
    """TrainStep.step captured once per pose-parameter set as a CUDA graph and replayed.

    A step is ~220 launches of this library plus the torch glue (weight-norm, pose MLP head, losses, fused Adam, NCCL
    all-reduce): eager execution is host-bound for the shipped 512/1024-ray batches and leaves launch gaps at 8 K rays.
    The captured step has static shapes: `batch_size` rays, pixel indices / jitter / frame index are device buffers that
    `step()` refills before each replay (from pinned host memory or device tensors), the learning rates are device
    scalars (`TrainStep.set_lr`), the activation stash is the recycled Stash pool buffer.  Nothing in the step reads
    back to the host, so one replay == one reference iteration (exp_runner.py:497-599, 772-816)."""

    def __init__(self, ts, batch_size, cos_anneal_ratio=1.0, two_frames=False, micro_batch=None):
        """`two_frames`: the maintain_shape iteration of the shipped confs (exp_runner.py:512-548): `batch_size` rays
        of the current frame + `batch_size` rays of an earlier frame in one render; `step()` then also takes the second
        frame's index and pixels, and t_rand has 2*batch_size rows.  `micro_batch`: ray batches whose activation stash
        does not fit HBM at once (config C3: 32 K / 16 K rays per GPU at N = 2 / 4) are captured as ONE graph of
        TrainStep's micro-batched step (whole-batch sampling and normalisers, fine stage + backward per micro-batch)."""
        assert ts.optimizer is None or ts.capturable, "build the TrainStep with capturable=True (or the default FlatAdam)"
        self.ts, self.B, self.car = ts, int(batch_size), float(cos_anneal_ratio)
        self.two = bool(two_frames)
        self.micro_batch = micro_batch
        assert not (self.two and micro_batch), "micro-batching takes one frame per step"
        dev = ts.params[0].device
        self.dev = dev
        self.px = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.py = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.tr = torch.zeros(self.B * (2 if self.two else 1), 1, dtype=torch.float32, device=dev)
        self.img = torch.zeros(1, dtype=torch.int64, device=dev)
        self.px2 = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.py2 = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.img2 = torch.zeros(1, dtype=torch.int64, device=dev)
        self.graphs = {}
        self.pool = None
        self.launches_per_step = 0
        self.kernels_per_step = 0
        self._keep = []

    def set_cos_anneal_ratio(self, ratio):
        """cos_anneal_ratio is a launch constant of the compositing kernels, so it is part of the graph key: constant
        for the shipped confs (anneal_end = 0 -> 1.0, exp_runner.py:1043-1047); while it ramps, every new value costs one
        capture (quantise it on the caller side if anneal_end > 0)."""
        self.car = float(ratio)

    def _kw(self, add_img_id):
        kw = dict(pixels=(self.px, self.py), t_rand=self.tr, cos_anneal_ratio=self.car, img_t=self.img)
        if self.micro_batch:
            kw.update(micro_batch=self.micro_batch)
        if self.two:
            kw.update(additional_img_id=add_img_id, add_pixels=(self.px2, self.py2), add_img_t=self.img2)
        return kw

    def _body(self, img_id, add_img_id=None):
        return self.ts.step(img_id, self.B, **self._kw(add_img_id))

    def _capture(self, key, img_id, add_img_id=None):
        from . import _lib as L
        from .fine import Stash
        ts = self.ts
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            # eager pass without the optimizer: one-time library/cuBLAS initialisation and the stash allocation
            ts.forward_backward(img_id, self.B, **self._kw(add_img_id))
        cur.wait_stream(side)
        torch.cuda.synchronize()
        for p in ts.all_params:
            p.grad = None
        self._keep = [b for lst in Stash._pool.values() for b in lst]      # keep the stash buffer alive with the graph
        if ts.own_optimizer and ts.optimizer.n_groups > 64:      # the active set's flag vector must pre-exist (H2D copy)
            ts.optimizer.flags_for(ts.active_groups(img_id, add_img_id))
        g = torch.cuda.CUDAGraph()
        n0, k0 = L.n_calls, L.kernel_launches()
        # thread_local: other threads (NCCL watchdog, autograd workers) may keep calling the CUDA runtime during capture
        with torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
            ls, out = self._body(img_id, add_img_id)
        self.launches_per_step = L.n_calls - n0                  # C-ABI calls recorded into the graph
        self.kernels_per_step = L.kernel_launches() - k0         # kernels of this library that one replay executes
        if self.pool is None:
            self.pool = g.pool()
        self.graphs[key] = (g, ls, out)

    def step(self, img_id, px, py, t_rand, add_img_id=None, add_px=None, add_py=None):
        """px, py int64 [B], t_rand fp32 [B,1] ([2B,1] with two_frames) (device or pinned host tensors) -> (losses,
        render dict): static tensors that the next replay overwrites."""
        self.px.copy_(px, non_blocking=True)
        self.py.copy_(py, non_blocking=True)
        self.tr.copy_(t_rand.reshape(self.tr.shape), non_blocking=True)
        self.img.fill_(int(img_id))
        key = (self.ts.graph_key(img_id), self.car)
        if self.two:
            self.px2.copy_(add_px, non_blocking=True)
            self.py2.copy_(add_py, non_blocking=True)
            self.img2.fill_(int(add_img_id))
            key = key + (self.ts.graph_key(add_img_id),)
        if key not in self.graphs:
            self._capture(key, img_id, add_img_id)
        g, ls, out = self.graphs[key]
        g.replay()
        return ls, out
