"""One train iteration of the reference's Runner.train (exp_runner.py:497-599, 772-816) on the B200
kernels, optionally ray-sharded across ranks: every rank renders its own ray shard with
replicated weights; the three whole-batch normalisers (sum(mask), sum(relax), ray count) and the
parameter gradients are all-reduced so that N-GPU results equal the single-GPU step on the union batch."""
import torch
import torch.nn.functional as F

from .models import camera as _camera


class TrainStep:
    def __init__(self, scene, igr_weight=0.1, mask_weight=5.0, lr=5e-4, pose_lr=5e-4, group=None, optimizer=True):
        self.s = scene
        self.igr_weight, self.mask_weight = igr_weight, mask_weight
        self.group = group
        self.world = torch.distributed.get_world_size(group) if group is not None else 1
        scene["renderer"].process_group = group
        nets = [scene["sdf_network"], scene["deviation_network"], scene["color_network"]]
        self.params = [p for n in nets for p in n.parameters() if p.requires_grad]
        if scene["pose_network"] is not None:
            self.pose_params = [p for p in scene["pose_network"].parameters() if p.requires_grad]
        else:
            self.pose_params = []
        self.all_params = self.params + self.pose_params
        self.optimizer = None
        if optimizer:       # exp_runner.py:264-269 (nets) and :258-262 (pose MLPs)
            groups = [dict(params=self.params, lr=lr)]
            if self.pose_params:
                groups.append(dict(params=self.pose_params, lr=pose_lr))
            self.optimizer = torch.optim.Adam(groups, fused=True)

    def pose_of(self, img_id):
        s = self.s
        if s["pose_network"] is not None:
            return s["pose_network"](img_id)[:3]
        sdf = s["sdf_network"]
        return _camera.barf_pose(sdf.se3_refine.weight[int(img_id)], sdf.noise_poses[int(img_id), :3, :])

    def losses(self, out, true_rgb, mask):
        """exp_runner.py:562-599, 772-779 with global normalisers when ray-sharded."""
        if self.mask_weight > 0.0:
            mask = (mask > 0.5).float()
        else:
            mask = torch.ones_like(mask)
        msum = mask.sum()
        n_rays = torch.tensor(float(mask.shape[0]), device=mask.device)
        if self.group is not None:
            pack = torch.stack([msum, n_rays])
            torch.distributed.all_reduce(pack, group=self.group)
            msum, n_rays = pack[0], pack[1]
        mask_sum = msum + 1e-5
        color_error = (out["color_fine"] - true_rgb) * mask
        color_loss = color_error.abs().sum() / mask_sum
        eik = out["gradient_error"]                       # already globally normalised by the renderer
        bce = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask, reduction="sum") / n_rays
        loss = color_loss + eik * self.igr_weight + bce * self.mask_weight
        return dict(loss=loss, color_loss=color_loss, eikonal_loss=eik, mask_loss=bce)

    def forward_backward(self, img_id, batch_size, pixels=None, t_rand=None, cos_anneal_ratio=1.0):
        s = self.s
        ds, rend = s["dataset"], s["renderer"]
        pose = self.pose_of(img_id)
        data, _ = ds.gen_random_rays_at(img_id, batch_size, pose, pixels=pixels)
        rays_o, rays_d, true_rgb, mask = data[:, :3], data[:, 3:6], data[:, 6:9], data[:, 9:10]
        near, far = ds.near_far_from_sphere(rays_o, rays_d)
        out = rend.render(rays_o, rays_d, near, far, cos_anneal_ratio=cos_anneal_ratio, t_rand=t_rand)
        ls = self.losses(out, true_rgb, mask)
        for p in self.all_params:
            p.grad = None
        ls["loss"].backward()
        if self.group is not None:
            self.allreduce_grads()
        return ls, out

    def allreduce_grads(self):
        """one SUM all-reduce of all MLP + pose gradients (~3.2 MB fp32) over NCCL/NVLink"""
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.all_params]
        flat = torch._utils._flatten_dense_tensors(grads)
        torch.distributed.all_reduce(flat, group=self.group)
        for p, g in zip(self.all_params, torch._utils._unflatten_dense_tensors(flat, grads)):
            p.grad = g

    def step(self, img_id, batch_size, pixels=None, t_rand=None, cos_anneal_ratio=1.0):
        ls, out = self.forward_backward(img_id, batch_size, pixels, t_rand, cos_anneal_ratio)
        if self.optimizer is not None:
            self.optimizer.step()
        return ls, out
