"""Ray generation side of models/dataset.py — the math of `gen_random_rays_at` (dataset.py:634-681),
`gen_rays_at` (:547-576) and `near_far_from_sphere` (:835-842) on the fused pose/ray-gen kernel.

Loading images / masks / cameras from disk (Dataset.__init__, dataset.py:146-545) is out of scope
(SURVEY.md §2 row 16): `RayDataset` is constructed from tensors that are already in memory (the
benchmark uses synthetic frames) and exposes the attributes the train loop reads."""
import os

import numpy as np
import torch

from .. import ops as _ops


class _RayGenFn(torch.autograd.Function):
    """(pose[3,4], K^-1, px, py) -> rays_o [B,3], rays_d [B,3]; backward -> d pose[3,4]"""

    @staticmethod
    def forward(ctx, pose34, intr_inv, px, py):
        p = pose34.detach().float().contiguous()
        rays_o, rays_d, _, _, _ = _ops.raygen_fwd(0, intr_inv, px, py, c2w34=p)
        ctx.save_for_backward(intr_inv, px, py, rays_o, rays_d)
        return rays_o, rays_d

    @staticmethod
    def backward(ctx, g_o, g_d):
        intr_inv, px, py, rays_o, rays_d = ctx.saved_tensors
        g34 = _ops.raygen_bwd(intr_inv, px, py, rays_o, rays_d, None if g_o is None else g_o.float().contiguous(),
                              None if g_d is None else g_d.float().contiguous(), None, None)
        return g34, None, None, None


class _RayGenNearFarFn(torch.autograd.Function):
    """as _RayGenFn, plus near / far of Dataset.near_far_from_sphere (models/dataset.py:835-842), which the kernel
    computes anyway: (pose[3,4], K^-1, px, py) -> rays_o, rays_d [B,3], near, far [B,1]; backward -> d pose[3,4]"""

    @staticmethod
    def forward(ctx, pose34, intr_inv, px, py):
        p = pose34.detach().float().contiguous()
        rays_o, rays_d, near, far, _ = _ops.raygen_fwd(0, intr_inv, px, py, c2w34=p)
        ctx.save_for_backward(intr_inv, px, py, rays_o, rays_d)
        return rays_o, rays_d, near, far

    @staticmethod
    def backward(ctx, g_o, g_d, g_near, g_far):
        intr_inv, px, py, rays_o, rays_d = ctx.saved_tensors
        c = lambda g: None if g is None else g.float().contiguous()
        g34 = _ops.raygen_bwd(intr_inv, px, py, rays_o, rays_d, c(g_o), c(g_d), c(g_near), c(g_far))
        return g34, None, None, None


class RayDataset:
    def __init__(self, images, masks, intrinsics, device="cuda"):
        """images [N,H,W,3] fp32 in [0,1], masks [N,H,W,3] fp32, intrinsics [N,3,3] or [3,3]."""
        self.device = torch.device(device)
        self.images = images.to(self.device)
        self.masks = masks.to(self.device)
        self.masks_np = None
        self.n_images, self.H, self.W = images.shape[0], images.shape[1], images.shape[2]
        intr = torch.as_tensor(intrinsics, dtype=torch.float32)
        if intr.dim() == 2:
            intr = intr[None].repeat(self.n_images, 1, 1)
        self.intrinsics_all = intr.to(self.device)
        self.intrinsics_all_inv = torch.linalg.inv(intr).contiguous().to(self.device)
        self.use_mono_depth = False
        # flow matches (dataset.py:324-415 builds these from LoFTR files on disk; here they are set from memory)
        self.index_to_frame = {i: "%04d" % i for i in range(self.n_images)}
        self.frame_to_index = {v: k for k, v in self.index_to_frame.items()}
        self.flow_pairs = {}
        self.loftr_interval_flows = {}

    def set_flows(self, matches):
        """matches: {(idx_a, idx_b): (xs_a, ys_a, xs_b, ys_b)} numpy float arrays of matched pixels.  Fills
        `flow_pairs` / `loftr_interval_flows` with both directions, keyed as the reference (dataset.py:395-415)."""
        for (a, b), (xa, ya, xb, yb) in matches.items():
            fa, fb = self.index_to_frame[int(a)], self.index_to_frame[int(b)]
            arr = [np.asarray(v, dtype=np.float32) for v in (xa, ya, xb, yb)]
            self.loftr_interval_flows[fa + "_" + fb] = (arr[0], arr[1], arr[2], arr[3])
            self.loftr_interval_flows[fb + "_" + fa] = (arr[2], arr[3], arr[0], arr[1])
            self.flow_pairs.setdefault(fa, set()).add(fb)
            self.flow_pairs.setdefault(fb, set()).add(fa)

    def gen_random_ray_pairs_at(self, img_id_corr, batch_size, pose_network, current_img_num, interval=1, img_id=None,
                                indexs=None):
        """Dataset.gen_random_ray_pairs_at (models/dataset.py:683-792): `batch_size` matched pixel pairs between
        frame img_id_corr and a random matched frame img_id -> (data [2*batch_size,10] = [rays of img_id_corr | rays of
        img_id], pixels_xy, pixels_xy_corr, img_id, depth) or five Nones when there is no usable pair.  Matches are
        sub-pixel floats: colours are read at the truncated pixel, ray directions use the float coordinates.
        `img_id` / `indexs` inject the two host RNG draws (np.random.choice) for reproducible runs."""
        img_id_corr = int(img_id_corr)
        name_corr = self.index_to_frame[img_id_corr]
        if name_corr not in self.flow_pairs:
            return None, None, None, None, None
        pairs_idx = sorted(self.frame_to_index[n] for n in self.flow_pairs[name_corr])
        pairs_idx = [i for i in pairs_idx if i < current_img_num and abs(i - img_id_corr) <= interval]
        if len(pairs_idx) == 0:
            return None, None, None, None, None
        if img_id is None:
            img_id = int(np.random.choice(pairs_idx))
        img_id = int(img_id)
        xs1, ys1, xs2, ys2 = self.loftr_interval_flows[name_corr + "_" + self.index_to_frame[img_id]]
        if indexs is None:
            indexs = np.random.choice(len(xs1), batch_size, replace=True)
        dev = self.device
        px_c = torch.from_numpy(xs1[indexs]).to(dev)
        py_c = torch.from_numpy(ys1[indexs]).to(dev)
        px = torch.from_numpy(xs2[indexs]).to(dev)
        py = torch.from_numpy(ys2[indexs]).to(dev)
        color_c = self.images[img_id_corr][(py_c.long(), px_c.long())]
        color = self.images[img_id][(py.long(), px.long())]
        o_c, v_c = _RayGenFn.apply(pose_network(img_id_corr)[:3, :4], self.intrinsics_all_inv[img_id_corr], px_c, py_c)
        o, v = _RayGenFn.apply(pose_network(img_id)[:3, :4], self.intrinsics_all_inv[img_id], px, py)
        mask = torch.ones(2 * len(indexs), 1, device=dev)     # matches were filtered to the masks (dataset.py:770-772)
        data = torch.cat([torch.cat([o_c, o], 0), torch.cat([v_c, v], 0), torch.cat([color_c, color], 0), mask], dim=-1)
        depth = torch.zeros(2 * len(indexs), device=dev)
        return (data, torch.stack([px, py], dim=-1), torch.stack([px_c, py_c], dim=-1),
                torch.tensor(img_id, device=dev).long(), depth)

    def _bbox(self, img_idx, patch_size):
        if self.masks_np is None:
            self.masks_np = self.masks[..., 0].cpu().numpy()
        ys, xs = np.where(self.masks_np[int(img_idx)] > 0.5)
        return (max(ys.min() - patch_size, 0), min(ys.max() + patch_size, self.H),
                max(xs.min() - patch_size, 0), min(xs.max() + patch_size, self.W))

    def gen_random_rays_at(self, img_idx, batch_size, pose, mask_guided_sampling=False, patch_size=30, pixels=None,
                           img_idx_t=None, with_near_far=False):
        """-> (data [B,10] = rays_o, rays_v, rgb, mask ; depth=None).  `pixels=(px,py)` injects the int64 pixel
        draw (for RNG parity / pinned-host pipelines); otherwise torch.randint on the device as the reference.
        `img_idx_t` (int64 device tensor [1]) makes the frame a device-side input (CUDA-graph replay).
        `with_near_far`: also return the kernel's near / far (-> data, None, near, far), saving the torch
        near_far_from_sphere call and its backward."""
        img_idx = int(img_idx)
        if pixels is None:
            if mask_guided_sampling and np.random.rand() < 0.7:
                ys_min, ys_max, xs_min, xs_max = self._bbox(img_idx, patch_size)
            else:
                ys_min, ys_max, xs_min, xs_max = 0, self.H, 0, self.W
            px = torch.randint(low=xs_min, high=xs_max, size=[batch_size], device=self.device)
            py = torch.randint(low=ys_min, high=ys_max, size=[batch_size], device=self.device)
        else:
            px, py = pixels
        if img_idx_t is not None:
            it = img_idx_t.reshape(1)
            color = self.images[it, py, px]
            mask = self.masks[it, py, px]
            intr_inv = self.intrinsics_all_inv.index_select(0, it)[0]
        else:
            color = self.images[img_idx][(py, px)]
            mask = self.masks[img_idx][(py, px)]
            intr_inv = self.intrinsics_all_inv[img_idx]
        if with_near_far:
            rays_o, rays_v, near, far = _RayGenNearFarFn.apply(pose[:3, :4], intr_inv, px, py)
            return torch.cat([rays_o, rays_v, color, mask[:, :1]], dim=-1), None, near, far
        rays_o, rays_v = _RayGenFn.apply(pose[:3, :4], intr_inv, px, py)
        return torch.cat([rays_o, rays_v, color, mask[:, :1]], dim=-1), None

    def gen_rays_at(self, img_idx, resolution_level=1, pose=None, with_mask=False):
        """full-frame rays [H/l, W/l, 3] (dataset.py:547-576); pixel grid = linspace(0, W-1, W//l)."""
        l = resolution_level
        img_idx = int(img_idx)
        tx = torch.linspace(0, self.W - 1, self.W // l, device=self.device)
        ty = torch.linspace(0, self.H - 1, self.H // l, device=self.device)
        if l == 1:
            px, py = torch.meshgrid(tx.long(), ty.long(), indexing="ij")          # [W,H] integer pixel centres
        else:       # linspace(0, W-1, W//l) is not integer: sub-pixel ray directions, colours at the truncated pixel
            px, py = torch.meshgrid(tx.float(), ty.float(), indexing="ij")
        rays_o, rays_v = _RayGenFn.apply(pose[:3, :4], self.intrinsics_all_inv[img_idx], px.reshape(-1).contiguous(),
                                         py.reshape(-1).contiguous())
        px, py = px.long(), py.long()
        Wl, Hl = px.shape
        rays_o = rays_o.reshape(Wl, Hl, 3).transpose(0, 1)
        rays_v = rays_v.reshape(Wl, Hl, 3).transpose(0, 1)
        if with_mask:
            mask = self.masks[img_idx][(py, px)][..., 0].transpose(0, 1)
            return rays_o, rays_v, mask
        return rays_o, rays_v

    def near_far_from_sphere(self, rays_o, rays_d):
        a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
        b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
        mid = 0.5 * (-b) / a
        return mid - 1.0, mid + 1.0


# ---- the reference's own loader with the ray functions on the kernels ---------------------------------------------
def make_dataset_class(ref_dataset_cls):
    """`Dataset.__init__` of the reference (models/dataset.py:146-545: images, masks, cameras, LoFTR matches, mask-based
    pose initialisation — disk I/O and one-off preprocessing) is kept as it is; the subclass returned here re-routes the
    per-iteration ray functions to the fused pose / ray-generation kernel.  Pixel draws use the same RNG calls as the
    reference (np.random.rand, torch.randint under exp_runner.py's default tensor type), so a run draws the same pixels."""

    class Dataset(ref_dataset_cls):
        def _intr_inv(self, img_idx):
            return self.intrinsics_all_inv[img_idx, :3, :3].float().contiguous()

        def gen_random_rays_at(self, img_idx, batch_size, pose, mask_guided_sampling=False, patch_size=30):
            """models/dataset.py:634-681"""
            if getattr(self, "use_mono_depth", False):
                raise NotImplementedError("use_mono_depth = True (per-ray monocular depth, models/dataset.py:673-677) is not "
                                          "on the B200 path: no shipped conf enables it")
            self.images = self.images.to("cuda:0")
            self.masks = self.masks.to("cuda:0")
            if mask_guided_sampling and np.random.rand() < 0.7:
                mask = self.masks_np[img_idx][:, :, 0]
                ys, xs = np.where(mask > 0.5)
                ys_min, ys_max = max(ys.min() - patch_size, 0), min(ys.max() + patch_size, self.H)
                xs_min, xs_max = max(xs.min() - patch_size, 0), min(xs.max() + patch_size, self.W)
            else:
                ys_min, ys_max, xs_min, xs_max = 0, self.H, 0, self.W
            pixels_x = torch.randint(low=xs_min, high=xs_max, size=[batch_size]).to(self.images.device)
            pixels_y = torch.randint(low=ys_min, high=ys_max, size=[batch_size]).to(self.images.device)
            color = self.images[img_idx][(pixels_y, pixels_x)]
            mask = self.masks[img_idx][(pixels_y, pixels_x)]
            rays_o, rays_v = _RayGenFn.apply(pose[:3, :4], self._intr_inv(img_idx), pixels_x, pixels_y)
            return torch.cat([rays_o, rays_v, color, mask[:, :1]], dim=-1), None

        def gen_rays_at(self, img_idx, resolution_level=1, pose=None, with_mask=False):
            """models/dataset.py:547-576"""
            if pose is None:
                pose = self.pose_all[img_idx]
            l = resolution_level
            dev = self.intrinsics_all_inv.device
            tx = torch.linspace(0, self.W - 1, self.W // l, device=dev)
            ty = torch.linspace(0, self.H - 1, self.H // l, device=dev)
            pixels_x, pixels_y = torch.meshgrid(tx, ty, indexing="ij")             # [W,H], float (sub-pixel when l > 1)
            rays_o, rays_v = _RayGenFn.apply(pose[:3, :4].float(), self._intr_inv(img_idx),
                                             pixels_x.reshape(-1).float().contiguous(),
                                             pixels_y.reshape(-1).float().contiguous())
            Wl, Hl = pixels_x.shape
            rays_o = rays_o.reshape(Wl, Hl, 3).transpose(0, 1)
            rays_v = rays_v.reshape(Wl, Hl, 3).transpose(0, 1)
            if with_mask:
                mask = self.masks[img_idx].to(dev)[(pixels_y.long(), pixels_x.long())]
                return rays_o, rays_v, mask[..., 0].transpose(0, 1)
            return rays_o, rays_v

        def gen_random_ray_pairs_at(self, img_id_corr, batch_size, pose_network, current_img_num, interval=1):
            """models/dataset.py:683-792: same host RNG draws (np.random.choice twice), rays from the sub-pixel kernel"""
            if getattr(self, "use_mono_depth", False):
                raise NotImplementedError("use_mono_depth = True is not on the B200 path (no shipped conf enables it)")
            ic = int(img_id_corr)
            img_name_corr = self.index_to_frame[ic]
            if img_name_corr not in self.flow_pairs:
                return None, None, None, None, None
            pairs_idx = [self.frame_to_index[n] for n in self.flow_pairs[img_name_corr]]
            pairs_idx = [i for i in pairs_idx if i < current_img_num and abs(i - ic) <= interval]
            if len(pairs_idx) == 0:
                return None, None, None, None, None
            dev = torch.device("cuda")
            img_id = torch.tensor(np.random.choice(pairs_idx)).long().to(dev)
            xs1, ys1, xs2, ys2 = self.loftr_interval_flows[img_name_corr + "_" + self.index_to_frame[int(img_id)]]
            indexs = np.random.choice(len(xs1), batch_size, replace=True)
            px_c = torch.from_numpy(xs1[indexs]).to(dev).float()
            py_c = torch.from_numpy(ys1[indexs]).to(dev).float()
            px = torch.from_numpy(xs2[indexs]).to(dev).float()
            py = torch.from_numpy(ys2[indexs]).to(dev).float()
            images = self.images.to(dev)
            color_c = images[ic][(py_c.long(), px_c.long())]
            color = images[int(img_id)][(py.long(), px.long())]
            o_c, v_c = _RayGenFn.apply(pose_network(img_id_corr)[:3, :4].float(), self._intr_inv(ic), px_c, py_c)
            o, v = _RayGenFn.apply(pose_network(img_id)[:3, :4].float(), self._intr_inv(int(img_id)), px, py)
            mask = torch.ones(2 * batch_size, 1, device=dev)          # the matches were filtered to the masks (:770-772)
            data = torch.cat([torch.cat([o_c, o], 0), torch.cat([v_c, v], 0), torch.cat([color_c, color], 0), mask], dim=-1)
            return (data, torch.stack([px, py], dim=-1), torch.stack([px_c, py_c], dim=-1), img_id,
                    torch.zeros(2 * batch_size, device=dev))

        def near_far_from_sphere(self, rays_o, rays_d):
            """models/dataset.py:835-842 (carries gradients to the pose through rays_o / rays_d)"""
            a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
            b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
            mid = 0.5 * (-b) / a
            return mid - 1.0, mid + 1.0

    Dataset.__qualname__ = "Dataset"
    Dataset.__doc__ = "reference Dataset (loader) + B200 ray generation; see make_dataset_class"
    return Dataset


_REF_NAMES = ("load_K_Rt_from_P", "load_unit_K_Rt", "get_crop_M_ori", "shrink_mask", "get_center_radius", "origin_to_new",
              "save_point_cloud", "another_epe")


def __getattr__(name):
    """`from models.dataset import Dataset` (exp_runner.py:13) and `load_K_Rt_from_P` (utils/align_poses.py:4): resolved on
    first use from the reference checkout's own models/dataset.py, which fmov_pose_b200.dropin loads by file path."""
    if name == "Dataset" or name in _REF_NAMES:
        from .. import dropin as _dropin
        ref = _dropin.load_reference_module(os.path.join("models", "dataset.py"), "_fmov_reference_models_dataset")
        if name == "Dataset":
            cls = make_dataset_class(ref.Dataset)
            globals()["Dataset"] = cls
            return cls
        return getattr(ref, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
