"""LearnPoseGF / SegLearnPose — drop-in for models/picture_pose.py:13-250.

Same parameters (lin1, lin2, lin3 | lin3_rot / lin3_trans(frozen) / lin3_scale, init_c2w), same
Gaussian-Fourier features (b ~ N(0, embedding_scale^2) drawn from np.random exactly like the reference,
picture_pose.py:74-78), same control methods.  The whole forward — Fourier features, the 21 k-parameter GELU MLP, its
heads and the tail (Rodrigues exponential + composition with the optionally scaled initial pose, picture_pose.py:140-186)
— is ONE launch per direction, fmov_pose_gf_fwd/_bwd (csrc/pose_raygen.cu): the torch formulation costs ~45 launches per
frame, a quarter of the shipped 512-ray iteration."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops as _ops


class _PoseGfFn(torch.autograd.Function):
    """Whole LearnPoseGF.forward (picture_pose.py:140-186) as one launch per direction (fmov_pose_gf_fwd/_bwd):
    (cid int64 [1], b, init_c2w [N,4,4] | None, rot_k, W1, b1, W2, b2, head W/b ...) -> c2w [3,4]"""

    @staticmethod
    def forward(ctx, cid_t, b, init_all, rot_k, *params):
        ps = [p.detach().float().contiguous() for p in params]
        heads = [(ps[4 + 2 * i], ps[5 + 2 * i]) for i in range((len(ps) - 4) // 2)]
        b_c = b.detach().float().contiguous().reshape(-1)
        init_c = None if init_all is None else init_all.detach().float().contiguous()
        c2w, save = _ops.pose_gf_fwd(cid_t, b_c, ps[0], ps[1], ps[2], ps[3], heads, float(rot_k), init_c)
        ctx.save_for_backward(cid_t, b_c, save, *ps, *(() if init_c is None else (init_c,)))
        ctx.n_params, ctx.rot_k, ctx.has_init = len(ps), float(rot_k), init_c is not None
        return c2w

    @staticmethod
    def backward(ctx, g):
        sv = ctx.saved_tensors
        cid_t, b_c, save = sv[0], sv[1], sv[2]
        ps = list(sv[3:3 + ctx.n_params])
        init_c = sv[3 + ctx.n_params] if ctx.has_init else None
        heads = [(ps[4 + 2 * i], ps[5 + 2 * i]) for i in range((len(ps) - 4) // 2)]
        need = list(ctx.needs_input_grad[4:])
        grads = _ops.pose_gf_bwd(cid_t, b_c, ps[0], ps[1], ps[2], ps[3], heads, ctx.rot_k, init_c, save,
                                 g.float().contiguous(), need)
        return (None, None, None, None) + tuple(grads)


class LearnPoseGF(nn.Module):
    def __init__(self, num_cams, init_c2w=None, pose_encoding=False, embedding_scale=10, emphasize_rot=False,
                 small_rot=False):
        super().__init__()
        self.emphasize_rot = emphasize_rot
        self.num_cams = num_cams
        self.embedding_size = 128
        self.init_c2w = nn.Parameter(init_c2w, requires_grad=False) if init_c2w is not None else None
        self.lin1 = nn.Linear(self.embedding_size * 2, 64)
        self.gelu1 = nn.GELU()
        self.lin2 = nn.Linear(64, 64)
        self.gelu2 = nn.GELU()
        if not emphasize_rot:
            self.lin3 = nn.Linear(64, 6)
            nn.init.normal_(self.lin3.weight, mean=0, std=0.01)
            nn.init.zeros_(self.lin3.bias)
        else:
            self.lin3_rot = nn.Linear(64, 3)
            nn.init.normal_(self.lin3_rot.weight, mean=0, std=0.01)
            nn.init.zeros_(self.lin3_rot.bias)
            self.lin3_trans = nn.Linear(64, 3)
            nn.init.zeros_(self.lin3_trans.weight)
            nn.init.zeros_(self.lin3_trans.bias)
            for p in self.lin3_trans.parameters():
                p.requires_grad = False
            self.lin3_scale = nn.Linear(64, 1)
            nn.init.normal_(self.lin3_scale.weight, mean=0, std=0.01)
            nn.init.ones_(self.lin3_scale.bias)
        self.embedding_scale = embedding_scale
        if pose_encoding:
            b = 2.0 ** np.linspace(0, 5, self.embedding_size // 2) - 1.0
            b = b[:, np.newaxis]
            b = np.concatenate([b, np.roll(b, 1, axis=-1)], 0) + 0
        else:
            b = np.random.normal(loc=0.0, scale=self.embedding_scale, size=[self.embedding_size, 1])
        # the reference ends up with `b` registered as a frozen Parameter (picture_pose.py:82: Parameter(...).to(dev)
        # returns the Parameter itself when it already lives on `dev`), so it is part of the state_dict
        self.b = nn.Parameter(torch.tensor(b).float(), requires_grad=False)
        self.small_rot = small_rot
        self._bottom = None         # cached [[0,0,0,1]] row on the module's device (no H2D copy per call)

    # ---- control surface used by exp_runner.py ----------------------------------------------------------
    def finish_warmup(self):
        pass

    def disable_trans(self):
        for p in self.lin3_scale.parameters():
            p.requires_grad = False

    def enable_trans(self):
        for p in self.lin3_scale.parameters():
            p.requires_grad = True

    def _heads(self):
        if not self.emphasize_rot:
            return [self.lin1, self.lin2, self.lin3]
        return [self.lin1, self.lin2, self.lin3_rot, self.lin3_trans, self.lin3_scale]

    def disable_grad(self):
        for m in self._heads():
            for p in m.parameters():
                p.requires_grad = False

    def enable_grad(self):
        for m in self._heads():
            for p in m.parameters():
                p.requires_grad = True

    def forward(self, cam_id, cam_id_t=None):
        """`cam_id_t`: optional int64 device tensor [1] holding cam_id.  When given, the frame index is read on the
        device only (Fourier features and the init_c2w row), which keeps the call CUDA-graph capturable with the
        frame as a replay-time input (train.GraphedTrainStep)."""
        if cam_id_t is not None:
            cid_t = cam_id_t.reshape(1)
        else:
            cid_t = torch.as_tensor(int(cam_id), dtype=torch.int64, device=self.b.device).reshape(1)
        if not self.emphasize_rot:
            heads = [self.lin3]
        else:
            heads = [self.lin3_rot, self.lin3_trans, self.lin3_scale]
        params = [self.lin1.weight, self.lin1.bias, self.lin2.weight, self.lin2.bias]
        for h in heads:
            params += [h.weight, h.bias]
        k = np.pi / 6 if self.small_rot else np.pi
        c2w34 = _PoseGfFn.apply(cid_t, self.b, self.init_c2w, float(k), *params)
        if self._bottom is None or self._bottom.device != c2w34.device:
            self._bottom = torch.cat([torch.zeros(1, 3, device=c2w34.device), torch.ones(1, 1, device=c2w34.device)], 1)
        return torch.cat([c2w34, self._bottom], dim=0)          # (4,4) like the reference


class SegLearnPose(nn.Module):
    def __init__(self, num_cams, segment_img_num, init_c2w=None, pose_encoding=False, embedding_scale=10,
                 emphasize_rot=False, small_rot=False):
        super().__init__()
        self.num_cams = num_cams
        self.segment_img_num = segment_img_num
        n = init_c2w.shape[0] // segment_img_num + (1 if init_c2w.shape[0] % segment_img_num else 0)
        self.pose_mlps = nn.ModuleList(
            [LearnPoseGF(num_cams, init_c2w.clone(), pose_encoding, embedding_scale, emphasize_rot, small_rot)
             for _ in range(n)])
        self.initialized_flag = nn.Parameter(torch.tensor([True] + [False] * (n - 1)), requires_grad=False)
        self.progress = nn.Parameter(torch.zeros(n), requires_grad=False)
        self._flags_host = None      # host mirror of initialized_flag (no device read-back per call)

    def _load_from_state_dict(self, *args, **kwargs):
        self._flags_host = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _flag(self, k):
        if self._flags_host is None:
            self._flags_host = [bool(v) for v in self.initialized_flag.tolist()]
        return self._flags_host[k]

    def _set_flag(self, k):
        self._flag(k)
        self._flags_host[k] = True
        self.initialized_flag[k] = True

    def forward(self, cam_id, cam_id_t=None):
        cid = int(cam_id)
        k = cid // self.segment_img_num
        if not self._flag(k):
            self._set_flag(k)
            with torch.no_grad():      # picture_pose.py:227-235: seed the new segment with the previous pose
                last_pose = self.pose_mlps[k - 1](k * self.segment_img_num - 1)
                last44 = torch.eye(4, device=last_pose.device)
                last44[:3] = last_pose[:3]
                self.pose_mlps[k].init_c2w.data.copy_(last44.clone().repeat(self.num_cams, 1, 1))
        return self.pose_mlps[k](cid, cam_id_t)

    def set_pose(self, cam_id, pose, force_update=False):
        k = cam_id // self.segment_img_num
        if not self._flag(k) or force_update:
            self._set_flag(k)
            with torch.no_grad():
                self.pose_mlps[k].init_c2w.data.copy_(pose.clone().repeat(self.num_cams, 1, 1))

    def step_progress(self, pose_mlp_index):
        self.progress[pose_mlp_index] += 1
        return self.progress[pose_mlp_index]
