"""Pose / Lie — the two classes of models/camera.py the train step uses (camera.py:8-60, 63-156):
`lie.se3_to_SE3` (10-term Taylor A/B/C) and `pose.compose`.  The BARF pose path of the train step
(exp_runner.py:419-424) runs in fmov_pose_fwd/_bwd mode 2; `barf_pose(...)` below is the differentiable
entry point for one frame.  The torch implementations serve batched host-side use (all frames at once, as
exp_runner.py:419-424 does) and validation code."""
import torch

from .. import ops as _ops


class Pose:
    def __call__(self, R=None, t=None):
        assert R is not None or t is not None
        if R is None:
            t = torch.as_tensor(t)
            R = torch.eye(3, device=t.device).repeat(*t.shape[:-1], 1, 1)
        elif t is None:
            R = torch.as_tensor(R)
            t = torch.zeros(R.shape[:-1], device=R.device)
        R, t = torch.as_tensor(R).float(), torch.as_tensor(t).float()
        return torch.cat([R, t[..., None]], dim=-1)

    def invert(self, pose, use_inverse=False):
        R, t = pose[..., :3], pose[..., 3:]
        R_inv = R.inverse() if use_inverse else R.transpose(-1, -2)
        return self(R=R_inv, t=(-R_inv @ t)[..., 0])

    def compose(self, pose_list):
        pose_new = pose_list[0]
        for p in pose_list[1:]:
            pose_new = self.compose_pair(pose_new, p)
        return pose_new

    def compose_pair(self, pose_a, pose_b):
        R_a, t_a = pose_a[..., :3], pose_a[..., 3:]
        R_b, t_b = pose_b[..., :3], pose_b[..., 3:]
        return self(R=R_b @ R_a, t=(R_b @ t_a + t_b)[..., 0])


class Lie:
    def skew_symmetric(self, w):
        w0, w1, w2 = w.unbind(dim=-1)
        O = torch.zeros_like(w0)
        return torch.stack([torch.stack([O, -w2, w1], dim=-1), torch.stack([w2, O, -w0], dim=-1),
                            torch.stack([-w1, w0, O], dim=-1)], dim=-2)

    def _taylor(self, x, first, nth=10):
        ans = torch.zeros_like(x)
        denom = 1.0
        for i in range(nth + 1):
            if first == 1:
                if i > 0:
                    denom *= (2 * i) * (2 * i + 1)
            else:
                denom *= (2 * i + first - 1) * (2 * i + first)
            ans = ans + (-1) ** i * x ** (2 * i) / denom
        return ans

    def taylor_A(self, x, nth=10):
        return self._taylor(x, 1, nth)

    def taylor_B(self, x, nth=10):
        return self._taylor(x, 2, nth)

    def taylor_C(self, x, nth=10):
        return self._taylor(x, 3, nth)

    def so3_to_SO3(self, w):
        wx = self.skew_symmetric(w)
        theta = w.norm(dim=-1)[..., None, None]
        I = torch.eye(3, device=w.device, dtype=torch.float32)
        return I + self.taylor_A(theta) * wx + self.taylor_B(theta) * wx @ wx

    def se3_to_SE3(self, wu, only_rot=False):
        w, u = wu.split([3, 3], dim=-1)
        wx = self.skew_symmetric(w)
        theta = w.norm(dim=-1)[..., None, None]
        I = torch.eye(3, device=w.device, dtype=torch.float32)
        A, B, C = self.taylor_A(theta), self.taylor_B(theta), self.taylor_C(theta)
        R = I + A * wx + B * wx @ wx
        V = I + B * wx + C * wx @ wx
        t = V @ u[..., None]
        if only_rot:
            t = torch.zeros_like(t.detach())
        return torch.cat([R, t], dim=-1)


class _BarfPoseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, se3, noise34):
        se3c, n = se3.detach().float().contiguous(), noise34.detach().float().contiguous()
        ctx.save_for_backward(se3c, n)
        return _ops.pose_fwd(2, se3=se3c, init34=n)

    @staticmethod
    def backward(ctx, g):
        se3c, n = ctx.saved_tensors
        _, _, _, g_se3 = _ops.pose_bwd(2, g.float().contiguous(), se3=se3c, init34=n)
        return g_se3, None


def barf_pose(se3_row, noise_pose):
    """compose([se3_to_SE3(se3_row), noise_pose[:3]]) for one frame (exp_runner.py:419-424) -> [3,4]"""
    n = noise_pose
    if n.shape[-2] == 3:
        # device-side fills only (CUDA-graph capturable; torch.tensor([...], device=...) is a pageable H2D copy)
        bottom = torch.cat([torch.zeros(1, 3, device=n.device), torch.ones(1, 1, device=n.device)], dim=1)
        n = torch.cat([n, bottom], dim=0)
    return _BarfPoseFn.apply(se3_row, n)


pose = Pose()
lie = Lie()


def to_hom(X):
    """[..., 3] -> [..., 4] homogeneous coordinates (models/camera.py:266-270; used by exp_runner.py:638, 671)"""
    return torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)
