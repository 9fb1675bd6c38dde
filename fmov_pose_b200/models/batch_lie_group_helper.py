"""vec2skew / Exp / batch_make_c2w — API mirror of models/batch_lie_group_helper.py:6-47.
Host-side torch helpers (used by pose initialisation code and tests); the train step evaluates the same
Rodrigues formula inside fmov_pose_fwd / fmov_raygen_fwd (csrc/pose_raygen.cu)."""
import torch


def vec2skew(v):
    zero = torch.zeros(v.shape[0], 1, dtype=torch.float32, device=v.device)
    rows = [torch.cat([zero, -v[:, 2:3], v[:, 1:2]], dim=-1),
            torch.cat([v[:, 2:3], zero, -v[:, 0:1]], dim=-1),
            torch.cat([-v[:, 1:2], v[:, 0:1], zero], dim=-1)]
    return torch.stack(rows, dim=1)


def Exp(r):
    K = vec2skew(r)
    th = r.norm(dim=1, keepdim=True) + 1e-15
    eye = torch.eye(3, dtype=torch.float32, device=r.device).unsqueeze(0).repeat(r.shape[0], 1, 1)
    return eye + (torch.sin(th) / th)[..., None] * K + ((1 - torch.cos(th)) / th ** 2)[..., None] * (K @ K)


def batch_make_c2w(r, t):
    return torch.cat([Exp(r), t.unsqueeze(-1)], dim=2)
