"""Placeholders for models/pixel_pose.py:28-388 (`PixelPose`, `DeepPixelPose`, `SegDeepPixelPose`).

`exp_runner.py:25` imports `SegDeepPixelPose` unconditionally, so the name has to resolve; the per-pixel pose MLPs are only
built when `model.pixel_level = True` (`exp_runner.py:237`, default False — no shipped conf sets it) and are outside the
train-step hot path.  Constructing one fails loudly instead of silently running something else."""
import torch.nn as nn


class _Unsupported(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError(
            f"{type(self).__name__} (model.pixel_level = True) is not part of the B200 train-step path: every shipped conf "
            "uses the per-frame pose modules of models/picture_pose.py (LearnPoseGF / SegLearnPose)")


class PixelPose(_Unsupported):
    pass


class DeepPixelPose(_Unsupported):
    pass


class SegDeepPixelPose(_Unsupported):
    pass
