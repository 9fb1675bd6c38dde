"""Import stub: models/camera.py:5 imports EasyDict but the hot path (Lie.se3_to_SE3,
Pose.compose) never uses it.  Test infrastructure only."""


class EasyDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__
