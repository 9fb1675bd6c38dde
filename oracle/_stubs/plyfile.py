"""Import stub: the reference's models/dataset.py:8 imports plyfile for `save_point_cloud` (a debugging export that the
train step never calls).  Test infrastructure only."""


class PlyData:          # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("plyfile is not installed; point-cloud export is outside the hot path")


class PlyElement:       # pragma: no cover
    @staticmethod
    def describe(*a, **k):
        raise RuntimeError("plyfile is not installed; point-cloud export is outside the hot path")
