"""Import stub: the reference's models/renderer.py imports PyMCubes at module scope
(renderer.py:6) but only calls it in extract_geometry (renderer.py:43), which is
downstream of the hot path.  Test infrastructure only."""


def marching_cubes(*a, **k):  # pragma: no cover
    raise RuntimeError("PyMCubes is not installed; marching cubes is out of scope (SURVEY.md §2 row 13)")
