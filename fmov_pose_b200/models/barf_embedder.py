"""API mirror of models/barf_embedder.py (get_embedder(multires) -> (embed(x, progress), out_dim))."""
from .embedder import Embedder, get_barf_embedder as get_embedder  # noqa: F401
