"""Shared helpers for the test-suite (fixture loading; oracle parameter dicts)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def params_from(d, prefix, dtype=torch.float32, requires_grad=False):
    """{'lin0.weight_g': tensor, ...} from fixture keys 'sdf.lin0.weight_g' ..."""
    out = {}
    for k, v in d.items():
        if k.startswith(prefix) and k[len(prefix):].startswith("lin"):
            t = torch.from_numpy(np.asarray(v)).to(dtype).clone()
            out[k[len(prefix):]] = t.requires_grad_(requires_grad)
    return out


def t(d, k, dtype=torch.float32):
    return torch.from_numpy(np.asarray(d[k])).to(dtype)
