"""BarfSDFNetwork / BarfRenderingNetwork — drop-in for models/barf_fields.py:8-211.

Adds the `noise_poses` buffer-parameter, the `se3_refine` Embedding(n_images, 6) and the `progress`
scalar (barf_fields.py:28-39, :172-174).  `progress` has no numerical effect: the reference's BARF
embedder never applies its coarse-to-fine weights."""
import torch
import torch.nn as nn

from .fields import RenderingNetwork, SDFNetwork


class BarfSDFNetwork(SDFNetwork):
    def __init__(self, noise_poses, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False, n_images=0, barf=True):
        nn.Module.__init__(self)
        self.noise_poses = nn.Parameter(noise_poses.clone(), requires_grad=False)
        self.se3_refine = nn.Embedding(n_images, 6)
        nn.init.zeros_(self.se3_refine.weight)
        if not barf:
            self.se3_refine.weight.requires_grad = False
        self.progress = nn.Parameter(torch.tensor(0.0))
        _sdf_init_layers(self, d_in, d_out, d_hidden, n_layers, skip_in, multires, bias, scale, geometric_init,
                         weight_norm, inside_outside)


def _sdf_init_layers(self, *args):
    # build the layers exactly like SDFNetwork.__init__ on an already-initialised nn.Module
    tmp = SDFNetwork(*args)
    self.multires, self.num_layers, self.skip_in, self.scale = tmp.multires, tmp.num_layers, tmp.skip_in, tmp.scale
    for l in range(tmp.num_layers - 1):
        setattr(self, "lin" + str(l), getattr(tmp, "lin" + str(l)))


class BarfRenderingNetwork(RenderingNetwork):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.progress = nn.Parameter(torch.tensor(0.0))
