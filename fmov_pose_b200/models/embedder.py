"""Positional encoders — API mirror of models/embedder.py:7-55 and models/barf_embedder.py:6-75.

The CUDA kernels evaluate the encoding in registers; these torch functions exist for API parity
(`get_embedder` is imported by user code) and for host-side checks.  Layout
[x, sin(2^k x), cos(2^k x)]_{k<L} (embedder.py:28-37).  The BARF variant takes a `progress` argument
but — exactly like the reference (barf_embedder.py:50-56) — does not apply the coarse-to-fine weights."""
import torch


class Embedder:
    def __init__(self, multires, input_dims=3):
        self.multires = multires
        self.input_dims = input_dims
        self.out_dim = input_dims * (1 + 2 * multires)
        self.freq_bands = 2.0 ** torch.linspace(0.0, multires - 1, multires) if multires > 0 else torch.zeros(0)

    def embed(self, inputs, progress=None):
        out = [inputs]
        for f in self.freq_bands.tolist():
            out.append(torch.sin(inputs * f))
            out.append(torch.cos(inputs * f))
        return torch.cat(out, -1)


def get_embedder(multires, input_dims=3):
    eo = Embedder(multires, input_dims)

    def embed(x, eo=eo):
        return eo.embed(x)

    return embed, eo.out_dim


def get_barf_embedder(multires, input_dims=3):
    eo = Embedder(multires, input_dims)

    def embed(x, progress, eo=eo):
        return eo.embed(x, progress)

    return embed, eo.out_dim
