"""Precision study behind the FMOV_RECOMPUTE_Q build switch of csrc/mlp_fine.cu (CPU emulation, fp64 truth).

The default backward kernel stores the second-order injection q_l = 100 * dbar_l * delta_l * (1 - sigma_l) as an fp16
tile; the experiment rebuilds it in the ordinary backward pass from tiles that exist anyway:
    q_l = vbar_{l+1} * delta_l * 100 * (1 - sigma_l) / sigma_l,      vbar_{l+1} = dbar_l * sigma_l  (fp16, loss-scaled)
with sigma_l = 1 - exp(-100 h_{l+1}) taken from the stored fp16 activation.  This test rounds the operands exactly as the
kernels store them and checks that the rebuilt q is as close to the closed form (oracle/explicit_adjoint.py:sdf_backward)
as the stored one, including where sigma underflows to 0 (q must be 0 there, not NaN)."""
import numpy as np
import torch

from oracle import explicit_adjoint as E
from oracle import neus_oracle as O
from tests._util import load_golden, params_from


def _h16(t):
    return t.to(torch.float16).to(torch.float64)


def _sat16(t):
    return torch.clamp(t, -65504.0, 65504.0).to(torch.float16).to(torch.float64)


def test_rebuilt_q_is_as_accurate_as_the_stored_q():
    torch.manual_seed(0)
    d = load_golden("full_6464_gf")
    sdf_p = params_from(d, "sdf.", dtype=torch.float64)
    W = [O.eff_weight(sdf_p, "", l) for l in range(9)]
    b = [sdf_p[f"lin{l}.bias"] for l in range(9)]
    P = 2048
    x = torch.randn(P, 3, dtype=torch.float64) * 0.5
    st = E.sdf_forward(W, b, x)
    _, stn = E.sdf_normal(W, x, st)
    sig, z, delta = st["sig"], st["z"], stn["delta"]
    nbar = torch.randn(P, 3, dtype=torch.float64)
    gbar_e = E.pe_j_apply(x, nbar, 6)
    scale = 2.0 ** (8 - np.floor(np.log2(nbar.abs().max().item())))          # amax * scale in [2^8, 2^9)
    dbar = (gbar_e * scale) @ W[0].T
    for l in range(1, 9):
        h = _h16(E.softplus(z[l - 1]))                                       # the stored activation H_l
        e = torch.exp(-100.0 * h)
        s_k = 1.0 - e                                                        # sigma as the kernels see it
        q_ref = 100.0 * dbar * delta[l - 1] * (1 - sig[l - 1]) * (z[l - 1] * 100.0 <= 20.0)
        q_stored = _sat16(100.0 * dbar * _h16(delta[l - 1]) * (1 - s_k))
        vbar = _sat16(dbar * s_k)
        r = torch.where(s_k > 0, 100.0 * e / s_k.clamp_min(1e-300), torch.zeros_like(s_k))
        q_rebuilt = vbar * _h16(delta[l - 1]) * r
        assert torch.isfinite(q_rebuilt).all()
        den = q_ref.abs().max().item()
        err_stored = (q_stored - q_ref).abs().max().item() / den
        err_rebuilt = (q_rebuilt - q_ref).abs().max().item() / den
        assert err_rebuilt < 2e-3, (l, err_rebuilt)
        assert err_rebuilt < 2.5 * err_stored + 1e-4, (l, err_rebuilt, err_stored)
        # what matters for dW / db: column sums over the points
        cs = (q_rebuilt - q_ref).sum(0).abs().max().item() / q_ref.sum(0).abs().max().item()
        assert cs < 2e-3, (l, cs)
        if l < 8:
            ab = dbar * sig[l - 1]
            vb = torch.cat([ab[:, :217] / E.SQ2, gbar_e * scale / E.SQ2], 1) if l == 4 else ab
            dbar = vb @ W[l].T
    # sigma == 0 (padding columns / underflowed activations): every factor is 0 and so is q
    z0 = torch.zeros(4, dtype=torch.float64)
    r0 = torch.where(z0 > 0, 100.0 / z0.clamp_min(1e-300), torch.zeros_like(z0))
    assert (r0 == 0).all()
