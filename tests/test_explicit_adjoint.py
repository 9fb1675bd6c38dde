"""The closed-form backward the CUDA kernels transcribe (oracle/explicit_adjoint.py) must agree
with autograd through the oracle's restatement of the reference.  CPU, fp64."""
import numpy as np
import torch

from oracle import explicit_adjoint as E
from oracle import neus_oracle as O
from tests._util import load_golden, params_from


def _nets():
    d = load_golden("full_6464_gf")
    sdf_p = params_from(d, "sdf.", dtype=torch.float64)
    col_p = params_from(d, "col.", dtype=torch.float64)
    return sdf_p, col_p


def test_sdf_forward_normal_backward_match_autograd():
    torch.manual_seed(0)
    sdf_p, _ = _nets()
    W = [O.eff_weight(sdf_p, "", l).clone().requires_grad_(True) for l in range(9)]
    b = [sdf_p[f"lin{l}.bias"].clone().requires_grad_(True) for l in range(9)]
    P = 96
    x = (torch.rand(P, 3, dtype=torch.float64) * 2 - 1).requires_grad_(True)
    # autograd path with plain weights
    p2 = {}
    for l in range(9):
        p2[f"lin{l}.weight"] = W[l]
        p2[f"lin{l}.bias"] = b[l]
    out = O.sdf_forward(p2, x)
    n_ag = O.sdf_gradient(p2, x)
    r1, r2, r3 = torch.randn(P, dtype=torch.float64), torch.randn(P, 256, dtype=torch.float64) * 0.1, \
        torch.randn(P, 3, dtype=torch.float64)
    loss = (out[:, 0] * r1).sum() + (out[:, 1:] * r2).sum() + (n_ag * r3).sum()
    grads = torch.autograd.grad(loss, [x] + W + b)
    with torch.no_grad():
        st = E.sdf_forward(W, b, x)
        n, stn = E.sdf_normal(W, x, st)
        np.testing.assert_allclose(st["sdf"].numpy(), out[:, 0].detach().numpy(), atol=1e-12)
        np.testing.assert_allclose(st["feat"].numpy(), out[:, 1:].detach().numpy(), atol=1e-12)
        np.testing.assert_allclose(n.numpy(), n_ag.detach().numpy(), atol=1e-11)
        xbar, dW, db = E.sdf_backward(W, x, st, stn, r1, r2, r3)
    np.testing.assert_allclose(xbar.numpy(), grads[0].numpy(), rtol=1e-8, atol=1e-9)
    for l in range(9):
        np.testing.assert_allclose(dW[l].numpy(), grads[1 + l].numpy(), rtol=1e-8, atol=1e-9, err_msg=f"dW{l}")
        np.testing.assert_allclose(db[l].numpy(), grads[10 + l].numpy(), rtol=1e-8, atol=1e-9, err_msg=f"db{l}")


def test_color_forward_backward_match_autograd():
    torch.manual_seed(1)
    _, col_p = _nets()
    W = [O.eff_weight(col_p, "", l).clone().requires_grad_(True) for l in range(5)]
    b = [col_p[f"lin{l}.bias"].clone().requires_grad_(True) for l in range(5)]
    P = 64
    mk = lambda *s: torch.randn(*s, dtype=torch.float64).requires_grad_(True)
    pts, dirs, nrm, feat = mk(P, 3), mk(P, 3), mk(P, 3), mk(P, 256)
    p2 = {}
    for l in range(5):
        p2[f"lin{l}.weight"] = W[l]
        p2[f"lin{l}.bias"] = b[l]
    rgb = O.color_forward(p2, pts, nrm, dirs, feat)
    r = torch.randn(P, 3, dtype=torch.float64)
    grads = torch.autograd.grad((rgb * r).sum(), [pts, dirs, nrm, feat] + W + b)
    with torch.no_grad():
        rgb2, st = E.color_forward(W, b, pts, dirs, nrm, feat)
        np.testing.assert_allclose(rgb2.numpy(), rgb.detach().numpy(), atol=1e-13)
        pb, dbar, nb, fb, dW, db = E.color_backward(W, dirs, st, r)
    for got, ref in zip([pb, dbar, nb, fb] + dW + db, grads):
        np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-9, atol=1e-11)
