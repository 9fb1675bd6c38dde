"""Closed-form (no autograd) forward / normal / backward of the SDF and colour MLPs.

TEST INFRASTRUCTURE ONLY — the mathematical blueprint the CUDA kernels in
`fmov_pose_b200/csrc/` transcribe.  Checked against autograd of
`oracle/neus_oracle.py` in `tests/test_explicit_adjoint.py`.

Reference semantics restated: SDFNetwork.forward/.gradient (models/fields.py:88-124),
RenderingNetwork.forward (models/fields.py:166-193); the reference obtains the normal and its
backward with autograd (create_graph=True) — this file spells the same derivatives out.

Notation (per point): e = PE6(x) (39); layers l=0..8; u_l input, z_l = W_l u_l + b_l;
h_{l+1} = softplus_100(z_l) (l<8); sigma_l = softplus'(z_l); u_4 = [h_4(217); e]/sqrt2.
Reverse sweep for n = d sdf/dx: delta_7 = W_8[0,:] * sigma_7; v_l = W_l^T delta_l;
delta_{l-1} = a_l * sigma_{l-1} with a_l = v_l (l=4: a_4 = v_4[:217]/sqrt2, g_e += v_4[217:]/sqrt2);
g_e += W_0^T delta_0; n = J_e^T g_e.
"""
import math

import torch

BETA = 100.0
THRESH = 20.0
SQ2 = math.sqrt(2.0)


def pe(x, L):
    out = [x]
    for k in range(L):
        out += [torch.sin(x * 2.0 ** k), torch.cos(x * 2.0 ** k)]
    return torch.cat(out, -1)


def pe_jt_apply(x, g_e, L):
    """n = J_e(x)^T g_e  (g_e [P, 3+6L])"""
    n = g_e[:, 0:3].clone()
    for k in range(L):
        f = 2.0 ** k
        n = n + g_e[:, 3 + 6 * k: 6 + 6 * k] * (f * torch.cos(x * f))
        n = n - g_e[:, 6 + 6 * k: 9 + 6 * k] * (f * torch.sin(x * f))
    return n


def pe_j_apply(x, nbar, L):
    """gbar_e = J_e(x) nbar  (nbar [P,3]) -> [P, 3+6L]"""
    out = [nbar]
    for k in range(L):
        f = 2.0 ** k
        out += [nbar * (f * torch.cos(x * f)), -nbar * (f * torch.sin(x * f))]
    return torch.cat(out, -1)


def pe_hess_term(x, g_e, nbar, L):
    """d/dx of (J_e^T g_e) . nbar with g_e, nbar held fixed: sum_i g_e,i * d2 e_i/dx2 * nbar"""
    t = torch.zeros_like(x)
    for k in range(L):
        f = 2.0 ** k
        t = t - g_e[:, 3 + 6 * k: 6 + 6 * k] * (f * f * torch.sin(x * f)) * nbar
        t = t - g_e[:, 6 + 6 * k: 9 + 6 * k] * (f * f * torch.cos(x * f)) * nbar
    return t


def softplus(z):
    return torch.where(z * BETA > THRESH, z, torch.log1p(torch.exp(torch.clamp(z * BETA, max=THRESH))) / BETA)


def dsoftplus(z):
    return torch.where(z * BETA > THRESH, torch.ones_like(z), torch.sigmoid(z * BETA))


def d2softplus(z):
    s = torch.sigmoid(z * BETA)
    return torch.where(z * BETA > THRESH, torch.zeros_like(z), BETA * s * (1 - s))


def sdf_forward(W, b, x, L=6):
    """W,b: lists of 9 effective weights/biases. returns dict with sdf [P], feat [P,256], stash."""
    e = pe(x, L)
    u, z, sig = [], [], []
    h = e
    for l in range(9):
        if l == 4:
            h = torch.cat([h, e], 1) / SQ2
        u.append(h)
        zl = h @ W[l].T + b[l]
        z.append(zl)
        if l < 8:
            sig.append(dsoftplus(zl))
            h = softplus(zl)
    return dict(e=e, u=u, z=z, sig=sig, sdf=z[8][:, 0], feat=z[8][:, 1:])


def sdf_normal(W, x, st, L=6):
    """reverse sweep; returns n [P,3] and stash (delta list, g_e)."""
    sig = st["sig"]
    delta = [None] * 8
    delta[7] = W[8][0:1, :] * sig[7]
    n_pe = st["e"].shape[1]
    g_e = torch.zeros_like(st["e"])
    for l in range(7, 0, -1):
        v = delta[l] @ W[l]                       # [P, in_l]
        if l == 4:
            g_e = g_e + v[:, -n_pe:] / SQ2
            a = v[:, :-n_pe] / SQ2
        else:
            a = v
        delta[l - 1] = a * sig[l - 1]
    g_e = g_e + delta[0] @ W[0]
    n = pe_jt_apply(x, g_e, L)
    return n, dict(delta=delta, g_e=g_e)


def sdf_backward(W, x, st, st_n, sbar, fbar, nbar, L=6):
    """Given dL/dsdf [P], dL/dfeat [P,256], dL/dn [P,3] -> dL/dx [P,3], dW list, db list."""
    sig, u, z, delta, g_e = st["sig"], st["u"], st["z"], st_n["delta"], st_n["g_e"]
    n_pe = st["e"].shape[1]
    dW = [torch.zeros_like(w) for w in W]
    db = [torch.zeros(w.shape[0], dtype=w.dtype) for w in W]
    # 1. n = J_e^T g_e
    gbar_e = pe_j_apply(x, nbar, L)
    xbar = pe_hess_term(x, g_e, nbar, L)
    # 2. adjoint of the reverse sweep (forward-shaped pass l = 0..7)
    q = [None] * 8                                  # injection into zbar_l
    dbar = gbar_e @ W[0].T                          # delta-bar_0
    dW[0] += delta[0].T @ gbar_e
    for l in range(1, 8):
        # delta_{l-1} = a_l * sigma_{l-1}
        abar = dbar * sig[l - 1]
        q[l - 1] = BETA * dbar * delta[l - 1] * (1 - sig[l - 1]) * (z[l - 1] * BETA <= THRESH)
        if l == 4:
            vbar = torch.cat([abar / SQ2, gbar_e / SQ2], 1)
        else:
            vbar = abar
        dW[l] += delta[l].T @ vbar                  # v_l = W_l^T delta_l
        dbar = vbar @ W[l].T                        # delta-bar_l
    # delta_7 = W_8[0,:] * sigma_7
    dW[8][0, :] += (dbar * sig[7]).sum(0)
    q[7] = BETA * dbar * delta[7] * (1 - sig[7]) * (z[7] * BETA <= THRESH)
    # 3. ordinary backward l = 8..0
    zbar = torch.cat([sbar[:, None], fbar], 1)
    ebar = torch.zeros_like(st["e"])
    for l in range(8, -1, -1):
        dW[l] += zbar.T @ u[l]
        db[l] += zbar.sum(0)
        ubar = zbar @ W[l]
        if l == 4:
            ebar = ebar + ubar[:, -n_pe:] / SQ2
            hbar = ubar[:, :-n_pe] / SQ2
        else:
            hbar = ubar
        if l > 0:
            zbar = hbar * sig[l - 1] + q[l - 1]
        else:
            ebar = ebar + hbar
    xbar = xbar + pe_jt_apply(x, ebar, L)
    return xbar, dW, db


def color_forward(W, b, pts, dirs, normals, feat, Lv=4):
    inp = torch.cat([pts, pe(dirs, Lv), normals, feat], -1)
    acts = [inp]
    h = inp
    for l in range(len(W)):
        h = h @ W[l].T + b[l]
        if l < len(W) - 1:
            h = torch.relu(h)
            acts.append(h)
    rgb = torch.sigmoid(h)
    return rgb, dict(acts=acts, rgb=rgb)


def color_backward(W, dirs, st, rgbbar, Lv=4):
    """-> pts-bar, dirs-bar, normals-bar, feat-bar, dW, db"""
    acts, rgb = st["acts"], st["rgb"]
    zbar = rgbbar * rgb * (1 - rgb)
    dW = [None] * len(W)
    db = [None] * len(W)
    for l in range(len(W) - 1, -1, -1):
        dW[l] = zbar.T @ acts[l]
        db[l] = zbar.sum(0)
        abar = zbar @ W[l]
        if l > 0:
            zbar = abar * (acts[l] > 0)
    nd = 3 + 6 * Lv
    ptsbar = abar[:, 0:3]
    dirs_e_bar = abar[:, 3:3 + nd]
    nbar = abar[:, 3 + nd: 6 + nd]
    fbar = abar[:, 6 + nd:]
    dirsbar = pe_jt_apply(dirs, dirs_e_bar, Lv)
    return ptsbar, dirsbar, nbar, fbar, dW, db
