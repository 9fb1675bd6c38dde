"""Accuracy and throughput of the three SDF grid-query chains on the reference's own 40^3 golden grid (tests/golden/grid40.npz,
bbox +-1.01) and on a 16.8 M-point slab of the 512^3 grid (config C5's per-GPU share): plain fp16 chain, activation-split
chain, full split-precision chain.   usage: python profiles/grid_modes.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from fmov_pose_b200 import ops, packing
from oracle import neus_oracle as O
from tests._util import load_golden, params_from

dev = torch.device("cuda:0")
d = load_golden("grid40")
p = params_from(d, "sdf.")
W = [O.eff_weight(p, "", l).to(dev) for l in range(9)]
b = [p[f"lin{l}.bias"].to(dev) for l in range(9)]
qw = packing.SdfQueryWeights(W, b, precise=True)
res = int(d["res"])
g = torch.Generator().manual_seed(5)
pts = (torch.rand(200000, 3, generator=g) * 2 - 1) * 1.2
ref_pts = O.sdf_value(p, pts)
count = 512 ** 3 // 8
buf = torch.empty(count, device=dev)
for name, mode in (("plain fp16 chain", False), ("activation-split chain", "act"), ("full split-precision chain", True)):
    out = torch.empty(res ** 3, device=dev)
    ops.sdf_query_grid(qw, [-1.01] * 3, [1.01] * 3, res, 0, res ** 3, out, precise=mode)
    err = np.abs(out.cpu().reshape(res, res, res).numpy() - d["u"]).max()
    got = ops.sdf_query_points(qw, pts.to(dev), precise=mode)
    err_pts = (got.cpu() - ref_pts).abs().max().item()
    for _ in range(2):
        ops.sdf_query_grid(qw, [-1.01] * 3, [1.01] * 3, 512, 0, count, buf, precise=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.sdf_query_grid(qw, [-1.01] * 3, [1.01] * 3, 512, 0, count, buf, precise=mode)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name:28s}: max |err| vs the reference's 40^3 grid {err:.2e}, vs fp32 oracle on 200 K points with |x| <= 2.1 {err_pts:.2e}; "
          f"{count} points in {ms:.2f} ms = {count / ms / 1e6:.3f} G queries/s")
