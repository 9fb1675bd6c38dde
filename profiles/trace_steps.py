"""Per-step timing of the chain engine on CTA 0 (library built with -DFMOV_TRACE), averaged over the CTA's tiles:
for every GEMM step of the fine_fwd / fine_bwd tile programs
   M    = accumulator seen by the first epilogue warp - last epilogue warp done with the previous step
          (operand hand-over + weight waits + MMAs + commit latency: the slot's epilogue warps idle)
   iss  = issuer: operand ready seen -> all MMAs of the step issued (includes waiting for weight slices / the other slot)
   E    = last epilogue warp done - accumulator seen (the step's epilogue)
usage: FMOV_LIB=<traced .so> python profiles/trace_steps.py [rays]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from fmov_pose_b200 import _lib as L, fine, synthetic
from fmov_pose_b200.train import TrainStep

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
NSTEP = 22
dev = torch.device("cuda:0")
scene = synthetic.build_scene(device=dev, n_images=4, H=120, W=160)
ds = scene["dataset"]
K = torch.tensor([[150.0, 0, 80.0], [0, 150.0, 60.0], [0, 0, 1.0]])
ds.intrinsics_all_inv = torch.linalg.inv(K)[None].repeat(4, 1, 1).contiguous().to(dev)
ts = TrainStep(scene, mask_weight=5.0, optimizer=False)
g = torch.Generator().manual_seed(0)
px = torch.randint(30, 130, [rays], generator=g).to(dev)
py = torch.randint(10, 110, [rays], generator=g).to(dev)
tr = torch.rand(rays, 1, generator=g).to(dev)
NEV = 24 * 16384
buf = (ctypes.c_longlong * (2 * NEV))()
lib = L.lib()
orig_fwd, orig_bwd = fine.fine_forward, fine.fine_backward
traces = {}


def wrap(name, fn):
    def f(*a, **k):
        lib.fmov_debug_trace(buf, NEV)
        r = fn(*a, **k)
        torch.cuda.synchronize()
        n = lib.fmov_debug_trace(buf, NEV)
        traces[name] = np.array(buf[: 2 * n], dtype=np.int64).reshape(n, 2).copy()
        return r
    return f


fine.fine_forward = wrap("fwd", orig_fwd)
fine.fine_backward = wrap("bwd", orig_bwd)
for i in range(2):
    ts.forward_backward(i, rays, pixels=(px, py), t_rand=tr)
torch.cuda.synchronize()
for name, t in traces.items():
    tag, clk = t[:, 0], t[:, 1]
    kind, who, idx = tag >> 32, (tag >> 16) & 0xFFFF, tag & 0xFFFF
    clk = clk - clk.min()
    ev = {}
    for k_, w_, i_, c_ in zip(kind, who, idx, clk):
        ev.setdefault((int(k_), int(w_), int(i_)), int(c_))
    nwarp = int(who[kind == 4].max()) + 1
    per_slot = nwarp // 2
    print(f"== {name}: {len(t)} events, span {clk.max()} cycles, {nwarp} epilogue warps")
    acc = {s: [[] for _ in range(4)] for s in range(NSTEP)}
    for slot in (0, 1):
        warps = range(slot * per_slot, (slot + 1) * per_slot)
        steps = sorted(i for (k_, w_, i) in ev if k_ == 2 and w_ == slot)
        for st in steps:
            if st < NSTEP:          # first tile of the slot: pipeline fill
                continue
            seen = [ev.get((5, w, st)) for w in warps]
            done_prev = [ev.get((6, w, st)) for w in warps]
            done = [ev.get((6, w, st + 1)) for w in warps]
            rdy, iss = ev.get((2, slot, st)), ev.get((3, slot, st))
            if None in seen or None in done_prev or rdy is None or iss is None:
                continue
            a = acc[st % NSTEP]
            a[0].append(min(seen) - max(done_prev))
            a[1].append(iss - rdy)
            a[3].append(min(seen) - iss)
            if None not in done:
                a[2].append(max(done) - min(seen))
    print("step     n      M    iss  acc-iss      E")
    tm = te = 0.0
    for s in range(NSTEP):
        a = acc[s]
        if not a[0]:
            continue
        m, i_, e, lag = np.mean(a[0]), np.mean(a[1]), (np.mean(a[2]) if a[2] else float("nan")), np.mean(a[3])
        tm += m
        te += 0 if np.isnan(e) else e
        print(f"{s:4d} {len(a[0]):5d} {m:6.0f} {i_:6.0f} {lag:8.0f} {e:6.0f}")
    print(f"sum over the tile program: M {tm:.0f}  E {te:.0f}  (cycles per tile and slot)")
