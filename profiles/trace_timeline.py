"""Debug timeline of the chain engine (library built with -DFMOV_TRACE): per-step intervals on CTA 0.
usage: python profiles/trace_timeline.py [rays]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from fmov_pose_b200 import _lib as L, synthetic
from fmov_pose_b200.train import TrainStep
rays = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
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
from fmov_pose_b200 import fine
orig_fwd, orig_bwd = fine.fine_forward, fine.fine_backward
traces = {}
def wrap(name, fn):
    def f(*a, **k):
        lib.fmov_debug_trace(buf, NEV)
        r = fn(*a, **k)
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
    print(f"== {name}: {len(t)} events, span {clk.max()} cycles")
    # issuer: per (slot, step): wait-start(1), ready(2), issued(3)
    ev = {}
    for k_, w_, i_, c_ in zip(kind, who, idx, clk):
        ev.setdefault((int(k_), int(w_), int(i_)), int(c_))
    nwarp = int(who[kind == 4].max()) + 1 if (kind == 4).any() else 0
    per_slot = nwarp // 2
    rows = []
    for slot in (0, 1):
        steps = sorted(i for (k_, w_, i) in ev if k_ == 2 and w_ == slot)
        for st in steps[:70]:
            t_wait, t_rdy, t_iss = ev.get((1, slot, st)), ev.get((2, slot, st)), ev.get((3, slot, st))
            warps = range(slot * per_slot, (slot + 1) * per_slot)
            arr = [ev.get((6, w, st)) for w in warps if ev.get((6, w, st)) is not None]          # epilogue done (before this MMA)
            acc = [ev.get((5, w, st)) for w in warps if ev.get((5, w, st)) is not None]          # acc ready seen
            wst = [ev.get((4, w, st)) for w in warps if ev.get((4, w, st)) is not None]
            rows.append((slot, st, t_wait, t_rdy, t_iss, min(arr) if arr else None, max(arr) if arr else None,
                         min(acc) if acc else None, max(acc) if acc else None, min(wst) if wst else None))
    print("slot step | issuer: wait_start ready issued | epi(prev step) first_done last_done | acc_seen first last | skew  rdy-lastdone  issue  acc-issued")
    for r in rows:
        slot, st, tw, trd, ti, a0, a1, c0, c1, w0 = r
        f = lambda v: "      -" if v is None else f"{v:7d}"
        d = lambda a, b: "     -" if a is None or b is None else f"{a - b:6d}"
        print(f"{slot} {st:4d} | {f(tw)} {f(trd)} {f(ti)} | {f(a0)} {f(a1)} | {f(c0)} {f(c1)} | {d(a1, a0)} {d(trd, a1)} {d(ti, trd)} {d(c0, ti)}")
