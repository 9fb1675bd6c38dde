"""Weight-ring timing on CTA 0 (library built with -DFMOV_TRACE): for every ring stage use `it`
   lat   = stage seen full by the issuer - copy issued by the producer     (L2 -> shared-memory latency as the ring sees it)
   wait  = stage seen full - issuer started waiting for it                  (0 when the slice was already resident)
   turn  = copy of use it+STAGES issued - stage of use it seen full         (MMAs read the stage + commit + producer wake-up)
usage: FMOV_LIB=<traced .so> python profiles/trace_ring.py [rays]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from fmov_pose_b200 import _lib as L, fine, synthetic
from fmov_pose_b200.train import TrainStep
rays = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
STAGES = 4
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
orig_fwd = fine.fine_forward
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
for i in range(2):
    ts.forward_backward(i, rays, pixels=(px, py), t_rand=tr)
torch.cuda.synchronize()
t = traces["fwd"]
tag, clk = t[:, 0], t[:, 1]
kind, who, idx = tag >> 32, (tag >> 16) & 0xFFFF, tag & 0xFFFF
ev = {}
for k_, w_, i_, c_ in zip(kind, who, idx, clk):
    if k_ in (7, 8, 9):
        ev.setdefault((int(k_), int(i_)), int(c_))
its = sorted(i for (k_, i) in ev if k_ == 9)
lat, wait, turn = [], [], []
for i in its:
    if i < 64 or (7, i) not in ev or (8, i) not in ev:
        continue
    lat.append(ev[(9, i)] - ev[(7, i)])
    wait.append(ev[(9, i)] - ev[(8, i)])
    if (7, i + STAGES) in ev:
        turn.append(ev[(7, i + STAGES)] - ev[(9, i)])
for name, v in (("lat", lat), ("wait", wait), ("turn", turn)):
    v = np.array(v)
    print(f"{name:5s} n {len(v):5d} mean {v.mean():7.0f} p10 {np.percentile(v, 10):7.0f} p50 {np.percentile(v, 50):7.0f} p90 {np.percentile(v, 90):7.0f} max {v.max():7.0f}")
