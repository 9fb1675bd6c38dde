"""Small drivers for ncu captures (one GPU): python profiles/run_kernels.py query|step [rays]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fmov_pose_b200 import ops, packing, synthetic  # noqa: E402
from fmov_pose_b200.train import TrainStep  # noqa: E402

what = sys.argv[1]
rays = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
scene = synthetic.build_scene(device=dev, n_images=4, H=120, W=160)
if what == "query":
    W, b = scene["sdf_network"].effective_weights()
    qw = packing.SdfQueryWeights(W, b)
    pts = torch.rand(rays * 128, 3, device=dev) * 2 - 1
    for _ in range(3):
        out = ops.sdf_query_points(qw, pts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.sdf_query_points(qw, pts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"query {rays * 128} pts: {ms:.3f} ms  {rays * 128 / ms / 1e6:.3f} Gpts/s  "
          f"{rays * 128 * 1049088 / ms / 1e9:.1f} TFLOP/s (algorithmic)")
elif what in ("step", "hostcost"):
    ds = scene["dataset"]
    K = torch.tensor([[150.0, 0, 80.0], [0, 150.0, 60.0], [0, 0, 1.0]])
    ds.intrinsics_all_inv = torch.linalg.inv(K)[None].repeat(4, 1, 1).contiguous().to(dev)
    ts = TrainStep(scene, mask_weight=5.0)
    g = torch.Generator().manual_seed(0)
    for i in range(3):
        px = torch.randint(30, 130, [rays], generator=g).to(dev)
        py = torch.randint(10, 110, [rays], generator=g).to(dev)
        ts.step(i % 4, rays, pixels=(px, py), t_rand=torch.rand(rays, 1, generator=g).to(dev))
    torch.cuda.synchronize()
    print("step ok")
if what == "hostcost":
    import time
    torch.cuda.synchronize()
    for rr in (rays,):
        px = torch.randint(30, 130, [rr], generator=g).to(dev)
        py = torch.randint(10, 110, [rr], generator=g).to(dev)
        tr = torch.rand(rr, 1, generator=g).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(10):
            ts.step(i % 4, rr, pixels=(px, py), t_rand=tr)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"rays {rr}: host issue time {1e3 * (t1 - t0) / 10:.2f} ms/step, wall incl. GPU {1e3 * (t2 - t0) / 10:.2f} ms/step")
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(5):
        ts.step(i % 4, rays, pixels=(px, py), t_rand=tr)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
