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
if what == "flow":
    # HBM-bound warp-per-ray kernels of csrc/flow.cu: algorithmic bytes per ray at S samples =
    #   fwd  z + weights (8 S) + o, d, xy (32) + err (8)                        = 8 S + 40
    #   bwd  z + weights (8 S) + o, d, xy, g_err (40) + d_weights (4 S) + d_o, d_d (24) = 12 S + 64   (d_z: + 4 S)
    from fmov_pose_b200 import flow
    S = 128
    B = rays
    g = torch.Generator(device=dev).manual_seed(0)
    o = torch.tensor([0.0, 0.0, -3.0], device=dev).repeat(B, 1).requires_grad_(True)
    d = torch.nn.functional.normalize(torch.randn(B, 3, device=dev, generator=g) * 0.1 + torch.tensor([0, 0, 1.0], device=dev), dim=-1).requires_grad_(True)
    z = (2.0 + torch.sort(torch.rand(B, S, device=dev, generator=g), dim=-1)[0] * 2.0)
    w = torch.rand(B, S, device=dev, generator=g).requires_grad_(True)
    c2w = torch.eye(4, device=dev)[:3].clone()
    c2w[:, 3] = torch.tensor([0.1, 0.0, -3.0], device=dev)
    c2w.requires_grad_(True)
    K = torch.tensor([[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1.0]], device=dev)
    xy = torch.rand(B, 2, device=dev, generator=g) * 400
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for it in range(3):
        w2c = flow.w2c_from_c2w(c2w)
        ev[0].record()
        err = flow.FlowReprojFunction.apply(o, d, z, w, w2c, K, xy, 2.0 / 64)
        ev[1].record()
        loss = err.abs().sum()
        gerr = torch.autograd.grad(loss, [err], retain_graph=True)[0]
        ev[2].record()
        grads = torch.autograd.grad(err, [o, d, w, w2c], gerr)
        ev[3].record()
    torch.cuda.synchronize()
    f_ms, b_ms = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
    fb, bb = B * (8 * S + 40), B * (12 * S + 64)
    print(f"flow_fwd {B} rays x {S}: {f_ms * 1e3:.1f} us  {fb / f_ms / 1e6:.0f} GB/s algorithmic;  "
          f"flow_bwd: {b_ms * 1e3:.1f} us  {bb / b_ms / 1e6:.0f} GB/s algorithmic")
    sys.exit(0)
scene = synthetic.build_scene(device=dev, n_images=4, H=120, W=160)
if what == "grid":
    # config C5 per-GPU slab through the value chains: plain fp16, activation-split (extract_fields' default), full split
    rend = scene["renderer"]
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
    count = 512 ** 3 // 64
    for precise in (False, "act", True):
        rend.extract_fields(bmin, bmax, 512, first=0, count=count, precise=precise)
    torch.cuda.synchronize()
    print("grid ok")
elif what == "query":
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
