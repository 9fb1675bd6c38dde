"""bench.py — NeuS train-step throughput (train rays/s, fwd+bwd+Adam) of the B200-native path.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N>1) prints ONE JSON line.
  value        whole-job rays/s with step inputs already resident in HBM
  e2e          same metric through the public API with per-step pinned-host -> device copies of the step
               inputs (pixel indices, jitter) and a device -> host read of the loss inside the timed region
  roofline     dominant kernel of the step (CUDA-event timed live, algorithmic FLOPs / duration)
  cpu_baseline the reference algorithm (oracle port, PyTorch CPU) timed on this box's host cores (rank 0, N=1)
`--impl reference` times that CPU path alone with all host threads on a bounded sample of the workload.
Workload: configs[1] — ho3d_virtual.conf-shaped joint shape+pose train step (SegLearnPose per-frame pose,
BarfSDFNetwork 8x256, colour 4x256, eikonal+colour+mask loss, mask_weight 5) with the 64+64 sampling the
north_star throughput target is quoted on, on synthetic 640x480 frames.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic FLOPs
F_S = 1_049_088
F_C = 542_720


def flops_per_ray(n, m):
    s = n + m
    samp = (n + (m - m // 4 if m > 0 else 0)) * F_S if m > 0 else 0   # coarse + 3 of 4 rounds queried
    return samp + s * (6 * F_S + 3 * F_C)


KERNEL_FLOPS_PER_POINT = {       # algorithmic FLOPs per fine-stage sample point of each MLP kernel
    "fine_fwd": 2 * F_S + F_C,   # value pass + reverse sweep (normal) + colour forward
    "fine_bwd": 2 * F_S + F_C,   # adjoint pass + ordinary backward (dX) + colour backward (dX)
    "dw": 2 * F_S + F_C,         # two outer-product accumulations per SDF layer + colour
}


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(o[0]))
                self.max_mhz = float(o[1])
                for nm, v in zip(names, o[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_step_fn(B, n, m, steps_up, threads):
    """The reference algorithm on CPU (oracle port, PyTorch): render + loss + backward + Adam, C1 shapes."""
    import torch
    from oracle import neus_oracle as O
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.models.barf_fields import BarfRenderingNetwork, BarfSDFNetwork
    from fmov_pose_b200.models.fields import SingleVarianceNetwork
    torch.set_num_threads(threads)
    torch.manual_seed(2024)
    init = synthetic.make_init_poses(20)
    sdf = BarfSDFNetwork(init, n_images=20, **synthetic.SDF_KW)
    col = BarfRenderingNetwork(**synthetic.COL_KW)
    var = SingleVarianceNetwork(0.3)
    params = [p for net in (sdf, col, var) for p in net.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=5e-4)
    g = torch.Generator().manual_seed(1)
    intr_inv = torch.linalg.inv(torch.tensor(synthetic.INTRINSICS))

    def step():
        sdf_p = {k: v for k, v in sdf.named_parameters()}
        col_p = {k: v for k, v in col.named_parameters()}
        px = torch.randint(170, 470, [B], generator=g)
        py = torch.randint(90, 390, [B], generator=g)
        pose = init[3, :3].clone().requires_grad_(True)
        mask = (((px - 320) ** 2 + (py - 240) ** 2) < 150 ** 2).float()[:, None]
        losses, _ = O.train_step(sdf_p, col_p, var.variance, pose, intr_inv, px, py, torch.rand(B, 3, generator=g), mask,
                                 t_rand=torch.rand(B, 1, generator=g), n_samples=n, n_importance=m,
                                 up_sample_steps=steps_up, cos_anneal_ratio=1.0, igr_weight=0.1, mask_weight=5.0)
        opt.zero_grad()
        losses["loss"].backward()
        opt.step()
        return float(losses["loss"])
    return step


def time_torch_gpu(B, n, m, steps_up, dev, warm=2, iters=5):
    """The same PyTorch restatement of the reference (fp32 eager autograd, double backward through
    sdf_network.gradient) run on THIS GPU — the "reference-PyTorch on one B200" arm of north_star's 50x target.
    Baseline only: nothing of it is on the product path."""
    import torch
    from oracle import neus_oracle as O
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.models.barf_fields import BarfRenderingNetwork, BarfSDFNetwork
    from fmov_pose_b200.models.fields import SingleVarianceNetwork
    torch.manual_seed(2024)
    init = synthetic.make_init_poses(20)
    sdf = BarfSDFNetwork(init, n_images=20, **synthetic.SDF_KW).to(dev)
    col = BarfRenderingNetwork(**synthetic.COL_KW).to(dev)
    var = SingleVarianceNetwork(0.3).to(dev)
    params = [p for net in (sdf, col, var) for p in net.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=5e-4)
    g = torch.Generator(device=dev).manual_seed(1)
    intr_inv = torch.linalg.inv(torch.tensor(synthetic.INTRINSICS)).to(dev)
    init = init.to(dev)

    def step():
        with torch.device(dev):      # the oracle's factory calls (linspace/ones/zeros) follow the default device
            sdf_p = {k: v for k, v in sdf.named_parameters()}
            col_p = {k: v for k, v in col.named_parameters()}
            px = torch.randint(170, 470, [B], generator=g, device=dev)
            py = torch.randint(90, 390, [B], generator=g, device=dev)
            pose = init[3, :3].clone().requires_grad_(True)
            mask = (((px - 320) ** 2 + (py - 240) ** 2) < 150 ** 2).float()[:, None]
            losses, _ = O.train_step(sdf_p, col_p, var.variance, pose, intr_inv, px, py,
                                     torch.rand(B, 3, generator=g, device=dev), mask,
                                     t_rand=torch.rand(B, 1, generator=g, device=dev), n_samples=n, n_importance=m,
                                     up_sample_steps=steps_up, cos_anneal_ratio=1.0, igr_weight=0.1, mask_weight=5.0)
            opt.zero_grad()
            losses["loss"].backward()
            opt.step()
        return losses["loss"]

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return B / ms * 1e3, ms


def time_cpu(B, n, m, steps_up, warm, iters):
    """-> (rays/s, s/step, threads, kind): the reference's own sources (oracle/_ref, `kind: "reference"`) when the build
    container installed them (oracle/build_ref.py), else the oracle port."""
    threads = os.cpu_count() or 1
    from oracle import ref_arm
    if ref_arm.available():
        rps, med, threads = ref_arm.time_cpu(B, n, m, steps_up, warm, iters, threads)
        return rps, med, threads, "reference"
    fn = cpu_reference_step_fn(B, n, m, steps_up, threads)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return B / med, med, threads, "port"


def reference_gpu_arm(rays, n, m, two=False, timeout=600):
    """reference-PyTorch on THIS GPU — the unmodified reference sources (oracle/_ref) under
    torch.set_default_tensor_type("torch.cuda.FloatTensor") as exp_runner.py:2030 runs them — in a child process (the
    reference claims the module name `models` and changes the default tensor type).  north_star's 50x denominator."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_arm.py"), "gpu", str(rays), str(n), str(m)] +
                       (["two"] if two else []), capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        return {"error": (r.stderr or r.stdout)[-300:]}
    return json.loads(lines[-1])


# algorithmic HBM bytes per sample point of the MLP kernels (16 KiB blocks per 128-point tile x 128 B;
# fine_fwd 119, fine_bwd 218, dw 176 blocks)
KERNEL_BYTES_PER_POINT = {"fine_fwd": 119 * 128.0, "fine_bwd": 218 * 128.0, "dw": 176 * 128.0}
# algorithmic HBM bytes per RAY of the HBM-class kernels (fp32 row-major I/O, S samples per ray)
HBM_KERNEL_BYTES_PER_RAY = {
    "composite_fwd": lambda S: 4.0 * (S * (1 + 3 + 3 + 1) + S * (1 + 1 + 1 + 1 + 3) + 6 + 8),      # in: sdf, n, rgb, z; out: w, cdf, inside, mid_z, pts, per-ray
    "composite_bwd": lambda S: 4.0 * (S * (1 + 3 + 3 + 1) + S * (1 + 3 + 3 + 1 + 1) + 6 + 12),     # in: same; out: d_sdf, d_n, d_rgb, d_dist, d_mid, per-ray
    "sample_round": lambda S: 4.0 * (2 * S + 2 * S + 16 + 6),                                       # z, sdf read + written back, new samples
}


def measured_traffic(kernel, rays):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` summary of
    this very configuration (profiles/r2b_ncu_traffic.json, keyed by rays per launch); None when not captured"""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "r2b_ncu_traffic.json")))
        return float(tab[str(rays)][kernel])
    except Exception:
        return None


def measure_extras(scene, dev, use_graph=True):
    """Secondary numbers of BASELINE.json's metric (not the headline): dense SDF grid query rate (config C5 shape,
    one GPU's 1/8 slab of the 512^3 grid) and the literal ho3d_virtual.conf step (2 x 512 rays, 32+0 samples)."""
    import torch
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import TrainStep
    out = {}
    rend = scene["renderer"]
    res, count = 512, 512 ** 3 // 8
    buf = torch.empty(count, dtype=torch.float32, device=dev)
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    notes = {"act": "activation-split chain (default of extract_fields: activations and encoded inputs as fp16 hi + lo, SDF within "
                    "3.6e-4 of the reference's own grid on the whole box; 2 MMA passes per layer over one weight stream, one tile in flight)",
             True: "full split-precision chain (weights split too: 1.3e-5; 3 MMA passes per layer over two weight streams)",
             False: "plain fp16 chain (6.7e-4 on the +-1.01 box, two tiles in flight)"}
    for tag, precise in (("c5_grid_query", "act"), ("c5_grid_query_full_split", True), ("c5_grid_query_fast_fp16_chain", False)):
        for _ in range(2):
            rend.extract_fields(bmin, bmax, res, first=0, count=count, out=buf, precise=precise)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            rend.extract_fields(bmin, bmax, res, first=0, count=count, out=buf, precise=precise)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[tag] = {"points": count, "ms": ms, "sdf_queries_per_s": count / ms * 1e3,
                    "tflops_algorithmic": count * F_S / ms / 1e9,
                    "note": "1/8 slab of the 512^3 grid (config C5 per-GPU share), includes weight packing; " + notes[precise]}
    # whole-frame forward-only render (validate_image shape: 640x480 = 307,200 rays, 64+64 samples)
    ds = scene["dataset"]
    with torch.no_grad():
        pose = scene["pose_network"](0)[:3]
        rays_o, rays_d = ds.gen_rays_at(0, pose=pose)
        rend.render_image(rays_o, rays_d, chunk_rays=16384)
        torch.cuda.synchronize()
        e0.record()
        rend.render_image(rays_o, rays_d, chunk_rays=16384)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_rays = rays_o.shape[0] * rays_o.shape[1]
    out["frame_render_640x480"] = {"ms_per_frame": ms, "rays_per_s": n_rays / ms * 1e3,
                                   "note": "NeuSRenderer.render_image: forward-only, 16384-ray launches, 64+64 samples "
                                           "(the reference issues 600 sequential 512-ray render() calls per frame)"}
    torch.cuda.empty_cache()
    tg = {}
    from oracle import ref_arm
    if ref_arm.available():
        for tag, (rays_t, nn_, mm_, two) in {"512_rays_64+64": (512, 64, 64, False), "4096_rays_64+64": (4096, 64, 64, False),
                                             "1024_rays_32+0_two_frames": (1024, 32, 0, True)}.items():
            try:
                tg[tag] = reference_gpu_arm(rays_t, nn_, mm_, two)
            except Exception as e:
                tg[tag] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        note = ("the UNMODIFIED reference sources (oracle/_ref: NeuSRenderer.render, fields, SegLearnPose, "
                "Dataset.gen_random_rays_at) + loss block + backward + Adam under torch.set_default_tensor_type(cuda), as "
                "exp_runner.py:2030 runs them, on the same B200: north_star's 50x denominator")
    else:
        for rays_t in (512, 4096):      # the reference's own batch size, and a large batch that amortises its launches
            try:
                rps, ms_t = time_torch_gpu(rays_t, 64, 64, 4, dev)
                tg[f"{rays_t}_rays_64+64"] = {"rays_per_s": rps, "ms_per_step": ms_t}
            except Exception as e:      # e.g. out of memory at the large batch: report, do not fail the bench
                tg[f"{rays_t}_rays_64+64"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            torch.cuda.empty_cache()
        note = "oracle port of the reference (oracle/_ref absent) on the same B200"
    out["reference_pytorch_on_this_gpu"] = dict(tg, note=note)
    out["c2_literal_1024rays_32+0"] = measure_literal(dev, use_graph)
    out["c4_barf_gf_512rays_64+64"] = measure_c4(dev, use_graph)
    return out


def measure_c4(dev, use_graph=True):
    """BASELINE config C4: the confs/ho3d_barf.conf iteration — BarfSDFNetwork / BarfRenderingNetwork (the BARF embedder
    never applies its coarse-to-fine weights), ONE LearnPoseGF pose MLP for all frames (pose_type gf),
    512 rays, 64+64 samples, mask_weight 1 — as a replayed graph."""
    import torch
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import GraphedTrainStep, TrainStep
    sc = synthetic.build_scene(device=dev, n_samples=64, n_importance=64, up_sample_steps=4, pose_type="gf")
    ts = TrainStep(sc, mask_weight=1.0)
    B, n_it = 512, 24
    g = torch.Generator().manual_seed(5)
    px = torch.randint(140, 500, [n_it, B], generator=g).to(dev)
    py = torch.randint(60, 420, [n_it, B], generator=g).to(dev)
    tr = torch.rand(n_it, B, 1, generator=g).to(dev)
    gts = GraphedTrainStep(ts, B) if use_graph else None
    step = (lambda i: gts.step(i % 20, px[i], py[i], tr[i])) if use_graph else \
        (lambda i: ts.step(i % 20, B, pixels=(px[i], py[i]), t_rand=tr[i]))
    for i in range(8):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8, n_it):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (n_it - 8)
    return {"ms_per_step": ms, "rays_per_s": B / ms * 1e3,
            "note": "confs/ho3d_barf.conf iteration (pose_type gf, 512 rays, 64+64 samples, mask_weight 1); " +
                    ("CUDA-graph replay" if use_graph else "eager")}


def measure_literal(dev, use_graph=True):
    """the shipped confs/ho3d_virtual.conf iteration: 2 x 512 rays of two frames, 32+0 samples, both pose MLPs trained"""
    import torch
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import TrainStep
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sc2 = synthetic.build_scene(device=dev, n_samples=32, n_importance=0, up_sample_steps=4, pose_type="seg")
    ts2 = TrainStep(sc2, mask_weight=5.0, capturable=use_graph)
    g = torch.Generator().manual_seed(3)
    B2, Bf = 1024, 512          # maintain_shape: 512 rays of the current frame + 512 rays of an earlier frame
    n_it = 24
    px = torch.randint(140, 500, [n_it, B2], generator=g).to(dev)
    py = torch.randint(60, 420, [n_it, B2], generator=g).to(dev)
    tr = torch.rand(n_it, B2, 1, generator=g).to(dev)
    frames = [(1 + i % 3, i % 3) for i in range(n_it)]
    if use_graph:
        from fmov_pose_b200.train import GraphedTrainStep
        g2 = GraphedTrainStep(ts2, Bf, two_frames=True)
        step2 = lambda i: g2.step(frames[i][0], px[i, :Bf], py[i, :Bf], tr[i], add_img_id=frames[i][1],
                                  add_px=px[i, Bf:], add_py=py[i, Bf:])
    else:
        step2 = lambda i: ts2.step(frames[i][0], Bf, pixels=(px[i, :Bf], py[i, :Bf]), t_rand=tr[i],
                                   additional_img_id=frames[i][1], add_pixels=(px[i, Bf:], py[i, Bf:]))
    for i in range(8):
        step2(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(8, n_it):
        step2(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (n_it - 8)
    return {"ms_per_step": ms, "rays_per_s": B2 / ms * 1e3,
            "library_calls_per_step": getattr(g2, "launches_per_step", None) if use_graph else None,
            "note": "confs/ho3d_virtual.conf as shipped (n_samples 32, n_importance 0, maintain_shape: 512 rays of the "
                    "current frame + 512 of an earlier frame, two pose MLPs trained); " +
                    ("CUDA-graph replay" if use_graph else "eager, host-launch bound")}


def measure_marching_cubes_child():
    """`bench.py --mc_only` (a child process of the default run): validate_mesh's whole device path at config C5's size
    — the 512^3 u = -sdf grid from the network, then marching cubes on it — prints one JSON object.  HBM roofline of the
    three marching-cubes passes: each reads the grid once (4 B/point); the mesh itself is small against that."""
    import torch
    from fmov_pose_b200 import mcubes_gpu, synthetic
    dev = torch.device("cuda:0")
    scene = synthetic.build_scene(device=dev, n_images=2, H=48, W=64)
    rend = scene["renderer"]
    res = 512
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
    u = rend.extract_fields(bmin, bmax, res).reshape(res, res, res)
    h = 2.02 / (res - 1)
    for _ in range(2):
        v, t = mcubes_gpu.marching_cubes(u, 0.0, scale=(h, h, h), offset=(-1.01,) * 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        v, t = mcubes_gpu.marching_cubes(u, 0.0, scale=(h, h, h), offset=(-1.01,) * 3)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    alg = 3 * 4 * res ** 3 + int(v.numel()) * 4 + int(t.numel()) * 4
    r = torch.linalg.norm(v, dim=1)
    print(json.dumps({"grid": "512^3", "ms": ms, "vertices": int(v.shape[0]), "triangles": int(t.shape[0]),
                      "radius_min_max": [float(r.min()), float(r.max())],
                      "algorithmic_gb_per_s": alg / ms / 1e6, "hbm_frac": alg / ms / 1e6 / hbm,
                      "note": "fmov_mc_count + fmov_mc_scan + fmov_mc_vertices + fmov_mc_triangles incl. the host sync for the "
                              "output sizes; the reference copies the grid to the host and runs PyMCubes on one core"}))


def measure_in_child(flag):
    """runs in a child process so that code paths that have not run on hardware yet cannot take the bench's CUDA context down"""
    import subprocess
    r = subprocess.run([sys.executable, os.path.abspath(__file__), flag], capture_output=True, text=True, timeout=300)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        return {"error": (r.stderr or r.stdout)[-300:]}
    return json.loads(lines[-1])


def measure_grid_sharded(scene, dev, group, world):
    """Config C5 as stated: the 512^3 validate_mesh grid partitioned into x-plane slabs across the ranks (one launch per
    rank) + all-gather of the slabs over NCCL.  Collective: every rank calls it; device-timed, max over ranks."""
    import torch
    from fmov_pose_b200.grid import extract_fields_sharded
    res = 512
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
    ms_all = []
    for it in range(3):
        if group is not None:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        u = extract_fields_sharded(scene["renderer"], bmin, bmax, res, group=group)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if group is not None:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        ms_all.append(ms.item())
        del u
    ms = min(ms_all[1:])
    torch.cuda.empty_cache()
    return {"grid": "512^3", "n_gpus": world, "ms": ms, "sdf_queries_per_s": res ** 3 / ms * 1e3,
            "tflops_algorithmic": res ** 3 * F_S / ms / 1e9,
            "note": "x-plane slabs, one fmov_sdf_query_grid launch per rank + NCCL all-gather of the slabs; max over ranks"}


def measure_c3_micro(dev):
    """Config C3's 65,536-ray batch on ONE GPU through TrainStep(micro_batch=8192): whole-batch sampling and normalisers,
    fine stage + backward per 8192-ray micro-batch (the 27 GB stash is reused), one Adam step."""
    import torch
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import TrainStep
    sc = synthetic.build_scene(device=dev, n_samples=64, n_importance=64, up_sample_steps=4, pose_type="seg")
    ts = TrainStep(sc, mask_weight=5.0)
    B = 65536
    g = torch.Generator().manual_seed(7)
    px = torch.randint(140, 500, [B], generator=g).to(dev)
    py = torch.randint(60, 420, [B], generator=g).to(dev)
    tr = torch.rand(B, 1, generator=g).to(dev)
    ts.step(2, B, pixels=(px, py), t_rand=tr, micro_batch=8192)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(2):
        ts.step(3 + i, B, pixels=(px, py), t_rand=tr, micro_batch=8192)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    return {"ms_per_step": ms, "rays_per_s": B / ms * 1e3,
            "note": "65,536 rays x (64+64) on one GPU, 8 micro-batches of 8192 rays, eager launches"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=8192, help="rays per GPU per step (weak scaling)")
    ap.add_argument("--n_samples", type=int, default=64)
    ap.add_argument("--n_importance", type=int, default=64)
    ap.add_argument("--cpu_rays", type=int, default=512, help="rays per CPU-baseline step (config C1)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_extras", action="store_true")
    ap.add_argument("--no_graph", action="store_true", help="eager launches instead of CUDA-graph replay of the step")
    ap.add_argument("--mc_only", action="store_true", help="internal: marching-cubes extra, run as a child process")
    args = ap.parse_args()
    if args.mc_only:
        measure_marching_cubes_child()
        return 0
    # stdout carries exactly ONE JSON line: libraries that printf to fd 1 (NCCL prints its version banner there when
    # NCCL_DEBUG is set on the box) are sent to stderr, the line is written to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n, m, up = args.n_samples, args.n_importance, 4
    workload = (f"C2 ho3d_virtual-shaped joint shape+pose train step (SegLearnPose, 8x256 SDF + 4x256 colour, "
                f"eikonal+colour+mask loss, Adam), {n}+{m} samples, synthetic 640x480 frames")

    if args.impl == "reference":
        if rank != 0:
            return 0
        # CPU arm: the reference's models/picture_pose.py:6-8 switches the default tensor type to CUDA whenever a GPU is
        # visible, so the GPUs are hidden from this process before torch is imported
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
        warm, iters = max(1, min(args.warmup, 2)), max(1, min(args.steps, 8))
        rps, sec, threads, kind = time_cpu(args.cpu_rays, n, m, up, warm, iters)
        what = ("the UNMODIFIED reference sources (oracle/_ref: NeuSRenderer.render, SDF / colour / variance networks, "
                "SegLearnPose pose MLP) + the loss block and Adam calls of exp_runner.py:562-599, 772-816" if kind == "reference"
                else "oracle port of the reference PyTorch path (oracle/_ref was not built: /root/reference only exists "
                     "in the build container)")
        line = {"impl": "reference", "metric": "train rays/s (fwd+bwd+Adam)", "value": rps, "unit": "rays/s",
                "n_gpus": args.gpus, "steps": iters, "warmup": warm, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "rays_per_step": args.cpu_rays, "n_samples": n, "n_importance": m},
                "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": kind,
                                 "sample": f"median of {iters} steps ({warm} warm-up) of {args.cpu_rays} rays ({n}+{m}), "
                                           f"all host threads, GPUs hidden; {what}"},
                "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    from fmov_pose_b200 import _lib as L
    from fmov_pose_b200 import synthetic
    from fmov_pose_b200.train import GraphedTrainStep, TrainStep
    # the captured step contains its NCCL all-reduces (tests/multi_gpu_graph_check.py: graphed == eager on 2 GPUs)
    use_graph = not args.no_graph
    scene = synthetic.build_scene(device=dev, n_samples=n, n_importance=m, up_sample_steps=up, pose_type="seg")
    ts = TrainStep(scene, igr_weight=0.1, mask_weight=5.0, group=group, capturable=use_graph)
    B = args.rays
    gts = GraphedTrainStep(ts, B) if use_graph else None
    total_steps = args.warmup + args.steps
    g = torch.Generator().manual_seed(1234 + rank)
    # step inputs: pixel draw inside the mask bbox (mask_guided_sampling), jitter, frame id
    px_h = torch.randint(140, 500, [2 * total_steps, B], generator=g).pin_memory()
    py_h = torch.randint(60, 420, [2 * total_steps, B], generator=g).pin_memory()
    tr_h = torch.rand(2 * total_steps, B, 1, generator=g).pin_memory()
    img_ids = [(7 * i + rank) % scene["n_images"] for i in range(2 * total_steps)]
    px_d, py_d, tr_d = px_h.to(dev), py_h.to(dev), tr_h.to(dev)
    h2d = B * (8 + 8 + 4)
    loss_host = torch.zeros(2 * total_steps).pin_memory()     # per-step loss read-back (async D2H, like a logger)

    def barrier():
        if group is not None:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def run(i, e2e, eager=False):
        if gts is not None and not eager:
            # replay of the captured step; its static input buffers are refilled from pinned host memory (e2e) or
            # from device-resident tensors (value)
            src = (px_h[i], py_h[i], tr_h[i]) if e2e else (px_d[i], py_d[i], tr_d[i])
            ls, _ = gts.step(img_ids[i], *src)
        else:
            if e2e:
                px, py, tr = px_h[i].to(dev, non_blocking=True), py_h[i].to(dev, non_blocking=True), tr_h[i].to(dev, non_blocking=True)
            else:
                px, py, tr = px_d[i], py_d[i], tr_d[i]
            ls, _ = ts.step(img_ids[i], B, pixels=(px, py), t_rand=tr)
        if e2e:
            loss_host[i:i + 1].copy_(ls["loss"].detach().reshape(1), non_blocking=True)
        return ls

    def timed(e2e, first):
        for i in range(args.warmup):
            run(first + i, e2e)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            run(first + args.warmup + i, e2e)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if group is not None:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    # initialisation (not part of the W warm-up steps of the contract): first-use work that a long training run pays
    # once — CUDA module loading, the caching allocator growing to the 27 GB activation stash, NCCL channel setup
    for i in range(5):
        run(2 * total_steps - 1 - (i % 2), False, eager=True)
    if gts is not None:          # capture one graph per pose-parameter set (here: per frame) before anything is timed
        seen = set()
        for i in range(2 * total_steps):
            if ts.graph_key(img_ids[i]) not in seen:
                seen.add(ts.graph_key(img_ids[i]))
                run(i, False)
    barrier()
    # per-kernel CUDA-event profile (roofline of the dominant kernel): eager launches of the same step, outside the
    # timed regions (events cannot be recorded inside a graph replay)
    L.profile_reset(True)
    calls0, k0 = L.n_calls, L.kernel_launches()
    n_prof = max(3, min(args.steps, 10))
    for i in range(n_prof):
        run(i, False, eager=True)
    calls_per_step = (L.n_calls - calls0) // n_prof
    kernels_per_step = (L.kernel_launches() - k0) // n_prof      # counted inside libfmov_b200.so at every launch site
    prof = L.profile_summary()
    L.profile_reset(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev = timed(False, 0)
    ms_e2e = timed(True, total_steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    value = world * B * args.steps / (ms_dev * 1e-3)
    e2e_val = world * B * args.steps / (ms_e2e * 1e-3)
    # roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
    roof = None
    # kernels of libfmov_b200.so executed in the timed region: what one graph replay holds (counted by the library while the
    # step was captured; equal to the eager count) x steps
    if gts is not None and gts.kernels_per_step:
        kernels_per_step = gts.kernels_per_step
    launches = kernels_per_step * args.steps
    mlp = {k: v for k, v in prof.items() if k in KERNEL_FLOPS_PER_POINT}
    if mlp:
        top = max(mlp, key=lambda k: mlp[k]["ms"])
        avg_ms = mlp[top]["ms"] / mlp[top]["count"]
        fl = KERNEL_FLOPS_PER_POINT[top] * B * (n + m)
        ach = fl / (avg_ms * 1e-3) / 1e12
        step_ms = ms_dev / args.steps
        traffic = measured_traffic(top, B)
        roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic,
                "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of one launch at this batch size, ncu --set full "
                                 "(profiles/r2b_ncu_traffic.json)" if traffic else
                                 "no ncu --set full capture committed for this batch size (profiles/r2b_ncu_traffic.json)"),
                "peak_source": peak_src, "avg_launch_ms": avg_ms,
                "timing_note": "per-kernel times: CUDA events around each launch in an eager pass of the same step, outside "
                               "the timed graph replay (events cannot be recorded inside a replay)",
                "share_of_step": (mlp[top]["ms"] / n_prof) / step_ms,
                "kernel_ms_per_step": {k: v["ms"] / n_prof for k, v in prof.items()}}
        # the same kernel against the HBM roofline (the fine-stage kernels stream their activation stash through HBM)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_ach = KERNEL_BYTES_PER_POINT[top] * B * (n + m) / (avg_ms * 1e-3) / 1e9
        roof["hbm_view"] = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                            "algorithmic_bytes_per_point": KERNEL_BYTES_PER_POINT[top]}
        # every MLP kernel on both rooflines + the HBM-class kernels (north_star: "achieved HBM GB/s for sampling and
        # compositing"), all from the same live CUDA-event pass
        per = {}
        for k, v in mlp.items():
            t_ms = v["ms"] / v["count"]
            per[k] = {"ms": t_ms, "tflops": KERNEL_FLOPS_PER_POINT[k] * B * (n + m) / (t_ms * 1e-3) / 1e12,
                      "tensor_frac": KERNEL_FLOPS_PER_POINT[k] * B * (n + m) / (t_ms * 1e-3) / 1e12 / peak_tf,
                      "hbm_gbs": KERNEL_BYTES_PER_POINT[k] * B * (n + m) / (t_ms * 1e-3) / 1e9,
                      "hbm_frac": KERNEL_BYTES_PER_POINT[k] * B * (n + m) / (t_ms * 1e-3) / 1e9 / hbm_peak,
                      "dram_bytes_ncu": measured_traffic(k, B)}
        if "sdf_query" in prof:
            v = prof["sdf_query"]
            pts_q = B * (n + (m - m // up if m > 0 else 0))
            t_ms = v["ms"] / n_prof
            per["sdf_query (all sampling queries of a step)"] = {"ms": t_ms, "tflops": pts_q * F_S / (t_ms * 1e-3) / 1e12,
                                                                 "tensor_frac": pts_q * F_S / (t_ms * 1e-3) / 1e12 / peak_tf}
        for k, fn in HBM_KERNEL_BYTES_PER_RAY.items():
            if k in prof:
                v = prof[k]
                t_ms = v["ms"] / v["count"]
                by = fn(n + m) * B
                per[k] = {"ms": t_ms, "hbm_gbs": by / (t_ms * 1e-3) / 1e9, "hbm_frac": by / (t_ms * 1e-3) / 1e9 / hbm_peak,
                          "algorithmic_bytes": by, "dram_bytes_ncu": measured_traffic(k, B)}
        roof["kernels"] = per
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            extras = measure_extras(scene, dev, use_graph)
        except Exception as e:          # secondary numbers must never cost the headline line
            extras = {"error": f"measure_extras: {type(e).__name__}: {str(e)[:300]}"}
        try:
            extras["c3_65536rays_micro_batched"] = measure_c3_micro(dev)
        except Exception as e:          # secondary number: report, do not fail the bench
            extras["c3_65536rays_micro_batched"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        try:
            extras["c5_marching_cubes_512"] = measure_in_child("--mc_only")
        except Exception as e:
            extras["c5_marching_cubes_512"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    if world > 1 and not args.no_extras:          # collective sections: every rank takes part
        extras_mg = {}
        # (1) attribution of the N-GPU step: the SAME step without any collective (group=None), all ranks running it at the
        # same time.  max over ranks = what the slowest GPU needs on its own under the node's concurrent load (power cap,
        # clock spread); N-GPU step time minus that = cost of the two collectives + the waiting they impose.
        try:
            sc_solo = synthetic.build_scene(device=dev, n_samples=n, n_importance=m, up_sample_steps=up, pose_type="seg")
            ts_solo = TrainStep(sc_solo, igr_weight=0.1, mask_weight=5.0, group=None, capturable=use_graph)
            g_solo = GraphedTrainStep(ts_solo, B)
            for i in range(2 * total_steps):
                if ts_solo.graph_key(img_ids[i]) not in {k[0] for k in g_solo.graphs}:
                    g_solo.step(img_ids[i], px_d[i], py_d[i], tr_d[i])
            for i in range(args.warmup):
                g_solo.step(img_ids[i], px_d[i], py_d[i], tr_d[i])
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                g_solo.step(img_ids[args.warmup + i], px_d[args.warmup + i], py_d[args.warmup + i], tr_d[args.warmup + i])
            e1.record()
            torch.cuda.synchronize()
            solo = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
            allsolo = [torch.zeros_like(solo) for _ in range(world)]
            torch.distributed.all_gather(allsolo, solo)
            solo_ms = [float(t.item()) for t in allsolo]
            step_ms_n = ms_dev / args.steps
            extras_mg["step_attribution"] = {
                "n_gpu_step_ms": step_ms_n, "solo_step_ms_per_rank": solo_ms, "solo_max_ms": max(solo_ms),
                "solo_min_ms": min(solo_ms), "collectives_and_waiting_ms": step_ms_n - max(solo_ms),
                "note": "solo = the identical captured step with group=None (no all-reduce), all ranks running it concurrently; "
                        "the N-GPU step cannot be faster than the slowest rank's solo step, the remainder is the two "
                        "collectives (4-float normaliser pack inside the forward, flat gradient buffer after dw) and the "
                        "waiting they impose"}
            g_solo.graphs.clear()
            del g_solo, ts_solo, sc_solo
        except Exception as e:
            extras_mg["step_attribution"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        # (2) config C3 as BASELINE.json states it: a 65,536-ray GLOBAL batch sharded over the N GPUs (32 K / 16 K rays per
        # GPU at N = 2 / 4), each rank running its shard as micro-batches of 8192 rays inside ONE captured graph
        try:
            per = 65536 // world
            if per > B:
                g3 = GraphedTrainStep(ts, per, micro_batch=B)
                gg = torch.Generator().manual_seed(99 + rank)
                px3 = torch.randint(140, 500, [per], generator=gg).to(dev)
                py3 = torch.randint(60, 420, [per], generator=gg).to(dev)
                tr3 = torch.rand(per, 1, generator=gg).to(dev)
                for i in range(2):
                    g3.step(3, px3, py3, tr3)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(4):
                    g3.step(3, px3, py3, tr3)
                e1.record()
                barrier()
                ms3 = torch.tensor([e0.elapsed_time(e1) / 4], device=dev)
                torch.distributed.all_reduce(ms3, op=torch.distributed.ReduceOp.MAX)
                extras_mg["c3_65536_global_rays"] = {
                    "ms_per_step": ms3.item(), "rays_per_s": 65536 / ms3.item() * 1e3, "rays_per_gpu": per,
                    "micro_batches_per_gpu": per // B,
                    "note": "config C3 as stated (strong-scaled 64 K-ray global batch): graph replay of the micro-batched "
                            "step, one normaliser all-reduce + one gradient all-reduce per step; max over ranks"}
                g3.graphs.clear()
                del g3
            else:
                extras_mg["c3_65536_global_rays"] = {"note": f"at N = {world} config C3 is {per} rays per GPU = the headline line"}
        except Exception as e:
            extras_mg["c3_65536_global_rays"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        if rank == 0:
            extras = dict(extras or {}, **extras_mg)
    if not args.no_extras:              # collective (all ranks): config C5 across the N GPUs of this run
        try:
            grid = measure_grid_sharded(scene, dev, group, world)
        except Exception as e:
            grid = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        if rank == 0:
            extras = extras or {}
            extras["c5_grid_512_sharded"] = grid
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # ~10-20 s of host work on the box's cores, in a child process with the GPUs hidden (see --impl reference)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "8", "--warmup", "2",
                                "--cpu_rays", str(args.cpu_rays), "--n_samples", str(n), "--n_importance", str(m)],
                               capture_output=True, text=True, timeout=900, cwd=ROOT)
            cpu = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])["cpu_baseline"]
        except Exception as e:
            cpu = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    if rank == 0:
        line = {"metric": "train rays/s (fwd+bwd+Adam)", "value": value, "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16/bf16 operands, f32 accumulate",
                "data": "synthetic",
                "config": {"workload": workload, "rays_per_gpu": B, "global_rays": world * B, "n_samples": n,
                           "n_importance": m, "up_sample_steps": up, "parallelism": f"ray-sharded dp{world}",
                           "launch": "CUDA-graph replay of the whole step" if use_graph else "eager",
                           "l2": "per-step working set (activation/gradient stash ~26 KB/sample) >> 126 MB L2",
                           "algorithmic_flops_per_ray": flops_per_ray(n, m)},
                "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "gpu_launches_note": f"{kernels_per_step} kernels of libfmov_b200.so per step "
                f"(fmov_launch_count: counted inside the library at every launch site; {calls_per_step} C-ABI calls), "
                "torch glue kernels not included", "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu,
                "extras": extras,
                "tensor_frac_whole_step": flops_per_ray(n, m) * value / world / 1e12 / peak_tf}
        emit(line)
    if group is not None:
        # teardown: graphs that hold NCCL kernels must be gone before the communicator is destroyed; measured in round 1:
        # destroy_process_group() with live captured graphs never returns (the run had already printed its line), so
        # the ranks leave through a barrier + hard exit instead
        if gts is not None:
            gts.graphs.clear()
        torch.cuda.synchronize()
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        if use_graph:
            os._exit(0)
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
