"""Manual multi-GPU parity check (not collected by pytest; needs >= 2 GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py

Every rank renders its shard of ONE global ray batch with replicated weights; after the normaliser and gradient
all-reduces the parameter gradients must equal those of a single-process step on the whole batch."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fmov_pose_b200 import synthetic  # noqa: E402
from fmov_pose_b200.train import TrainStep  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    scene = synthetic.build_scene(device=dev, n_images=4, H=120, W=160)
    K = torch.tensor([[150.0, 0, 80.0], [0, 150.0, 60.0], [0, 0, 1.0]])
    scene["dataset"].intrinsics_all_inv = torch.linalg.inv(K)[None].repeat(4, 1, 1).contiguous().to(dev)
    B = 1024
    g = torch.Generator().manual_seed(7)
    px = torch.randint(30, 130, [B], generator=g).to(dev)
    py = torch.randint(10, 110, [B], generator=g).to(dev)
    tr = torch.rand(B, 1, generator=g).to(dev)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    ts = TrainStep(scene, mask_weight=5.0, group=dist.group.WORLD, optimizer=False)
    ls, _ = ts.forward_backward(2, B // world, pixels=(px[sl], py[sl]), t_rand=tr[sl])
    sharded = [None if p.grad is None else p.grad.clone() for p in ts.all_params]
    loss_parts = torch.stack([ls["color_loss"].detach(), ls["mask_loss"].detach()])
    dist.all_reduce(loss_parts)
    ts1 = TrainStep(scene, mask_weight=5.0, group=None, optimizer=False)
    ls1, _ = ts1.forward_backward(2, B, pixels=(px, py), t_rand=tr)
    worst = 0.0
    for a, p in zip(sharded, ts1.all_params):
        if p.grad is None:          # pose MLPs of frames nobody rendered: no gradient on one GPU, none after the all-reduce
            assert a is None, "a parameter without a gradient on one GPU came out of the sharded step with one"
            continue
        ref = p.grad
        assert a is not None
        if ref.norm().item() == 0.0:
            assert a.norm().item() == 0.0
            continue
        worst = max(worst, ((a - ref).norm() / ref.norm()).item())
    ok = worst < 2e-3 and abs(loss_parts[0].item() - ls1["color_loss"].item()) < 1e-5 and \
        abs(ls["eikonal_loss"].item() - ls1["eikonal_loss"].item()) < 1e-6
    print(f"rank {rank}: worst grad rel diff sharded-vs-single = {worst:.2e}; colour loss {loss_parts[0].item():.6f} vs "
          f"{ls1['color_loss'].item():.6f}; eikonal {ls['eikonal_loss'].item():.6f} vs {ls1['eikonal_loss'].item():.6f} -> "
          f"{'OK' if ok else 'MISMATCH'}")
    # config C5: grid partitioned into x-plane slabs across the ranks == the single-GPU grid (ragged: 50 planes / world)
    from fmov_pose_b200.grid import extract_fields_sharded
    rend = scene["renderer"]
    bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
    for res in (48, 50):
        u_sh = extract_fields_sharded(rend, bmin, bmax, res, group=dist.group.WORLD)
        u_1 = rend.extract_fields(bmin, bmax, res).view(res, res, res)
        same = torch.equal(u_sh, u_1)
        print(f"rank {rank}: sharded {res}^3 grid == single-GPU grid: {same}")
        ok = ok and same
    # micro-batched + ray-sharded step == one-shot sharded step
    ls_m, _ = ts.forward_backward(2, B // world, pixels=(px[sl], py[sl]), t_rand=tr[sl], micro_batch=B // world // 4)
    worst_m = 0.0
    for a, p in zip(sharded, ts.all_params):
        if a is not None and a.norm().item() > 0:
            worst_m = max(worst_m, ((p.grad - a).norm() / a.norm()).item())
    print(f"rank {rank}: micro-batched sharded step vs one-shot sharded step: worst grad rel diff {worst_m:.2e}")
    ok = ok and worst_m < 5e-3
    # Two optimiser steps with the ranks rendering frames of DIFFERENT pose MLPs (the advisor's case): FlatAdam all-reduces
    # gradients + group flags, so the union {networks, pose MLP of rank 0's frame, pose MLP of rank 1's frame} is stepped and
    # the pose MLPs of frames nobody rendered keep parameters, moments and step counters — equal to one process that renders
    # both frames in one batch (maintain_shape-style two-frame step) on the union.
    from fmov_pose_b200 import synthetic as syn
    Bh = 256
    frames = [(0, 1), (2, 1)]                     # (rank-0 frame, rank-1 frame) per step
    def fresh():
        sc = syn.build_scene(device=dev, n_images=4, n_samples=16, n_importance=16, up_sample_steps=2, pose_type="seg", H=120, W=160)
        sc["dataset"].intrinsics_all_inv = torch.linalg.inv(K)[None].repeat(4, 1, 1).contiguous().to(dev)
        return sc
    sc_a, sc_b = fresh(), fresh()
    ts_a = TrainStep(sc_a, mask_weight=5.0, group=dist.group.WORLD)
    ts_b = TrainStep(sc_b, mask_weight=5.0, group=None)
    for it, fr in enumerate(frames):
        sl0, sl1 = slice(0, Bh), slice(Bh, 2 * Bh)
        mine = sl0 if rank == 0 else sl1
        if world == 2:
            ts_a.step(fr[rank], Bh, pixels=(px[mine], py[mine]), t_rand=tr[mine])
        ts_b.step(fr[0], Bh, pixels=(px[sl0], py[sl0]), t_rand=tr[:2 * Bh], additional_img_id=fr[1], add_pixels=(px[sl1], py[sl1]))
    if world == 2:
        worst_o, mean_o, cnt = 0.0, 0.0, 0
        for a, b in zip(ts_a.all_params, ts_b.all_params):
            worst_o = max(worst_o, float((a - b).abs().max()))
            mean_o += float((a - b).abs().sum())
            cnt += a.numel()
        mean_o /= cnt
        steps_a = [float(v) for v in ts_a.optimizer.step_count]
        steps_b = [float(v) for v in ts_b.optimizer.step_count]
        print(f"rank {rank}: FlatAdam 2 steps, frames differ per rank: |param diff| vs single process max {worst_o:.2e} mean {mean_o:.2e}; "
              f"group step counters {steps_a} vs {steps_b}")
        # Adam's first steps are sign-like (lr per step whatever the gradient's size): where a gradient is fp16-rounding noise
        # around zero the two runs may step in opposite directions, so the bound is one step (lr = 5e-4) and a tiny mean
        ok = ok and worst_o < 1e-3 and mean_o < 2e-5 and steps_a == steps_b and steps_a[4] == 0.0      # pose MLP 3 was never rendered
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
