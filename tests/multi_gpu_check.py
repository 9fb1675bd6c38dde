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
    sharded = [p.grad.clone() for p in ts.all_params]
    loss_parts = torch.stack([ls["color_loss"].detach(), ls["mask_loss"].detach()])
    dist.all_reduce(loss_parts)
    ts1 = TrainStep(scene, mask_weight=5.0, group=None, optimizer=False)
    scene["renderer"].process_group = None
    ls1, _ = ts1.forward_backward(2, B, pixels=(px, py), t_rand=tr)
    worst = 0.0
    for a, p in zip(sharded, ts1.all_params):
        ref = p.grad if p.grad is not None else torch.zeros_like(p)      # pose MLPs of other frames get no gradient
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
    scene["renderer"].process_group = dist.group.WORLD
    ls_m, _ = ts.forward_backward(2, B // world, pixels=(px[sl], py[sl]), t_rand=tr[sl], micro_batch=B // world // 4)
    worst_m = 0.0
    for a, p in zip(sharded, ts.all_params):
        if a.norm().item() > 0:
            worst_m = max(worst_m, ((p.grad - a).norm() / a.norm()).item())
    print(f"rank {rank}: micro-batched sharded step vs one-shot sharded step: worst grad rel diff {worst_m:.2e}")
    ok = ok and worst_m < 5e-3
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
