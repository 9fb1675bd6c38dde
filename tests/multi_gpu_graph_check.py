"""Manual check (needs >= 2 GPUs, not collected by pytest): CUDA-graph replay of the ray-sharded train step with the
NCCL all-reduces captured inside the graph.  Round-1 result on 2 x B200: graphed losses == eager losses on both ranks.
destroy_process_group() with live captured graphs never returns, so the script leaves through os._exit after a barrier.

    timeout 120 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/multi_gpu_graph_check.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fmov_pose_b200 import synthetic  # noqa: E402
from fmov_pose_b200.train import GraphedTrainStep, TrainStep  # noqa: E402


def say(rank, msg):
    print(f"[{time.strftime('%H:%M:%S')}] rank {rank}: {msg}", flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, n_steps = 256, 4
    g = torch.Generator().manual_seed(7 + rank)
    px = torch.randint(200, 440, [n_steps, B], generator=g).to(dev)
    py = torch.randint(120, 360, [n_steps, B], generator=g).to(dev)
    tr = torch.rand(n_steps, B, 1, generator=g).to(dev)
    imgs = [1, 1, 2, 1]
    res = []
    for graphed in (False, True):
        scene = synthetic.build_scene(device=dev, n_images=4, n_samples=16, n_importance=16, up_sample_steps=2)
        ts = TrainStep(scene, mask_weight=5.0, group=dist.group.WORLD, capturable=graphed)
        gts = GraphedTrainStep(ts, B) if graphed else None
        losses = []
        for i in range(n_steps):
            say(rank, f"graphed={graphed} step {i}")
            if graphed:
                ls, _ = gts.step(imgs[i], px[i], py[i], tr[i])
            else:
                ls, _ = ts.step(imgs[i], B, pixels=(px[i], py[i]), t_rand=tr[i])
            torch.cuda.synchronize()
            losses.append(float(ls["loss"].detach()))
        res.append(losses)
    ok = all(abs(a - b) <= 2e-3 * max(1.0, abs(a)) for a, b in zip(*res))
    say(rank, f"eager {res[0]} graphed {res[1]} -> {'OK' if ok else 'MISMATCH'}")
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
