"""Kernel launches of one eager train iteration (torch.profiler): python profiles/count_launches.py
literal shipped config (2 x 512 rays of two frames, 32+0) and the 8192-ray 64+64 step."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from fmov_pose_b200 import synthetic  # noqa: E402
from fmov_pose_b200.train import TrainStep  # noqa: E402

dev = torch.device("cuda:0")
for name, kw, B, two in (("literal 2x512 rays 32+0", dict(n_samples=32, n_importance=0), 512, True),
                         ("8192 rays 64+64", dict(n_samples=64, n_importance=64), 8192, False)):
    sc = synthetic.build_scene(device=dev, pose_type="seg", **kw)
    ts = TrainStep(sc, mask_weight=5.0, fused_loss="--torch_loss" not in sys.argv,
                   fused_rays=None if "--torch_rays" not in sys.argv else False)
    g = torch.Generator().manual_seed(0)
    px = torch.randint(140, 500, [2 * B], generator=g).to(dev)
    py = torch.randint(60, 420, [2 * B], generator=g).to(dev)
    tr = torch.rand(2 * B if two else B, 1, generator=g).to(dev)
    add = dict(additional_img_id=1, add_pixels=(px[B:], py[B:])) if two else {}
    for _ in range(3):
        ts.step(2, B, pixels=(px[:B], py[:B]), t_rand=tr, **add)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ts.step(2, B, pixels=(px[:B], py[:B]), t_rand=tr, **add)
        torch.cuda.synchronize()
    cnt = collections.Counter()
    names = collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in ev.name.lower() and "memset" not in ev.name.lower():
            cnt["fmov" if "fmov::" in ev.name else "torch/other"] += 1
            names[ev.name[:110]] += 1
    print(f"{name}: {sum(cnt.values())} kernel launches per iteration ({dict(cnt)})")
    if "--names" in sys.argv:          # which torch glue is left to fuse
        for k, v in names.most_common(40):
            print(f"    {v:4d} x {k}")
    del ts, sc
    torch.cuda.empty_cache()
