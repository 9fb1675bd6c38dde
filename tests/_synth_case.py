"""A tiny synthetic HO3D-style case on disk for running the reference's exp_runner.py (test infrastructure): images,
object masks, cameras_sphere.npz, crop transforms and an (empty) LoFTR match directory, in the layout
models/dataset.py:146-545 loads, plus a test copy of confs/ho3d_virtual.conf with the schedule shortened."""
import os
import re

import numpy as np

H, W, N_IMAGES, FOCAL = 48, 64, 4, 60.0


def write_case(work, case="SYN_ori"):
    import cv2 as cv
    data = os.path.join(work, "data", "HO3Dv3", case)
    os.makedirs(os.path.join(data, "image"))
    os.makedirs(os.path.join(data, "mask_obj"))
    os.makedirs(os.path.join(work, "data", "HO3Dv3", "matches", case.split("_")[0]))
    rng = np.random.RandomState(7)
    ys, xs = np.mgrid[0:H, 0:W]
    K = np.array([[FOCAL, 0, W / 2], [0, FOCAL, H / 2], [0, 0, 1.0]])
    cams, crops = {}, {}
    for i in range(N_IMAGES):
        name = f"{i:04d}"
        disc = (((xs - W / 2) ** 2 + (ys - H / 2) ** 2) < 15 ** 2)
        img = (rng.rand(H, W, 3) * 255).astype(np.uint8)
        img[~disc] = 0
        cv.imwrite(os.path.join(data, "image", name + ".png"), img)
        cv.imwrite(os.path.join(data, "mask_obj", name + ".png"), (disc[..., None] * np.ones(3) * 255).astype(np.uint8))
        # world-to-camera [R|t] of a camera at distance 3.6 looking at the origin, rotated a little per frame
        a = 0.15 * i
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        t = np.array([0.0, 0.0, 3.6])
        P = np.eye(4)
        P[:3, :4] = K @ np.concatenate([R, t[:, None]], axis=1)
        cams[f"world_mat_{name}"] = P.astype(np.float32)
        cams[f"scale_mat_{name}"] = np.eye(4, dtype=np.float32)
        crops[name] = np.eye(3, dtype=np.float32)
    np.savez(os.path.join(data, "cameras_sphere.npz"), **cams)
    np.save(os.path.join(data, "transform_matrixs.npy"), crops, allow_pickle=True)
    return data


def write_conf(ref_root, work, end_iter=6, batch_size=128, name="ho3d_virtual.conf"):
    """the reference's own conf with the schedule shortened (textual edits of values only)"""
    text = open(os.path.join(ref_root, "confs", name)).read()

    def sub(key, val):
        nonlocal text
        text, n = re.subn(rf"(\n\s*{key}\s*=\s*)[^\n]+", rf"\g<1>{val}", text, count=1)
        assert n == 1 or key in optional, key
    optional = ("max_pro_iteration", "pro_warm_up_end", "mesh_warmup_step")          # absent from the non-progressive confs
    for key, val in dict(end_iter=end_iter, batch_size=batch_size, save_freq=end_iter, val_freq=100000, val_mesh_freq=100000,
                         report_freq=1, pose_freq=100000, mesh_warmup_step=0, max_pro_iteration=2, pro_warm_up_end=1,
                         warm_up_end=2).items():
        sub(key, val)
    text, n = re.subn(r"recording\s*=\s*\[[^\]]*\]", "recording = [ ./ ]", text, count=1)
    assert n == 1
    path = os.path.join(work, "test_" + name)
    with open(path, "w") as fh:
        fh.write(text)
    return path
