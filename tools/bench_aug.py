"""Measurement of the GPU crop / rotate / resize augmentation (pylib/HumanAug.py::crop_batch, csrc/warp.cu) next to the
reference's CPU path (oracle/aug_oracle.py::crop = the reference's crop over PIL, one process as inside a DataLoader worker):
a batch of 24 MPII-sized photographs (720 x 1280) with agent-style sampled scales / rotations.
  python tools/bench_aug.py [--batch 24] [--reps 20] [--cpu-baseline]
The GPU arm imports nothing from oracle/; the --cpu-baseline leg is the one place this tool executes the checker."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pose_adv_aug_b200 import synth                          # noqa: E402
from pose_adv_aug_b200.pylib import HumanAug as H            # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=24)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--cpu-baseline", action="store_true",
                help="also time the reference's CPU path (oracle/aug_oracle.py, the checker) on the same batch and compare the bytes")
args = ap.parse_args()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
photos = [synth.make_photo(720, 1280, 300 + k) for k in range(args.batch)]
centers = np.stack([[rng.uniform(300, 980), rng.uniform(200, 520)] for _ in range(args.batch)]).astype(np.float32)
# scale_provided * 1.25 * 2^N(0, 0.25 clipped), rotation from the agent's 7 Gaussians (ref data/joint_train_s_r_agent.py:30-35)
scales = (rng.uniform(0.8, 3.2, args.batch) * 1.25 * 2 ** rng.choice(np.arange(-0.6, 0.61, 0.2), args.batch)).astype(np.float32)
rots = rng.choice(np.arange(-60, 61, 20), args.batch) + rng.normal(0, 5, args.batch).clip(-5, 5)
imgs = [torch.from_numpy(p).to(dev) for p in photos]
for _ in range(3):
    out = H.crop_batch(imgs, centers, scales, rots, 256, 200)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for _ in range(args.reps):
    out = H.crop_batch(imgs, centers, scales, rots, 256, 200)
e1.record()
torch.cuda.synchronize()
wall = (time.time() - t0) / args.reps
gpu_ms = e0.elapsed_time(e1) / args.reps
for _ in range(2):
    H.crop_batch(imgs, centers, scales, rots, 256, 200, batched=False)
torch.cuda.synchronize()
t1 = time.time()
for _ in range(max(args.reps // 4, 2)):
    H.crop_batch(imgs, centers, scales, rots, 256, 200, batched=False)
torch.cuda.synchronize()
wall_per_image_path = (time.time() - t1) / max(args.reps // 4, 2)
line = {"what": "HumanAug.crop_batch: %d photographs 720x1280 -> [N,3,256,256] float32, sampled scale / rotation" % args.batch,
        "gpu_ms_per_batch": gpu_ms, "per_image_launch_path_ms_per_batch": wall_per_image_path * 1e3, "wall_ms_per_batch": wall * 1e3, "images_per_s": args.batch / wall,
        "n_shrunk_first": int((scales * 200 / 256 >= 2).sum()), "n_rotated": int((rots != 0).sum())}
if args.cpu_baseline:
    from oracle import aug_oracle as A
    t0 = time.time()
    ref = [A.im_to_torch_float(A.crop(p, c, s, r, 256, 200)) for p, c, s, r in zip(photos, centers, scales, rots)]
    cpu = time.time() - t0
    line["cpu_reference_ms_per_batch"] = cpu * 1e3
    line["cpu_images_per_s_one_core"] = args.batch / cpu
    line["bit_exact_vs_cpu"] = bool(np.array_equal(np.stack(ref), out.cpu().numpy()))
print(json.dumps(line))
