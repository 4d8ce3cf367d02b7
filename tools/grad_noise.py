"""Diagnostic (GPU): per-parameter gradient error of the CUDA path vs the fp64 oracle, next to
the fp32 oracle's own error (the noise floor of SURVEY 0.5).  python tools/grad_noise.py"""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hg_oracle as O, synth                      # noqa: E402
from pose_adv_aug_b200.models import asn_stacked_hg as M      # noqa: E402

S, Mo, K, C, N, R = 2, 1, 16, 64, 4, 128
dev = torch.device("cuda", 0)
sd64 = synth.make_state_dict(O.hg_schema(S, Mo, K, C), seed=31, dtype=torch.float64)
x64 = synth.make_images(N, R, seed=32, dtype=torch.float64)
t64 = synth.make_heatmaps(N, R, K, seed=33, dtype=torch.float64)
outs64, loss64, g64, _ = O.train_step(sd64, x64, t64, S, Mo)
sd32 = OrderedDict((k, v.float() if v.is_floating_point() else v) for k, v in sd64.items())
outs32, loss32, g32, _ = O.train_step(sd32, x64.float(), t64.float(), S, Mo)
net = M.create_hg(S, Mo, K, C)
net.load_state_dict(sd32)
net.to(dev).train()
outs = net(x64.float().to(dev))
loss = O.mse_loss(outs, t64.float().to(dev))
loss.backward()
params = dict(net.named_parameters())
print("heatmap err ours/ref32 vs f64:", [float((a.cpu().double() - b).abs().max()) for a, b in zip(outs, outs64)],
      [float((a.double() - b).abs().max()) for a, b in zip(outs32, outs64)])
rows = []
for k in g64:
    if k.endswith("bias") and "bn" not in k and "linear.0.1" not in k and "linear.1.1" not in k and "out_conv" not in k:
        continue
    d = float(g64[k].norm())
    if d == 0:
        continue
    e_ours = float((params[k].grad.cpu().double() - g64[k]).norm()) / d
    e_ref = float((g32[k].double() - g64[k]).norm()) / d
    rows.append((k, e_ours, e_ref))
for k, a, b in rows:
    if ".conv" in k or "bn3.weight" in k or k.startswith("conv1") or "out_conv" in k or "linear" in k:
        print("%-34s ours %.2e  ref32 %.2e  ratio %.1f" % (k, a, b, a / max(b, 1e-30)))
