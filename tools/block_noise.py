"""Diagnostic (GPU): gradient error of ONE _Residual block vs fp64, ours vs the fp32 oracle."""
import os, sys
from collections import OrderedDict
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import hg_oracle as O, synth
from pose_adv_aug_b200.models import asn_stacked_hg as M
dev = torch.device("cuda", 0)
for (cin, cout, N, H, seed) in [(64, 64, 4, 32, 5), (256, 256, 4, 16, 6), (64, 64, 4, 32, 7)]:
    schema = O.residual_schema("r", cin, cout, False)
    sd64 = synth.make_state_dict(schema, seed=seed, dtype=torch.float64)
    x64 = synth.make_tensor("x", (N, cin, H, H), seed=seed, lo=-1, hi=2, dtype=torch.float64)
    gy = synth.make_tensor("gy", (N, cout, H, H), seed=seed, dtype=torch.float64)
    res = {}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        leaves = OrderedDict((k, v.to(dt).clone().requires_grad_(True)) for k, v in sd64.items() if O.is_trainable(k))
        work = OrderedDict((k, leaves.get(k, v.to(dt) if v.is_floating_point() else v)) for k, v in sd64.items())
        xr = x64.to(dt).clone().requires_grad_(True)
        yr = O.residual(work, "r", xr, O.BNState(True))
        yr.backward(gy.to(dt))
        res[tag] = dict((k, v.grad.double()) for k, v in leaves.items())
        res[tag]["x"] = xr.grad.double()
        res[tag]["y"] = yr.detach().double()
    blk = M._Residual(cin, cout)
    blk.load_state_dict(OrderedDict((k[2:], v.float()) for k, v in sd64.items()))
    blk.to(dev).train()
    x = x64.float().to(dev).requires_grad_(True)
    y = blk(x)
    y.backward(gy.float().to(dev))
    ours = dict(("r." + k, p.grad.cpu().double()) for k, p in blk.named_parameters())
    ours["x"] = x.grad.cpu().double(); ours["y"] = y.detach().cpu().double()
    print("block", cin, cout, N, H)
    for k in ["y", "r.bn3.weight", "r.conv3.weight", "r.bn2.weight", "r.conv2.weight", "r.bn1.weight", "r.conv1.weight", "x"]:
        d = float(res["f64"][k].norm())
        print("  %-16s ours %.2e  ref32 %.2e" % (k, float((ours[k] - res["f64"][k]).norm()) / d,
                                                 float((res["f32"][k] - res["f64"][k]).norm()) / d))
