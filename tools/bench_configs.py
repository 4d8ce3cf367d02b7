"""Throughput of the BASELINE.json configurations that are NOT the bench line (they are parity-test cases; this tool
records their speed for profiles/):  config 3 = 2-stack HG + ASN agent joint-train iteration (bs 24, 256x256),
config 5 = 8-stack hourglass (bs 16, 256x256).  Device-resident inputs, CUDA events, synthetic data.
usage (GPU box): python tools/bench_configs.py [--steps 10]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F
from pose_adv_aug_b200 import synth, HourglassTrainer, FlatRMSprop, agent
from pose_adv_aug_b200.models import asn_stacked_hg as M
from pose_adv_aug_b200.pylib import Evaluation, HumanPts

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda", 0)


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {}
# ---- config 5: 8-stack hourglass, bs 16 ----
S, C, N, R = 8, 256, 16, 256
net = M.create_hg(S, 1, 16, C)
net.load_state_dict(synth.make_state_dict(synth.schema_of(net), seed=1, perturb_bn=False))
tr = HourglassTrainer(net, N, R, device=dev, use_graph=True)
tr.x.copy_(synth.make_images(N, R, seed=100)); tr.t.copy_(synth.make_heatmaps(N, R, 16, seed=200))
ms = timed(tr.step_resident, args.steps)
out["config5_8stack_bs16"] = {"ms_per_step": ms, "images_per_s": N / ms * 1e3, "launches_per_step": tr.launches_per_step,
                              "loss": float(tr.loss), "hbm_floor_ms": 7.81,
                              "note": "S=8 C=256 bs=16 256x256 train step, one CUDA graph; HBM floor from SURVEY 8d"}
del tr, net
torch.cuda.empty_cache()

# ---- config 3: joint-train iteration (joint-train-pose-s-r-agent.py:245-296) ----
S, C, N, R = 2, 256, 24, 256
net = M.create_hg(S, 1, 16, C)
net.load_state_dict(synth.make_state_dict(synth.schema_of(net), seed=1, perturb_bn=False))
asn = M.create_asn(C, C, 7, 7, is_aug=True)
asn.load_state_dict(synth.make_state_dict(synth.schema_of(asn), seed=2, perturb_bn=False))
asn.to(dev)
tr = HourglassTrainer(net, N, R, device=dev, use_graph=True)
x = synth.make_images(N, R, seed=100).to(dev)
tr.x.copy_(x); tr.t.copy_(synth.make_heatmaps(N, R, 16, seed=200))
pts = torch.randint(4, 60, (N, 16, 2), device=dev).float()
np.random.seed(0)


def agent_aug_iteration():
    # half-hg forward (hg.train(), agent.eval()) -> on-GPU sampling -> [data pipeline: out of scope; the targets of the
    # re-augmented batch are rendered on the GPU from their joint coordinates] -> full train step -> PCK on the GPU
    net.train(); asn.eval()
    with torch.no_grad():
        ps, pr = net(x, asn, is_half_hg=True, is_aug=True)
    _, _, si, ri = agent.sample_scale_rotation(ps, pr)
    hm, _ = HumanPts.pts2heatmap(pts, [64, 64])
    tr.t.copy_(hm)
    tr.step_resident()
    return Evaluation.accuracy(tr.heatmaps()[-1], tr.t, list(range(16)))


ms = timed(agent_aug_iteration, args.steps)
out["config3_agent_aug_iteration"] = {"ms_per_iteration": ms, "images_per_s": N / ms * 1e3,
                                      "note": "half-hg(train BN)+ASN(eval) fwd, on-GPU sampling, GPU target rendering, full "
                                              "train step (CUDA graph), GPU PCK; bs 24"}
ms_half = timed(lambda: net(x, asn, is_half_hg=True, is_aug=True), args.steps)
out["config3_half_hg_asn_forward_ms"] = ms_half

# agent update (train_agent_sr, :323-410): hg.eval(), agent.train(); KL loss on [N,7]; gradients reach the ASN only
opt = FlatRMSprop(asn, lr=2.5e-4)
tgt = torch.softmax(torch.randn(N, 7, device=dev), dim=1)


def agent_update():
    net.eval(); asn.train()
    ps, pr = net(x, asn, is_half_hg=True, is_aug=True)
    loss = F.kl_div(torch.log(F.softmax(ps, dim=1) + 1e-7), tgt, reduction="mean") * 7 + \
        F.kl_div(torch.log(F.softmax(pr, dim=1) + 1e-7), tgt, reduction="mean") * 7
    opt.zero_grad()
    loss.backward()
    opt.step()


ms = timed(agent_update, args.steps)
out["config3_agent_update"] = {"ms_per_update": ms, "images_per_s": N / ms * 1e3,
                               "note": "half-hg(eval) + ASN(train) fwd + KL + ASN bwd + flat RMSprop; module path (no graph)"}
print(json.dumps(out))
