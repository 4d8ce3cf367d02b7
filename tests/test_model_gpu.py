"""Block- and whole-network parity (-m gpu) of the drop-in modules against the CPU oracle
(oracle/hg_oracle.py, itself pinned to the reference by tests/golden) and directly against the
committed golden outputs of the reference model.

Tolerances (north_star): heat-maps and loss <= 1e-3 relative; gradients by the noise-floor rule
of SURVEY.md 0.5 (fp32 vs fp64 of the reference itself differ by ~1e-2 whole-net)."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from hgk_testlib import DEV, relerr, rnd
from oracle import hg_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
META = json.load(open(os.path.join(GOLD, "meta.json")))


def _mods():
    from pose_adv_aug_b200.models import asn_stacked_hg as M
    return M


def _load(net, sd):
    missing = net.load_state_dict(OrderedDict((k, v.float()) for k, v in sd.items()), strict=True)
    return net


def _grads64(sd64, fwd):
    """fp64 oracle gradients of sum(outputs * weights) for a block test."""
    leaves = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd64.items() if O.is_trainable(k))
    work = OrderedDict((k, leaves.get(k, v)) for k, v in sd64.items())
    return leaves, work


# relative-L2 gradient bounds of test_residual_block, ~5x the values measured on B200: plain-TF32 gradients 4.9e-4 .. 5.9e-4,
# 3xTF32 data gradients + fp32 weight gradients 5e-7 .. 2e-6, fp32 SIMT 2e-7 .. 8e-7
L2TOL_TC, L2TOL_TCP, L2TOL_SIMT = 3e-3, 1e-5, 5e-6


def _rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("cfg", [(64, 128, True, 2, 16), (128, 128, False, 3, 8), (256, 256, False, 2, 4),
                                 (32, 32, False, 2, 1), (12, 24, True, 3, 6)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("mode", ["simt", "tc", "tc_precise"])
def test_residual_block(cfg, training, mode, monkeypatch):
    """simt: fp32 SIMT kernels; tc: tcgen05 (3xTF32 forward, TF32 gradients); tc_precise: tcgen05 with
    3xTF32 data gradients + fp32 weight gradients."""
    M = _mods()
    monkeypatch.setattr(M, "CONV_PATH", 1 if mode == "simt" else 0)
    monkeypatch.setattr(M, "PRECISE_GRADS", mode == "tc_precise")
    cin, cout, adapter, N, H = cfg
    import torch.nn as nn
    ad = nn.Conv2d(cin, cout, 1) if adapter else None
    blk = M._Residual(cin, cout, ad)
    schema = O.residual_schema("r", cin, cout, adapter)
    sd64 = synth.make_state_dict(schema, seed=5, dtype=torch.float64)
    _load(blk, OrderedDict((k[2:], v) for k, v in sd64.items()))
    blk.to(DEV)
    blk.train(training)
    x64 = rnd("x", (N, cin, H, H), -1.0, 2.0)
    leaves, work = _grads64(sd64, None)
    xr = x64.clone().requires_grad_(True)
    st = O.BNState(training)
    yr = O.residual(work, "r", xr, st)
    gy = rnd("gy", tuple(yr.shape))
    yr.backward(gy)
    x = x64.float().to(DEV).requires_grad_(True)
    y = blk(x)
    assert y.shape == yr.shape
    tol = 5e-4 if (N * H * H <= 4 and training) else (1e-5 if mode == "simt" else 3e-5)     # 3xTF32 forward: fp32-class
    assert relerr(y, yr) < tol
    y.backward(gy.float().to(DEV))
    # "tc": gradients run on plain TF32 operands (10-bit mantissa), two to three chained convolutions deep
    # On small tensors a single ReLU-mask flip (3xTF32 forward differs from fp64 by ~1e-6, fp32 by ~1e-7)
    # already moves a gradient by ~1e-3 of its max, hence the wider tc_precise bound; the kernels
    # themselves are checked at fp32 / TF32 class in test_kernels_gpu.py.
    gtol = 5e-2 if (N * H * H <= 4 and training) else {"simt": 5e-5, "tc": 1.5e-1, "tc_precise": 5e-3}[mode]
    if mode != "simt" and N * H * H <= 64:
        gtol = 1.5e-1        # 32 pixels: one flipped ReLU mask is 1/32 of a channel's gradient
    assert relerr(x.grad, xr.grad) < gtol
    # next to the max-abs bound (which has to absorb single ReLU-mask flips) a relative-L2 bound of the arithmetic's class:
    # a wrong tap, halo or BatchNorm-apply fusion moves the L2 distance by O(1), mask flips and TF32 rounding do not
    l2 = {"x": _rel_l2(x.grad, xr.grad)}
    for k, p in blk.named_parameters():
        ref = leaves["r." + k].grad
        if training and k.endswith("bias") and "bn" not in k:
            continue      # conv bias in front of a train-mode BN: true gradient is 0 (rounding noise on both sides)
        assert relerr(p.grad, ref) < gtol, k
        l2[k] = _rel_l2(p.grad, ref)
    print("residual block %s %s train=%s: worst gradient rel-L2 %.2e (%s)" % (cfg, mode, training, max(l2.values()),
                                                                            max(l2, key=l2.get)))
    # (not on <= 64 pixels: BatchNorm over 2 samples / one flipped ReLU mask = 1/32 of a bias gradient, see above)
    if not (N * H * H <= 64 and (training or mode != "simt")):
        l2tol = {"simt": L2TOL_SIMT, "tc": L2TOL_TC, "tc_precise": L2TOL_TCP}[mode]
        assert max(l2.values()) < l2tol, l2
    if training:
        for k, v in st.updates.items():
            if "num_batches" in k:
                continue
            assert relerr(blk.state_dict()[k[2:]], v) < 1e-5, k
        assert int(blk.bn1.num_batches_tracked) == 1


def _hg_case(case, dtype=torch.float32):
    c = META["cases"][case]
    S, Mo, K, C, N, R = c["stacks"], c["modules"], c["classes"], c["chan"], c["batch"], c["res"]
    sd = synth.make_state_dict(O.hg_schema(S, Mo, K, C), seed=1, dtype=dtype)
    x = synth.make_images(N, R, seed=2, dtype=dtype)
    t = synth.make_heatmaps(N, R, K, seed=3, dtype=dtype)
    return c, sd, x, t


@pytest.mark.parametrize("case", [k for k in META["cases"] if k.startswith("hg_")])
def test_whole_net_vs_reference_golden(case):
    """Heat-maps / loss / gradients / running stats / RMSprop step against the REFERENCE's outputs."""
    M = _mods()
    c, sd, x, t = _hg_case(case)
    g32 = np.load(os.path.join(GOLD, case + "_f32.npz"))
    g64 = np.load(os.path.join(GOLD, case + "_f64.npz"))
    net = _load(M.create_hg(c["stacks"], c["modules"], c["classes"], c["chan"]), sd).to(DEV)
    assert list(net.state_dict().keys()) == list(sd.keys())
    xd, td = x.to(DEV), t.to(DEV)
    net.eval()
    with torch.no_grad():
        oe = net(xd)
    assert isinstance(oe, list) and len(oe) == c["stacks"]
    for i, o in enumerate(oe):
        ref = torch.from_numpy(g32["eval_out%d" % i])
        assert tuple(o.shape) == tuple(ref.shape)
        assert relerr(o, ref) < 1e-3           # north_star: <= 1e-3 relative on fp32 heat-maps
        assert relerr(o, ref) < 1e-4           # what fp32 arithmetic actually achieves
    net.train()
    outs = net(xd)
    loss = 0
    for o in outs:
        tmp = (o - td) ** 2
        loss = loss + tmp.sum() / tmp.numel()
    opt = torch.optim.RMSprop(net.parameters(), lr=2.5e-4, alpha=0.99, eps=1e-8, momentum=0, weight_decay=0)
    opt.zero_grad()
    loss.backward()
    for i, o in enumerate(outs):
        ref64 = torch.from_numpy(g64["train_out%d" % i])
        floor = relerr(torch.from_numpy(g32["train_out%d" % i]), ref64)
        assert relerr(o, ref64) < 1e-3
        assert relerr(o, ref64) < 3 * floor + 1e-4
    assert abs(float(loss) - float(g64["loss"])) / float(g64["loss"]) < 1e-3
    names = c["param_names"]
    params = dict(net.named_parameters())
    norms = np.array([float(params[k].grad.double().norm()) for k in names])
    gn32, gn64 = g32["grad_norms"], g64["grad_norms"]
    floor = np.abs(gn32 - gn64).max() / gn64.max()
    # whole-net: block-level error equals the fp32 reference's (test_residual_block); the whole-net figure is
    # dominated by a handful of discrete ReLU-mask flips amplified by the BN chain, hence the wide factor
    # default mode: TF32 gradients add ~4e-3 rel-L2, below the 1.2e-2 fp32 noise floor of the headline config
    assert np.abs(norms - gn64).max() / gn64.max() < max(8 * floor + 2e-4, 1e-2)
    for k in g32.files:
        if k.startswith("grad:"):
            name = k[5:]
            if name.endswith("bias") and "bn" not in name and "linear.0.1" not in name and "out_conv" not in name:
                continue
            assert relerr(params[name].grad, torch.from_numpy(g32[k])) < max(5e-2, 4 * floor), name
        if k.startswith("stat:"):
            assert relerr(net.state_dict()[k[5:]], torch.from_numpy(g32[k])) < 2e-3, k
    opt.step()
    for k in g32.files:
        if k.startswith("step:"):
            # RMSprop's first step is ~ lr*10*sign(g): robust except where g is at rounding-noise level
            d = (net.state_dict()[k[5:]].cpu() - torch.from_numpy(g32[k])).abs()
            assert float((d > 1e-5).float().mean()) < 0.02, k
            assert float(d.max()) < 6e-3, k


@pytest.mark.parametrize("precise", [False, True])
def test_whole_net_vs_oracle_two_stack_c64(precise, monkeypatch):
    """A shape class the goldens do not hold (C=64, N=4, 256x256), oracle run live in fp64.  (256x256 so that the deepest rung is
    4x4 as in the reference's configurations: at 128x128 it is 2x2 -- BatchNorm statistics over 16 samples -- and the whole-net
    gradient becomes a lottery of single ReLU flips: mathematically equivalent forward variants, 1e-6 apart, measured anywhere
    between 2.3e-3 and 4.7e-2.)"""
    M = _mods()
    monkeypatch.setattr(M, "PRECISE_GRADS", precise)
    S, Mo, K, C, N, R = 2, 1, 16, 64, 4, 256
    sd64 = synth.make_state_dict(O.hg_schema(S, Mo, K, C), seed=31, dtype=torch.float64)
    x64 = synth.make_images(N, R, seed=32, dtype=torch.float64)
    t64 = synth.make_heatmaps(N, R, K, seed=33, dtype=torch.float64)
    outs64, loss64, grads64, st = O.train_step(sd64, x64, t64, S, Mo)
    sd32 = OrderedDict((k, v.float() if v.is_floating_point() else v) for k, v in sd64.items())
    outs32, loss32, grads32, _ = O.train_step(sd32, x64.float(), t64.float(), S, Mo)
    net = _load(M.create_hg(S, Mo, K, C), sd64).to(DEV)
    net.train()
    outs = net(x64.float().to(DEV))
    loss = O.mse_loss(outs, t64.float().to(DEV))
    loss.backward()
    for o, r in zip(outs, outs64):
        assert relerr(o, r) < 1e-3
    assert abs(float(loss) - float(loss64)) / float(loss64) < 1e-4
    names = [k for k in sd64 if O.is_trainable(k)]
    params = dict(net.named_parameters())
    num = sum(float((params[k].grad.cpu().double() - grads64[k]).pow(2).sum()) for k in names)
    den = sum(float(grads64[k].pow(2).sum()) for k in names)
    num32 = sum(float((grads32[k].double() - grads64[k]).pow(2).sum()) for k in names)
    ours, floor = (num / den) ** 0.5, (num32 / den) ** 0.5
    print("two-stack C=64 gradient rel-L2 vs fp64: ours %.3e, fp32 oracle %.3e (precise=%s)" % (ours, floor, precise))
    # see the note in test_whole_net_vs_reference_golden; TF32 gradients: absolute bound instead
    assert ours < (8 * floor + 1e-4 if precise else 3 * floor + 1e-3), (ours, floor)
    for k, v in st.updates.items():
        if "num_batches" not in k:
            assert relerr(net.state_dict()[k], v) < 1e-4, k


def test_state_dict_schema_and_checkpoint_roundtrip():
    M = _mods()
    net = M.create_hg(2, 1, 16, 256)
    got = [[k, list(v.shape)] for k, v in net.state_dict().items()]
    assert got == META["schema_hg_s2_m1_k16_c256"]
    asn = M.create_asn(256, 256, 7, 7, is_aug=True)
    assert [[k, list(v.shape)] for k, v in asn.state_dict().items()] == META["schema_asn_aug_c256"]
    # DataParallel-style `module.` prefix is a pure key rename (utils/checkpoint.py:57-67 copies by name)
    small = M.create_hg(1, 1, 16, 32).to(DEV)
    x = synth.make_images(2, 64, seed=4).to(DEV)
    small.eval()
    with torch.no_grad():
        a = small(x)[0].clone()
    sd = OrderedDict(("module." + k, v.clone()) for k, v in small.state_dict().items())
    other = M.create_hg(1, 1, 16, 32).to(DEV)
    own = other.state_dict()
    for k, v in sd.items():                      # the reference loader's name-wise copy
        own[k[len("module."):]].copy_(v)
    other.eval()
    with torch.no_grad():
        b = other(x)[0]
    assert torch.equal(a, b)


def test_asn_half_hg_and_agent_update():
    """half-hg + ASN (ref models/asn_stacked_hg.py:159-171) in both mode combinations the scripts use."""
    M = _mods()
    g = np.load(os.path.join(GOLD, "asn_c32_n2_r256_f32.npz"))
    C, N, R = 32, 2, 256
    sd = synth.make_state_dict(O.hg_schema(1, 1, 16, C), seed=11)
    asd = synth.make_state_dict(O.asn_schema(C, C, 7, 7, is_aug=True), seed=12)
    net = _load(M.create_hg(1, 1, 16, C), sd).to(DEV)
    asn = _load(M.create_asn(C, C, 7, 7, is_aug=True), asd).to(DEV)
    x = synth.make_images(N, R, seed=13).to(DEV)
    # joint-train-pose-s-r-agent.py:207-208,250: hg.train(), agent.eval()
    net.train()
    asn.eval()
    rm0 = net.state_dict()["bn1.running_mean"].clone()
    ps, pr = net(x, asn, is_half_hg=True, is_aug=True)
    assert tuple(ps.shape) == (N, 7) and tuple(pr.shape) == (N, 7)
    assert relerr(ps, torch.from_numpy(g["half_scale_asneval"])) < 1e-3
    assert relerr(pr, torch.from_numpy(g["half_rot_asneval"])) < 1e-3
    assert not torch.equal(rm0, net.state_dict()["bn1.running_mean"])     # train-mode half pass updates stats
    # :323-324,342,399-410: hg.eval(), agent.train(); KL loss; gradients reach the ASN only
    _load(net, sd)
    net.eval()
    asn.train()
    ps, pr = net(x, asn, is_half_hg=True, is_aug=True)
    assert relerr(ps, torch.from_numpy(g["half_scale_asntrain"])) < 1e-3
    assert relerr(pr, torch.from_numpy(g["half_rot_asntrain"])) < 1e-3
    tgt = torch.softmax(synth.make_tensor("asn_target", (N, 7), seed=14), dim=1).to(DEV)
    loss = O.agent_kl_loss(ps, tgt, 7) + O.agent_kl_loss(pr, tgt, 7)
    assert abs(float(loss) - float(g["agent_loss"])) < 1e-3 * abs(float(g["agent_loss"])) + 1e-6
    for p in list(net.parameters()) + list(asn.parameters()):
        if p.grad is not None:
            p.grad.zero_()
    loss.backward()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in net.parameters() if p.grad is not None)
    assert relerr(asn.fc_scale.weight.grad, torch.from_numpy(g["grad:fc_scale.weight"])) < 1e-3
    assert relerr(asn.merge1.conv2.weight.grad, torch.from_numpy(g["grad:merge1.conv2.weight"])) < 2e-2
    assert relerr(asn.residual_skip1.conv1.weight.grad, torch.from_numpy(g["grad:residual_skip1.conv1.weight"])) < 2e-2
    names = META["cases"]["asn_c32_n2_r256"]["param_names"]
    params = dict(asn.named_parameters())
    norms = np.array([float(params[k].grad.double().norm()) for k in names])
    g64 = np.load(os.path.join(GOLD, "asn_c32_n2_r256_f64.npz"))["grad_norms"]
    floor = np.abs(g["grad_norms"] - g64).max() / g64.max()
    assert np.abs(norms - g64).max() / g64.max() < 2 * floor + 1e-3
    # whole hg + ASN returns (outs, scale, rot) (ref :338)
    net.train()
    asn.eval()
    outs, ps2, pr2 = net(x, asn, is_aug=True)
    assert isinstance(outs, list) and len(outs) == 1 and tuple(outs[0].shape) == (N, 16, 64, 64)
    # standalone ASN on NCHW feature dict (ref :401)
    feats = {"neck": torch.rand(N, C, 4, 4, device=DEV), "skip1": torch.rand(N, C, 64, 64, device=DEV),
             "skip2": torch.rand(N, C, 32, 32, device=DEV), "skip3": torch.rand(N, C, 16, 16, device=DEV),
             "skip4": torch.rand(N, C, 8, 8, device=DEV)}
    s1, r1 = asn(feats, is_aug=True)
    (s_ref, r_ref), _ = O.asn_forward(OrderedDict((k, v.cpu().double()) for k, v in asn.state_dict().items()),
                                      dict((k, v.cpu().double()) for k, v in feats.items()), training=False)
    assert relerr(s1, s_ref) < 1e-3 and relerr(r1, r_ref) < 1e-3


def test_trainer_matches_module_path_and_graph_replay(monkeypatch):
    """HourglassTrainer (fused step, CUDA graph) == module forward + torch loss + backward + RMSprop.
    Run with fp32-class gradients so that three RMSprop steps stay comparable with the CPU oracle (the test
    net has a 1x1 neck over 4 samples: TF32 gradient noise is amplified there)."""
    M = _mods()
    monkeypatch.setattr(M, "PRECISE_GRADS", True)
    from pose_adv_aug_b200 import HourglassTrainer, FlatRMSprop
    S, Mo, K, C, N, R = 2, 1, 16, 32, 4, 64
    sd = synth.make_state_dict(O.hg_schema(S, Mo, K, C), seed=41)
    x = synth.make_images(N, R, seed=42)
    t = synth.make_heatmaps(N, R, K, seed=43)
    nets = [_load(M.create_hg(S, Mo, K, C), sd).to(DEV) for _ in range(3)]
    # (a) module path + flat optimizer
    a = nets[0]
    a.train()
    opt = FlatRMSprop(a, lr=2.5e-4)
    losses_a = []
    grad_a = None
    for _ in range(3):
        outs = a(x.to(DEV))
        loss = O.mse_loss(outs, t.to(DEV))
        opt.zero_grad()
        loss.backward()
        first = grad_a is None
        if first:
            grad_a = opt.store.grad.clone()
        opt.step()
        if first:
            head_a = a.out_conv[1].weight.detach().clone()       # after ONE update: the two paths have not diverged yet
        losses_a.append(float(loss))
    # (b) fused trainer without graph, (c) with graph
    for net, use_graph in ((nets[1], False), (nets[2], True)):
        tr = HourglassTrainer(net, N, R, lr=2.5e-4, use_graph=use_graph)
        losses = [float(tr.step(x.pin_memory(), t.pin_memory()))]
        # same kernels on both paths: the first step's flat gradient agrees to rounding (dL/dout comes from torch ops on one
        # side and from the fused MSE kernel on the other; weight gradients are float atomics)
        g_b = tr.store.grad
        assert float((g_b - grad_a).norm() / grad_a.norm()) < 1e-4
        # one RMSprop update later the last head agrees tightly: an element moves by ~lr*10*sign(g), so only elements whose
        # gradient is at the rounding-noise level (sign flips) may differ (1e-4 = 4 % of one step)
        d1 = (net.out_conv[1].weight.detach() - head_a).abs()
        assert float((d1 > 1e-4).float().mean()) < 0.02
        losses += [float(tr.step(x.pin_memory(), t.pin_memory())) for _ in range(2)]
        # ... after which RMSprop's first steps (~lr*10*sign(g) per weight whatever |g| is) turn that rounding noise into
        # visible loss differences on this tiny, ill-conditioned net: tight on steps 0-1, loose on step 2
        for i, (la, lb) in enumerate(zip(losses_a, losses)):
            assert abs(la - lb) < (1e-4 if i < 2 else 1e-2) * abs(la), (losses_a, losses)
        # Parameters: the two paths differ only by rounding of dL/dout (torch ops vs the fused MSE kernel)
        # and atomics order, but that 1e-7 noise is amplified towards the stem by the BN chain
        # and RMSprop's first steps move a weight by ~lr*10*sign(g) whatever |g| is.  So: the heads agree
        # tightly, everything stays within the 3-step RMSprop travel.
        for (k, pa), (_, pb) in zip(a.state_dict().items(), net.state_dict().items()):
            if pa.is_floating_point():
                d = (pa - pb).abs()
                # parameters travel at most ~3 * lr * 10 = 7.5e-3 in three RMSprop steps whatever the gradient is, which bounds
                # their difference; BatchNorm running statistics are batch statistics of a diverging (chaotic) pair of nets --
                # 64 samples per channel at the 4x4 rung of this 4-image batch -- and only have to stay in the same ballpark
                if "running_" in k:
                    assert float(d.max()) < 1e-1 * max(1.0, float(pa.abs().max())), k
                else:
                    assert float(d.max()) < 2e-2, k
        hm = tr.heatmaps()
        assert len(hm) == S and tuple(hm[0].shape) == (N, K, R // 4, R // 4)
    # oracle: same three steps on CPU fp32.  Step 0 (before any update) must agree tightly; after that every
    # weight has moved by ~lr*10*sign(g) (RMSprop's first steps), which turns rounding-level gradient
    # differences into visible loss differences on this tiny, ill-conditioned net (1x1 neck over 4 samples).
    sdo = OrderedDict((k, v.clone()) for k, v in sd.items())
    sq = OrderedDict((k, torch.zeros_like(v)) for k, v in sdo.items() if O.is_trainable(k))
    for i in range(3):
        _, lo, _, _ = O.train_step(sdo, x, t, S, Mo, square_avg=sq)
        assert abs(float(lo) - losses_a[i]) < (1e-4 if i == 0 else 5e-2) * abs(float(lo))


def test_multistream_graph_equals_single_stream():
    """The dependency-scheduled multi-stream CUDA graph computes exactly what the single-stream list does."""
    M = _mods()
    from pose_adv_aug_b200 import HourglassTrainer
    S, Mo, K, C, N, R = 2, 1, 16, 64, 4, 128
    sd = synth.make_state_dict(O.hg_schema(S, Mo, K, C), seed=51)
    x = synth.make_images(N, R, seed=52).to(DEV)
    t = synth.make_heatmaps(N, R, K, seed=53).to(DEV)
    res = []
    for ns in (1, 4):
        net = _load(M.create_hg(S, Mo, K, C), sd).to(DEV)
        tr = HourglassTrainer(net, N, R, use_graph=True, n_streams=ns)
        losses, hms, grads = [], [], []
        for _ in range(4):
            losses.append(float(tr.step(x, t)))
            hms.append([h.clone() for h in tr.heatmaps()])
            grads.append(tr.store.grad.clone())
        res.append((losses, hms, grads))
    (la, ha, ga), (lb, hb, gb) = res
    # step 0 sees identical parameters: the forward is order-independent up to the fp64-atomic BN statistics and the
    # gradients up to the order of the float atomics of the weight-gradient reduction
    assert abs(la[0] - lb[0]) < 1e-6 * abs(la[0])
    for a, b in zip(ha[0], hb[0]):
        assert relerr(a, b) < 1e-5
    assert float((ga[0] - gb[0]).norm() / ga[0].norm()) < 1e-4
    # step 1 runs on parameters updated from those gradients (a missing dependency edge would show here)
    assert abs(la[1] - lb[1]) < 2e-6 * abs(la[1]), (la, lb)
    for a, b in zip(ha[1], hb[1]):
        assert relerr(a, b) < 5e-3
    # later steps inherit the atomics ordering noise through RMSprop's sign-like first updates (lr * g / sqrt(0.01 g^2):
    # a sign flip of a near-zero gradient moves a weight by 2 * 10 * lr); the same spread exists between two
    # single-stream runs, so only the loss is compared from here on
    for i in (2, 3):
        assert abs(la[i] - lb[i]) < (2e-3 if i == 2 else 5e-3) * abs(la[i]), (la, lb)


@pytest.mark.gpu
def test_in_graph_bucketed_collective_sees_final_gradients():
    """The bucketed all-reduce captured inside the step graph (trainer.py) on ONE GPU: the collective is replaced by
    `bucket *= 2` on the communication stream.  Every bucket must be doubled exactly once and only after its last writer
    (weight-gradient atomics, the per-bucket 3x3 unpack, BatchNorm-backward finalisers) has run: the gradient buffer after a
    replay equals twice the plain trainer's, replay after replay."""
    M = _mods()
    from pose_adv_aug_b200 import HourglassTrainer
    S, Mo, K, C, N, R = 2, 1, 16, 64, 4, 128
    sd = synth.make_state_dict(O.hg_schema(S, Mo, K, C), seed=81)
    x = synth.make_images(N, R, seed=82).to(DEV)
    t = synth.make_heatmaps(N, R, K, seed=83).to(DEV)
    tr_a = HourglassTrainer(_load(M.create_hg(S, Mo, K, C), sd).to(DEV), N, R, use_graph=True)
    tr_b = HourglassTrainer(_load(M.create_hg(S, Mo, K, C), sd).to(DEV), N, R, use_graph=True, ar_buckets=5,
                            fake_collective=lambda g: g.mul_(2.0))
    assert tr_b.ar_in_graph and not tr_a.ar_in_graph
    rng = tr_b.bucket_ranges()
    assert len(rng) >= 3 and rng[0][0] == 0 and rng[-1][1] == tr_b.store.numel
    assert sum(1 for r in tr_b.plan.bwd if r[2] == "unpack_add_grads") >= 1
    for it in range(3):
        # same parameters on both sides before every step (the doubled gradient changes RMSprop's eps term only, but the
        # comparison should not depend on that)
        tr_b.store.flat.copy_(tr_a.store.flat)
        tr_b.square_avg.copy_(tr_a.square_avg)
        tr_b.store.fbuf_flat.copy_(tr_a.store.fbuf_flat)
        la, lb = float(tr_a.step(x, t)), float(tr_b.step(x, t))
        assert abs(la - lb) < 1e-6 * abs(la)
        ga, gb = tr_a.store.grad, tr_b.store.grad
        for lo, hi in rng:
            d = float((gb[lo:hi] - 2.0 * ga[lo:hi]).norm() / (2.0 * ga[lo:hi]).norm())
            assert d < 1e-4, (it, lo, hi, d)


@pytest.mark.gpu
def test_tensor_core_stem_equals_ffma_stem_in_a_train_step(monkeypatch):
    """HGK_STEM_TC=1 (space-to-depth + tcgen05 image-tile kernel with ksize = 4, 3xTF32) against the default FFMA stem:
    same heat-maps, loss and gradients of one captured train step up to the products' rounding."""
    M = _mods()
    from pose_adv_aug_b200 import HourglassTrainer
    S, Mo, K, C, N, R = 2, 1, 16, 64, 4, 128
    sd = synth.make_state_dict(O.hg_schema(S, Mo, K, C), seed=91)
    x = synth.make_images(N, R, seed=92).to(DEV)
    t = synth.make_heatmaps(N, R, K, seed=93).to(DEV)
    res = []
    for on in ("0", "1"):
        monkeypatch.setenv("HGK_STEM_TC", on)
        tr = HourglassTrainer(_load(M.create_hg(S, Mo, K, C), sd).to(DEV), N, R, use_graph=True)
        names = [r[2] for r in tr.plan.fwd]
        assert ("stem_s2d_image" in names) == (on == "1") and ("stem_conv7_fwd" in names) == (on == "0")
        loss = float(tr.step(x, t))
        res.append((loss, [h.clone() for h in tr.heatmaps()], tr.store.grad.clone()))
    (la, ha, ga), (lb, hb, gb) = res
    assert abs(la - lb) < 1e-5 * abs(la)
    for a, b in zip(ha, hb):
        assert relerr(a, b) < 1e-4
    assert float((ga - gb).norm() / ga.norm()) < 5e-3


def test_cpu_input_fails_loudly():
    M = _mods()
    from pose_adv_aug_b200 import HGKError
    net = M.create_hg(1, 1, 16, 32)
    with pytest.raises(HGKError):
        net(torch.rand(1, 3, 64, 64))


def test_asn_dropout_mode_vs_reference_golden():
    """ASN dropout mode (ref models/asn_stacked_hg.py:79-136,172-190,308-322,340): half-hg mask logits,
    whole two-stack net with np.random-sampled masks (same numpy seed as the reference run -> same cells), loss,
    gradients, the agent-side gradient, and the `dropout_masks=` path of _Hourglass.forward against the oracle."""
    M = _mods()
    g = np.load(os.path.join(GOLD, "dropout_s2_c32_n2_r256_f32.npz"))
    S, C, N, R = 2, 32, 2, 256
    sd = synth.make_state_dict(O.hg_schema(S, 1, 16, C), seed=21)
    asd = synth.make_state_dict(O.asn_schema(C, C, is_dropout=True), seed=22)
    net = _load(M.create_hg(S, 1, 16, C), sd).to(DEV)
    asn = _load(M.create_asn(C, C, is_dropout=True), asd).to(DEV)
    x = synth.make_images(N, R, seed=23).to(DEV)
    t = synth.make_heatmaps(N, R, 16, seed=24).to(DEV)
    net.train()
    asn.eval()
    pm = net(x, asn, is_half_hg=True, is_dropout=True)
    assert tuple(pm.shape) == (N, 1, 4, 4)
    assert relerr(pm, torch.from_numpy(g["half_pred_mask"])) < 1e-3
    _load(net, sd)
    np.random.seed(4321)
    outs, pm2, indexes = net(x, asn, is_dropout=True)
    assert isinstance(outs, list) and len(outs) == S and tuple(pm2.shape) == (N, 1, 4, 4)
    assert indexes.dtype == torch.long and np.array_equal(indexes.cpu().numpy(), g["indexes"])
    loss = 0
    for o in outs:
        tmp = (o - t) ** 2
        loss = loss + tmp.sum() / tmp.numel()
    assert abs(float(loss) - float(g["loss"])) < 1e-3 * float(g["loss"])
    for i, o in enumerate(outs):
        ref = torch.from_numpy(g["out%d_sub" % i])
        assert relerr(o[:, :, ::3, ::3], ref) < 1e-3
    for p in net.parameters():
        if p.grad is not None:
            p.grad.zero_()
    loss.backward()
    params = dict(net.named_parameters())
    for k, tol in (("out_conv.1.weight", 1e-3), ("hg.1.skip2.0.conv1.weight", 3e-2), ("hg.0.skip1.0.conv3.weight", 3e-2),
                   ("hg.0.neck.0.conv2.weight", 3e-2), ("residual1.conv1.weight", 5e-2)):
        assert relerr(params[k].grad, torch.from_numpy(g["grad:" + k])) < tol, k
    # agent side: hg.eval(), asn.train(), gradient of sum(pred_mask * w) reaches the ASN only
    _load(net, sd)
    net.eval()
    asn.train()
    pm3 = net(x, asn, is_half_hg=True, is_dropout=True)
    assert relerr(pm3, torch.from_numpy(g["half_pred_mask_asntrain"])) < 1e-3
    w = synth.make_tensor("mask_w", tuple(pm3.shape), seed=25).to(DEV)
    for p in list(asn.parameters()) + list(net.parameters()):
        if p.grad is not None:
            p.grad.zero_()
    (pm3 * w).sum().backward()
    assert relerr(asn.out_conv.weight.grad, torch.from_numpy(g["grad:asn.out_conv.weight"])) < 1e-3
    assert relerr(asn.merge4.conv3.weight.grad, torch.from_numpy(g["grad:asn.merge4.conv3.weight"])) < 2e-2
    assert all(float(p.grad.abs().sum()) == 0.0 for p in net.parameters() if p.grad is not None)
    # _Hourglass.forward(x, dropout_masks=...) (ref:184-190) and _Hourglass whole dropout mode (ref:210) vs the oracle
    hg = net.hg[0]
    hg.train()
    feat = synth.make_tensor("hg_feat", (N, C, 64, 64), seed=26, lo=0.0, hi=1.0)
    masks = torch.ones(N, 1, 4, 4)
    masks[0, 0, 1, 2] = 0
    masks[0, 0, 3, 3] = 0
    masks[1, 0, 0, 0] = 0
    y = hg(feat.to(DEV), dropout_masks=masks.to(DEV))
    sd0 = OrderedDict((k[len("hg.0."):], v.double()) for k, v in sd.items() if k.startswith("hg.0."))
    st = O.BNState(True)
    neck, s1, s2, s3, s4 = O.hourglass_down(sd0, "", feat.double(), st, 1) if False else (None,) * 5
    sdp = OrderedDict(("h." + k, v) for k, v in sd0.items())
    neck, s1, s2, s3, s4 = O.hourglass_down(sdp, "h", feat.double(), st, 1)
    neck, s1, s2, s3, s4 = (O.dropout(v, masks.double()) for v in (neck, s1, s2, s3, s4))
    yref = O.hourglass_up(sdp, "h", neck, s1, s2, s3, s4, st, 1)
    assert relerr(y, yref) < 1e-3
    asn.eval()
    np.random.seed(7)
    y2, pmh, idxh, dm = hg(feat.to(DEV), asn, is_dropout=True)
    assert tuple(dm.shape) == (N, 1, 4, 4) and float(dm.sum()) == N * 14 and tuple(idxh.shape) == (N, 2)
    for i in range(N):
        for j in range(2):
            assert float(dm[i, 0, int(idxh[i, j]) // 4, int(idxh[i, j]) % 4]) == 0.0
    # standalone ASN dropout head on an NCHW feature dict (ref:401,437-439)
    feats = {"neck": torch.rand(N, C, 4, 4, device=DEV), "skip1": torch.rand(N, C, 64, 64, device=DEV),
             "skip2": torch.rand(N, C, 32, 32, device=DEV), "skip3": torch.rand(N, C, 16, 16, device=DEV),
             "skip4": torch.rand(N, C, 8, 8, device=DEV)}
    m1 = asn(feats, is_dropout=True)
    mref, _ = O.asn_forward(OrderedDict((k, v.cpu().double()) for k, v in asn.state_dict().items()),
                            dict((k, v.cpu().double()) for k, v in feats.items()), training=False)
    assert relerr(m1, mref) < 1e-3


def test_config5_eight_stack_forward_vs_oracle():
    """BASELINE.json configs[4] (8-stack hourglass, C=256, 256x256) at a batch the CPU oracle finishes in seconds:
    every one of the 8 heat-map outputs within 1e-3 of the fp32 oracle, loss within 1e-3; and one full-size (bs 16)
    train step runs with a finite loss equal to the sum of its per-stack MSEs (size-independent property)."""
    M = _mods()
    S, C, N, R = 8, 256, 2, 256
    sd = synth.make_state_dict(O.hg_schema(S, 1, 16, C), seed=51, perturb_bn=False)
    net = _load(M.create_hg(S, 1, 16, C), sd).to(DEV)
    net.train()
    x = synth.make_images(N, R, seed=52)
    t = synth.make_heatmaps(N, R, 16, seed=53)
    with torch.no_grad():
        outs = net(x.to(DEV))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref, _ = O.hg_forward(OrderedDict((k, v.clone()) for k, v in sd.items()), x, S, 1, training=True)
    assert len(outs) == S
    for i, (a, b) in enumerate(zip(outs, ref)):
        assert relerr(a, b) < 1e-3, "stack %d" % i
    la, lb = float(O.mse_loss(outs, t.to(DEV))), float(O.mse_loss(ref, t))
    assert abs(la - lb) < 1e-3 * abs(lb)
    from pose_adv_aug_b200 import HourglassTrainer
    _load(net, sd)
    tr = HourglassTrainer(net, 16, R, use_graph=False)
    x16, t16 = synth.make_images(16, R, seed=54), synth.make_heatmaps(16, R, 16, seed=55)
    loss = float(tr.step(x16.pin_memory(), t16.pin_memory()))
    hm = tr.heatmaps()
    per_stack = sum(float(((h - t16.to(DEV)) ** 2).mean()) for h in hm)
    assert np.isfinite(loss) and abs(loss - per_stack) < 1e-4 * per_stack


def test_prefetched_step_equals_serial_step():
    """HourglassTrainer.prefetch()/step_prefetched() (upload of the next batch on a copy stream under the current
    step) computes exactly what step(images, heatmaps) does on the same batch sequence."""
    M = _mods()
    from pose_adv_aug_b200 import HourglassTrainer, HGKError
    S, C, N, R = 1, 32, 2, 64
    sd = synth.make_state_dict(O.hg_schema(S, 1, 16, C), seed=61)
    batches = [(synth.make_images(N, R, seed=70 + i).pin_memory(), synth.make_heatmaps(N, R, 16, seed=80 + i).pin_memory())
               for i in range(4)]
    # lr = 0: the loss of step i depends on batch i only (train-mode BN uses batch statistics), so the comparison is not
    # blurred by the chaotic amplification of last-bit gradient differences that weight updates cause on this tiny net
    ta = HourglassTrainer(_load(M.create_hg(S, 1, 16, C), sd), N, R, lr=0.0, use_graph=True, n_streams=1)
    tb = HourglassTrainer(_load(M.create_hg(S, 1, 16, C), sd), N, R, lr=0.0, use_graph=True, n_streams=1)
    la = [float(ta.step(x, t)) for x, t in batches]
    with pytest.raises(HGKError):
        tb.step_prefetched()
    tb.prefetch(*batches[0])
    lb = []
    for i in range(4):
        loss = tb.step_prefetched()
        if i + 1 < 4:
            tb.prefetch(*batches[i + 1])
        lb.append(float(loss))
    # (same kernels, same order; fp64 atomics may differ in the last bit between two runs, and every batch is different
    #  data, so a batch consumed out of order would show up at the 1e-1 level)
    for a, b in zip(la, lb):
        assert abs(a - b) < 1e-5 * abs(a), (la, lb)
    assert len(set(round(v, 6) for v in la)) == 4
    # lagged read-back: every step's loss arrives, in order, through the pinned queue while the next step runs
    tc = HourglassTrainer(_load(M.create_hg(S, 1, 16, C), sd), N, R, lr=0.0, use_graph=True, n_streams=1)
    with pytest.raises(HGKError):
        tc.pop_loss()
    tc.prefetch(*batches[0])
    lc = []
    for i in range(4):
        tc.step_prefetched(loss_to_host=True)
        if i + 1 < 4:
            tc.prefetch(*batches[i + 1])
        if i > 0:
            lc.append(tc.pop_loss())
    lc.append(tc.pop_loss())
    for a, c in zip(la, lc):
        assert abs(a - c) < 1e-5 * abs(a), (la, lc)


def test_headline_shape_train_step_vs_fp64_oracle():
    """BASELINE configs[1] shapes (2 stacks, nFeat 256, 256x256; 4 images so that the CPU oracle finishes in seconds): one
    full train step on the SHIPPING path (3xTF32 forward, plain-TF32 backward, tcgen05 instantiations <128|256, ...>) against
    the fp64 oracle.  Heat-maps and loss within 1e-3 (north_star); gradients by the noise-floor rule: our distance to fp64 is
    at most twice the fp32 reference arithmetic's own distance to fp64 (1.2e-2 rel-L2 at this config)."""
    M = _mods()
    from pose_adv_aug_b200 import HourglassTrainer
    S, Mo, K, C, N, R = 2, 1, 16, 256, 4, 256
    sd = synth.make_state_dict(O.hg_schema(S, Mo, K, C), seed=71)
    x, t = synth.make_images(N, R, seed=72), synth.make_heatmaps(N, R, K, seed=73)
    sd64 = OrderedDict((k, v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items())
    outs64, loss64, g64, _ = O.train_step(sd64, x.double(), t.double(), S, Mo)
    sd32 = OrderedDict((k, v.clone()) for k, v in sd.items())
    outs32, loss32, g32, _ = O.train_step(sd32, x, t, S, Mo)
    net = _load(M.create_hg(S, Mo, K, C), sd)
    tr = HourglassTrainer(net, N, R, device=DEV, use_graph=False)
    loss = float(tr.step(x.pin_memory(), t.pin_memory()))
    torch.cuda.synchronize()
    for h, o64 in zip(tr.heatmaps(), outs64):
        assert float((h.cpu().double() - o64).abs().max() / o64.abs().max()) < 1e-3
    assert abs(loss - float(loss64)) < 1e-3 * abs(float(loss64))
    grads = dict((k, p.grad.detach().cpu().double()) for k, p in net.named_parameters())
    num = den = num32 = 0.0
    dot = n1 = 0.0
    for k, g in g64.items():
        num += float((grads[k] - g).pow(2).sum())
        num32 += float((g32[k].double() - g).pow(2).sum())
        den += float(g.pow(2).sum())
        dot += float((grads[k] * g).sum())
        n1 += float(grads[k].pow(2).sum())
    ours, floor = (num / den) ** 0.5, (num32 / den) ** 0.5
    cosine = dot / (n1 ** 0.5 * den ** 0.5)
    print("headline-shape gradient rel-L2 vs fp64: ours %.3e, fp32 oracle %.3e, cosine %.6f" % (ours, floor, cosine))
    assert ours < 2.0 * floor + 1e-3, (ours, floor)
    assert cosine > 0.999


def test_virtual_rank_data_parallel_step():
    """nn.DataParallel semantics of stack-hg.py:49 without a second GPU: W virtual ranks run their shard one after the other
    on this device (each with its OWN BatchNorm batch statistics, as the reference's replicas), the flat gradient buffers are
    summed on the device -- what the one ncclAllReduce(sum) of the step produces -- and the flat RMSprop kernel applies the
    update with 1/W folded in.  Checked against the oracle: per-shard fp32 gradients, averaged."""
    M = _mods()
    from pose_adv_aug_b200 import HourglassTrainer
    S, Mo, K, C, Nloc, R, W = 2, 1, 16, 64, 2, 128, 2
    sd = synth.make_state_dict(O.hg_schema(S, Mo, K, C), seed=81)
    xs = [synth.make_images(Nloc, R, seed=82 + r) for r in range(W)]
    ts = [synth.make_heatmaps(Nloc, R, K, seed=92 + r) for r in range(W)]
    net = _load(M.create_hg(S, Mo, K, C), sd)
    monkey = M.PRECISE_GRADS
    M.PRECISE_GRADS = True                      # fp32-class gradients: the comparison below is on the data-parallel arithmetic
    try:
        tr = HourglassTrainer(net, Nloc, R, device=DEV, use_graph=False, distributed=False)
        acc = torch.zeros_like(tr.store.grad)
        losses = []
        for r in range(W):
            tr.x.copy_(xs[r]); tr.t.copy_(ts[r])
            tr._body_grads()                    # zero grads/loss -> forward -> MSE -> backward of this rank's shard
            acc += tr.store.grad
            losses.append(float(tr.loss_acc))
        # oracle gradients of the W shards on the same (not yet updated) weights, in fp32 (the reference's arithmetic) and fp64
        gsum = gsum64 = None
        sd64 = OrderedDict((k, v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items())
        for r in range(W):
            _, l_ref, g, _ = O.train_step(OrderedDict((k, v.clone()) for k, v in sd.items()), xs[r], ts[r], S, Mo)
            assert abs(losses[r] - float(l_ref)) < 1e-3 * abs(float(l_ref))
            _, _, g64, _ = O.train_step(OrderedDict((k, v.clone()) for k, v in sd64.items()), xs[r].double(), ts[r].double(), S, Mo)
            gsum = g if gsum is None else OrderedDict((k, gsum[k] + g[k]) for k in g)
            gsum64 = g64 if gsum64 is None else OrderedDict((k, gsum64[k] + g64[k]) for k in g64)
        tr.store.grad.copy_(acc)                # == the all-reduced buffer
        # noise-floor rule: the summed gradient is about as close to fp64 as the fp32 reference arithmetic is (factor
        # 3 instead of the usual 2: two images per virtual rank, the ReLU-mask flips behind the whole-net noise do not average)
        num = num32 = den = 0.0
        for k, p in net.named_parameters():
            num += float((p.grad.detach().cpu().double() - gsum64[k]).pow(2).sum())
            num32 += float((gsum[k].double() - gsum64[k]).pow(2).sum())
            den += float(gsum64[k].pow(2).sum())
        assert (num / den) ** 0.5 < 3.0 * (num32 / den) ** 0.5 + 1e-3, ((num / den) ** 0.5, (num32 / den) ** 0.5)
        # update with 1/W folded into the kernel: p -= lr * (g/W) / (sqrt((1-alpha) (g/W)^2) + eps)
        before = tr.store.flat.clone()
        tr.hyper[3:4].fill_(1.0 / W)
        tr._body_update()
        torch.cuda.synchronize()
        gbar = acc / W
        want = before - 2.5e-4 * gbar / ((0.01 * gbar * gbar).sqrt() + 1e-8)
        assert float((tr.store.flat - want).abs().max()) < 1e-6
    finally:
        M.PRECISE_GRADS = monkey


def test_two_rank_nccl_equals_sum_of_local_gradients():
    """tools/check_dp.py under torchrun (2 ranks, NCCL): identical parameters on every rank after 3 steps and all-reduced
    gradient == sum of the local gradients.  Needs two GPUs; skipped on a one-GPU box."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29611", os.path.join(root, "tools", "check_dp.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode()
    assert r.returncode == 0 and "DP CHECK OK" in out, out[-2000:]
