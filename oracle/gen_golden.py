"""Generate tests/golden/* by running THE REFERENCE ITSELF (oracle/_ref, produced from
/root/reference by oracle/make_ref.py) on deterministic synthetic weights and inputs.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):
    python oracle/make_ref.py && python oracle/gen_golden.py
The outputs are small .npz / .json fixtures committed under tests/golden/; they pin
oracle/hg_oracle.py (tests/test_oracle.py) and, through it, the CUDA path.
"""
import json
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_ref, synth          # noqa: E402
from oracle import hg_oracle as O           # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (name, stacks, modules, classes, chan, batch, res)
HG_CASES = [
    ("hg_s1_c128_n2_r64", 1, 1, 16, 128, 2, 64),      # BASELINE.json configs[0]
    ("hg_s2_c32_n2_r64", 2, 1, 16, 32, 2, 64),        # two stacks: forth_conv / in_conv path
    ("hg_s1_c16_m2_n3_r64", 1, 2, 16, 16, 3, 64),     # num_modules=2, odd batch, 1x1 neck
]
FULL_GRAD_KEYS = ["conv1.weight", "bn1.weight", "bn1.bias", "residual1.adapter.weight",
                  "residual1.adapter.bias", "residual3.conv2.weight", "hg.0.neck.0.conv2.weight",
                  "hg.0.neck.0.bn2.weight", "hg.0.up1.0.conv3.weight", "out_conv.0.weight",
                  "out_conv.0.bias", "linear.0.1.bias", "forth_conv.0.weight", "in_conv.0.weight",
                  "in_conv.0.bias"]
STAT_KEYS = ["bn1.running_mean", "bn1.running_var", "residual1.bn3.running_mean",
             "hg.0.neck.0.bn2.running_var", "hg.0.neck.0.bn2.running_mean", "linear.0.1.running_var"]


def _np(t):
    return t.detach().cpu().numpy()


def run_hg_case(ref, name, S, M, K, C, N, R, dtype):
    net = ref.create_hg(num_stacks=S, num_modules=M, num_classes=K, chan=C)
    schema = O.hg_schema(S, M, K, C)
    sd = synth.make_state_dict(schema, seed=1, dtype=dtype)
    net = net.to(dtype)
    missing = net.load_state_dict(sd, strict=True)
    x = synth.make_images(N, R, seed=2, dtype=dtype)
    t = synth.make_heatmaps(N, R, K, seed=3, dtype=dtype)
    out = OrderedDict()
    # eval-mode forward (running statistics)
    net.eval()
    with torch.no_grad():
        oe = net(x)
    for i, o in enumerate(oe):
        out["eval_out%d" % i] = _np(o)
    # train-mode forward + loss + backward (stack-hg.py:153-164)
    net.train()
    outs = net(x)
    loss = 0
    for o in outs:
        tmp = (o - t) ** 2
        loss = loss + tmp.sum() / tmp.numel()
    net.zero_grad()
    loss.backward()
    for i, o in enumerate(outs):
        out["train_out%d" % i] = _np(o)
    out["loss"] = _np(loss)
    names, norms = [], []
    for k, p in net.named_parameters():
        names.append(k)
        norms.append(float(p.grad.double().norm()))
        if k in FULL_GRAD_KEYS:
            out["grad:" + k] = _np(p.grad)
    out["grad_norms"] = np.asarray(norms, dtype=np.float64)
    new_sd = net.state_dict()
    for k in STAT_KEYS:
        if k in new_sd:
            out["stat:" + k] = _np(new_sd[k])
    # one RMSprop step (stack-hg.py:51-52,165), record a few updated parameters
    opt = torch.optim.RMSprop(net.parameters(), lr=2.5e-4, alpha=0.99, eps=1e-8, momentum=0, weight_decay=0)
    opt.step()
    new_sd = net.state_dict()
    # (only well-conditioned ones: the first RMSprop step is ~ lr*10*sign(g), so parameters whose
    # true gradient is 0 -- conv biases feeding a BN -- move by +-2.5e-3 according to noise sign)
    for k in ("out_conv.0.weight", "out_conv.0.bias"):
        out["step:" + k] = _np(new_sd[k])
    return names, out


def run_asn_case(ref, dtype):
    """half-hg + ASN (is_aug) and whole-hg + ASN, models/asn_stacked_hg.py:159-171,298-307."""
    S, M, K, C, N, R = 1, 1, 16, 32, 2, 256     # AvgPool2d(4) + Linear(C,7) needs a 4x4 neck => R=256
    net = ref.create_hg(num_stacks=S, num_modules=M, num_classes=K, chan=C).to(dtype)
    asn = ref.create_asn(chan_in=C, chan_out=C, scale_num=7, rotation_num=7, is_aug=True).to(dtype)
    sd = synth.make_state_dict(O.hg_schema(S, M, K, C), seed=11, dtype=dtype)
    asd = synth.make_state_dict(O.asn_schema(C, C, 7, 7, is_aug=True), seed=12, dtype=dtype)
    net.load_state_dict(sd, strict=True)
    asn.load_state_dict(asd, strict=True)
    x = synth.make_images(N, R, seed=13, dtype=dtype)
    out = OrderedDict()
    import contextlib
    import io
    net.train()
    asn.eval()
    with contextlib.redirect_stdout(io.StringIO()):
        ps, pr = net(x, asn, is_half_hg=True, is_aug=True)
    out["half_scale_asneval"] = _np(ps)
    out["half_rot_asneval"] = _np(pr)
    net.load_state_dict(sd, strict=True)
    net.eval()
    asn.train()
    with contextlib.redirect_stdout(io.StringIO()):
        ps, pr = net(x, asn, is_half_hg=True, is_aug=True)
    out["half_scale_asntrain"] = _np(ps)
    out["half_rot_asntrain"] = _np(pr)
    # agent loss of joint-train-pose-s-r-agent.py:399-407 on a fixed target distribution
    tgt = torch.softmax(synth.make_tensor("asn_target", (N, 7), seed=14, dtype=dtype), dim=1)
    import torch.nn.functional as F
    ls = F.kl_div(torch.log(F.softmax(ps, dim=1) + 1e-7), tgt, reduction="mean") * 7
    lr_ = F.kl_div(torch.log(F.softmax(pr, dim=1) + 1e-7), tgt, reduction="mean") * 7
    loss = ls + lr_
    asn.zero_grad()
    net.zero_grad()
    loss.backward()
    out["agent_loss"] = _np(loss)
    names, norms = [], []
    for k, p in asn.named_parameters():
        names.append(k)
        norms.append(float(p.grad.double().norm()))
    out["grad_norms"] = np.asarray(norms)
    out["grad:fc_scale.weight"] = _np(asn.fc_scale.weight.grad)
    out["grad:merge1.conv2.weight"] = _np(asn.merge1.conv2.weight.grad)
    out["grad:residual_skip1.conv1.weight"] = _np(asn.residual_skip1.conv1.weight.grad)
    out["hg_has_grad"] = np.asarray([int(any(p.grad is not None and float(p.grad.abs().sum()) > 0
                                              for p in net.parameters()))])
    return names, out


def main():
    ref, crit = make_ref.load()
    if ref is None:
        make_ref.generate()
        ref, crit = make_ref.load()
    assert ref is not None, "reference not available; run in the build container"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    meta = {"cases": {}, "generator": "oracle/gen_golden.py", "torch": torch.__version__,
            "reference": "zhiqiangdon/pose-adv-aug models/asn_stacked_hg.py via oracle/make_ref.py"}
    import contextlib
    import io
    for (name, S, M, K, C, N, R) in HG_CASES:
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            with contextlib.redirect_stdout(io.StringIO()):
                names, out = run_hg_case(ref, name, S, M, K, C, N, R, dtype)
            if tag == "f64":      # fp64 "truth": keep only what the noise-floor rules need
                out = OrderedDict((k, v) for k, v in out.items()
                                  if k.startswith("train_out") or k in ("loss", "grad_norms"))
            np.savez_compressed(os.path.join(GOLD, "%s_%s.npz" % (name, tag)), **out)
        meta["cases"][name] = {"stacks": S, "modules": M, "classes": K, "chan": C, "batch": N, "res": R,
                               "param_names": names}
        print("golden", name, "loss", float(out["loss"]))
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        with contextlib.redirect_stdout(io.StringIO()):
            names, out = run_asn_case(ref, dtype)
        np.savez_compressed(os.path.join(GOLD, "asn_c32_n2_r256_%s.npz" % tag), **out)
    meta["cases"]["asn_c32_n2_r256"] = {"chan": 32, "batch": 2, "res": 256, "param_names": names}
    print("golden asn agent_loss", float(out["agent_loss"]))
    # state_dict schemas of the headline models (names, shapes, order)
    with contextlib.redirect_stdout(io.StringIO()):
        hg = ref.create_hg(num_stacks=2, num_modules=1, num_classes=16, chan=256)
        asn = ref.create_asn(chan_in=256, chan_out=256, scale_num=7, rotation_num=7, is_aug=True)
        asn_d = ref.create_asn(chan_in=256, chan_out=256, is_dropout=True)
    meta["schema_hg_s2_m1_k16_c256"] = [[k, list(v.shape)] for k, v in hg.state_dict().items()]
    meta["schema_asn_aug_c256"] = [[k, list(v.shape)] for k, v in asn.state_dict().items()]
    meta["schema_asn_dropout_c256"] = [[k, list(v.shape)] for k, v in asn_d.state_dict().items()]
    meta["n_params_hg_s2_c256"] = sum(p.numel() for p in hg.parameters())
    # Criterion golden (pylib/Criterion.py)
    p = torch.sigmoid(synth.make_tensor("crit_pred", (2, 4, 8, 8), seed=5))
    g = (synth.make_tensor("crit_gt", (2, 4, 8, 8), seed=6) > 0.5).float()
    w = 1.0 + 4.0 * g
    p1 = p.clone().requires_grad_(True)
    l2 = crit.weighted_L2(p1, g, w)
    l2.backward()
    p2 = p.clone().requires_grad_(True)
    ce = crit.weighted_sigmoid_crossentropy(p2, g, w)
    ce.backward()
    np.savez_compressed(os.path.join(GOLD, "criterion_f32.npz"), l2=_np(l2), ce=_np(ce),
                        l2_grad=_np(p1.grad), ce_grad=_np(p2.grad))
    with open(os.path.join(GOLD, "meta.json"), "w") as f:
        json.dump(meta, f, indent=0)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
