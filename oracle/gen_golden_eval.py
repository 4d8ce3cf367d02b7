"""Generate tests/golden/eval_*.npz by running THE REFERENCE's own pylib/Evaluation.py, HumanAug flip helpers and
numpy sampler (oracle/_ref, produced from /root/reference by oracle/make_ref.py) on deterministic inputs.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):
    python oracle/make_ref.py && python oracle/gen_golden_eval.py
The fixtures pin oracle/eval_oracle.py (tests/test_oracle_eval.py) and, through it, the CUDA kernels of
pose_adv_aug_b200/csrc/eval.cu (tests/test_eval_gpu.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_ref, synth          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def eval_inputs(n=6, seed=40):
    """Heat-map pairs with the edge cases the reference code branches on: absent joints (all-zero map -> (0,0)),
    an all-negative map, peaks on the border (no quarter-pixel shift), exact ties (first maximum wins)."""
    tgt = synth.make_heatmaps(n, 256, 16, seed=seed)                       # [n,16,64,64]
    noise = synth.make_tensor("eval_noise", tuple(tgt.shape), seed=seed + 1, lo=-0.05, hi=0.05)
    shift = torch.roll(tgt, shifts=(1, -2), dims=(2, 3))
    out = 0.8 * tgt + 0.35 * shift + noise
    out[0, 3] = -out[0, 3].abs() - 0.01                                     # max <= 0
    out[1, 5] = 0.0                                                         # all zeros: max == 0
    out[2, 7] = 0.0
    out[2, 7, 0, 0] = 1.0                                                   # peak in the corner
    out[2, 8] = 0.0
    out[2, 8, 63, 63] = 1.0
    out[3, 9] = 0.0
    out[3, 9, 10, 20] = 0.7
    out[3, 9, 30, 5] = 0.7                                                  # tie: first (row-major) wins
    out[4, 2] = 0.0
    out[4, 2, 1, 40] = 0.9                                                  # py == 2 boundary of the refinement test
    out[4, 2, 0, 40] = 0.3
    r = np.random.Generator(np.random.PCG64(seed + 2))
    center = torch.from_numpy(r.uniform(200, 900, size=(n, 2))).float()
    scale = torch.from_numpy(r.uniform(0.8, 3.5, size=(n,))).float()
    rot = torch.from_numpy(np.where(r.random(n) < 0.5, 0.0, r.uniform(-40, 40, size=n))).float()
    grnd_pts = torch.from_numpy(r.uniform(-50, 1200, size=(n, 16, 2))).float()
    grnd_pts[0, 0] = 0.0                                                    # invisible joint (use_zero boundary)
    grnd_pts[1, 1, 0] = -3.0
    normalizer = torch.from_numpy(r.uniform(30, 120, size=(n,))).float()
    return out.contiguous(), tgt.contiguous(), center, scale, rot, grnd_pts, normalizer


def pts_inputs(n=5, j=16, seed=60):
    """Joint coordinates as the datasets hand them to pts2heatmap (1-based crop coordinates after TransformPts, floats):
    interior points, points within 3 px of every border (clipped blobs), exactly on the accept/reject limits, outside."""
    r = np.random.Generator(np.random.PCG64(seed))
    p = r.integers(-6, 72, size=(n, j, 2)).astype(np.float64)
    p[0, 0] = [1, 1]
    p[0, 1] = [64, 64]
    p[0, 2] = [0, 10]
    p[0, 3] = [10, 65]
    p[0, 4] = [64, 1]
    p[0, 5] = [2.6, 61.4]                                        # non-integer: int() truncation of pt -+ tmp_size
    p[0, 6] = [33.5, 0.5]
    p[1, 0] = [3, 3]
    p[1, 1] = [62, 2]
    return p


def run_dropout_case(ref):
    """ASN dropout mode (models/asn_stacked_hg.py:79-136,172-190,308-322,340), never invoked by the shipped scripts:
    half-hg mask logits, then the whole two-stack net with np.random-sampled masks, loss and a few gradients."""
    import contextlib
    import io
    from collections import OrderedDict
    from oracle import hg_oracle as O
    S, C, N, R = 2, 32, 2, 256
    net = ref.create_hg(num_stacks=S, num_modules=1, num_classes=16, chan=C)
    asn = ref.create_asn(chan_in=C, chan_out=C, is_dropout=True)
    sd = synth.make_state_dict(O.hg_schema(S, 1, 16, C), seed=21)
    asd = synth.make_state_dict(O.asn_schema(C, C, is_dropout=True), seed=22)
    net.load_state_dict(sd, strict=True)
    asn.load_state_dict(asd, strict=True)
    x = synth.make_images(N, R, seed=23)
    t = synth.make_heatmaps(N, R, 16, seed=24)
    out = OrderedDict()
    net.train()
    asn.eval()
    with contextlib.redirect_stdout(io.StringIO()):
        pm = net(x, asn, is_half_hg=True, is_dropout=True)
    out["half_pred_mask"] = pm.detach().numpy()
    net.load_state_dict(sd, strict=True)
    np.random.seed(4321)
    with contextlib.redirect_stdout(io.StringIO()):
        outs, pm2, indexes = net(x, asn, is_dropout=True)
    loss = 0
    for o in outs:
        tmp = (o - t) ** 2
        loss = loss + tmp.sum() / tmp.numel()
    net.zero_grad()
    loss.backward()
    for i, o in enumerate(outs):
        out["out%d_sub" % i] = o.detach().numpy()[:, :, ::3, ::3].copy()          # sub-sampled (fixture size)
        out["out%d_sum" % i] = np.asarray(o.detach().double().sum().item())
    out["pred_mask"] = pm2.detach().numpy()
    out["indexes"] = indexes.numpy()
    out["loss"] = loss.detach().numpy()
    for k in ("hg.0.skip1.0.conv3.weight", "hg.0.neck.0.conv2.weight", "hg.1.skip2.0.conv1.weight", "out_conv.1.weight",
              "residual1.conv1.weight"):
        out["grad:" + k] = dict(net.named_parameters())[k].grad.numpy()
    # agent-side gradient through the mask logits (pretrain-style loss on pred_mask), hg in eval mode
    net.load_state_dict(sd, strict=True)
    net.eval()
    asn.train()
    with contextlib.redirect_stdout(io.StringIO()):
        pm3 = net(x, asn, is_half_hg=True, is_dropout=True)
    w = synth.make_tensor("mask_w", tuple(pm3.shape), seed=25)
    asn.zero_grad()
    (pm3 * w).sum().backward()
    out["half_pred_mask_asntrain"] = pm3.detach().numpy()
    out["grad:asn.out_conv.weight"] = asn.out_conv.weight.grad.numpy()
    out["grad:asn.merge4.conv3.weight"] = asn.merge4.conv3.weight.grad.numpy()
    return out


def main():
    ref, crit, ev = make_ref.load(with_eval=True)
    if ev is None:
        make_ref.generate()
        ref, crit, ev = make_ref.load(with_eval=True)
    assert ev is not None, "reference not available; run in the build container"
    os.makedirs(GOLD, exist_ok=True)
    out, tgt, center, scale, rot, grnd_pts, normalizer = eval_inputs()
    res = [64, 64]
    idx = [0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13, 14, 15]
    g = {}
    g["get_preds_out"] = ev.get_preds(out.clone()).numpy()
    g["get_preds_tgt"] = ev.get_preds(tgt.clone()).numpy()
    g["final_preds"] = ev.final_preds(out.clone(), center, scale, res, rot).numpy()
    preds, gts = ev.get_preds(out.clone()), ev.get_preds(tgt.clone())
    norm = torch.ones(preds.size(0)) * out.size(3) / 10
    g["dists_hm"] = ev.calc_dists(preds, gts, norm).numpy()
    g["accuracy"] = ev.accuracy(out.clone(), tgt.clone(), idx).numpy()
    g["accuracy_thr02"] = ev.accuracy(out.clone(), tgt.clone(), [0, 3, 9], thr=0.2).numpy()
    g["accuracy_origin_res"] = ev.accuracy_origin_res(out.clone(), center, scale, res, grnd_pts, normalizer, rot).numpy()
    fp = ev.final_preds(out.clone(), center, scale, res, rot)
    g["dists_origin"] = ev.calc_dists(fp, grnd_pts, normalizer, use_zero=True).numpy()
    # per_person_pckh against ground truth that is consistent with the crops (points near the predictions)
    near = fp.clone() + torch.from_numpy(np.random.Generator(np.random.PCG64(7)).uniform(-60, 60, size=tuple(fp.shape))).float()
    g["near_pts"] = near.numpy()
    g["per_person_pckh"] = ev.per_person_pckh(out.clone(), tgt.clone(), center, scale, res, near, normalizer, rot).numpy()
    # flip test, stack-hg.py:225-232: HumanAug.flip_channels + shuffle_channels_for_horizontal_flipping, mean
    out2 = synth.make_tensor("flip_out2", tuple(out.shape), seed=44, lo=0.0, hi=1.0)
    o2 = torch.from_numpy(out2.numpy()[:, :, :, ::-1].copy()).float()                     # flip_channels (HumanAug.py:199-210)
    flip_indxs = np.array([[1, 4], [0, 5], [12, 13], [11, 14], [10, 15], [2, 3]])       # HumanAug.py:182
    for i in range(flip_indxs.shape[0]):
        i1, i2 = flip_indxs[i]
        tmp = o2.narrow(1, int(i1), 1).clone()
        o2.narrow(1, int(i1), 1).copy_(o2.narrow(1, int(i2), 1))
        o2.narrow(1, int(i2), 1).copy_(tmp)
    g["flip_merged_sample"] = ((out + o2) / 2).numpy()[:, :, ::7, ::5].copy()           # sub-sampled (fixture size)
    g["flip_merged_sum"] = np.asarray(((out + o2) / 2).double().sum().item())
    # agent sampling, joint-train-pose-s-r-agent.py:252-271: softmax, then np.random.choice per sample (scale, rotation)
    logits_s = synth.make_tensor("agent_logits_s", (24, 7), seed=45, lo=-3, hi=3)
    logits_r = synth.make_tensor("agent_logits_r", (24, 7), seed=46, lo=-3, hi=3)
    ps = torch.softmax(logits_s, dim=1).numpy()
    pr = torch.softmax(logits_r, dim=1).numpy()
    np.random.seed(1234)
    si, ri = [], []
    for j in range(ps.shape[0]):
        si.append(np.random.choice(7, 1, p=ps[j])[0])
        ri.append(np.random.choice(7, 1, p=pr[j])[0])
    g["agent_probs_s"], g["agent_probs_r"] = ps, pr
    g["agent_idx_s"], g["agent_idx_r"] = np.asarray(si, dtype=np.int64), np.asarray(ri, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "eval_n6_f32.npz"), **g)
    hp = make_ref.load(with_eval=True, with_pts=True)[3]
    pp = pts_inputs()
    hms, vps = [], []
    for n in range(pp.shape[0]):
        hm, vp = hp.pts2heatmap(pp[n].copy(), [64, 64], sigma=1)
        hms.append(hm)
        vps.append(vp)
    hm = torch.from_numpy(np.stack(hms)).float()               # as the datasets do (data/mpii_for_mpii.py:151-153)
    gp = {"pts": pp, "valid_pts": np.stack(vps).astype(np.float32), "heatmap_sum": hm.double().sum(dim=(2, 3)).numpy(),
          "heatmap_rows": hm[:, :, ::9].numpy().copy()}
    # (HumanPts.heatmap2pts cannot be executed: `max.gt(0).repeat(1, 1, 2)` at :133 assumes a kept dimension that
    #  torch.max(dim) has not returned since torch 0.2 -- it is dead code in the reference, only called from
    #  commented-out lines; oracle/eval_oracle.py restates its evident intent, parity unpinned for that one function)
    hm2, vp2 = hp.pts2heatmap(pp[0].copy(), [64, 64], sigma=2)
    gp["heatmap_sigma2_sum"] = hm2.sum(axis=(1, 2))
    np.savez_compressed(os.path.join(GOLD, "humanpts_f32.npz"), **gp)
    d = run_dropout_case(ref)
    np.savez_compressed(os.path.join(GOLD, "dropout_s2_c32_n2_r256_f32.npz"), **d)
    print("dropout golden: loss", float(d["loss"]), "indexes", d["indexes"].tolist())
    print("wrote eval golden:", {k: v.shape for k, v in g.items()})
    print("accuracy", g["accuracy"][:4], "origin", g["accuracy_origin_res"][:4], "pp", g["per_person_pckh"])


if __name__ == "__main__":
    main()
