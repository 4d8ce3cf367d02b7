"""-m gpu parity tests of the callers either side of the training path (csrc/eval.cu through the C-ABI and the
drop-in modules pylib/Evaluation.py, pylib/HumanAug.py, agent.py): bit-exact against oracle/eval_oracle.py and
against the reference's own outputs in tests/golden/eval_n6_f32.npz (integer / index / compare work), 1e-6 relative
for the fp32 distances and probabilities."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hgk_testlib import DEV, call, ptr                      # noqa: E402
from oracle import eval_oracle as E, synth                  # noqa: E402
from oracle.gen_golden_eval import eval_inputs              # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "eval_n6_f32.npz"))


def _mods():
    from pose_adv_aug_b200.pylib import Evaluation, HumanAug
    from pose_adv_aug_b200 import agent
    return Evaluation, HumanAug, agent


def test_get_preds_final_preds_vs_reference_golden():
    Ev = _mods()[0]
    out, tgt, center, scale, rot, grnd_pts, normalizer = eval_inputs()
    p = Ev.get_preds(out.to(DEV))
    assert p.is_cuda and tuple(p.shape) == (6, 16, 2)
    assert np.array_equal(p.cpu().numpy(), G["get_preds_out"])
    assert np.array_equal(Ev.get_preds(tgt.to(DEV)).cpu().numpy(), G["get_preds_tgt"])
    fp = Ev.final_preds(out.to(DEV), center, scale, [64, 64], rot)
    assert np.array_equal(fp.cpu().numpy(), G["final_preds"])


def test_accuracy_family_vs_reference_golden():
    Ev = _mods()[0]
    out, tgt, center, scale, rot, grnd_pts, normalizer = eval_inputs()
    o, t = out.to(DEV), tgt.to(DEV)
    idx = E.MPII_IDXS
    acc = Ev.accuracy(o, t, idx)
    assert acc.is_cuda
    np.testing.assert_allclose(acc.cpu().numpy(), G["accuracy"], rtol=1e-6)
    np.testing.assert_allclose(Ev.accuracy(o, t, [0, 3, 9], thr=0.2).cpu().numpy(), G["accuracy_thr02"], rtol=1e-6)
    norm = torch.ones(6) * 64 / 10
    d = Ev.calc_dists(Ev.get_preds(o), Ev.get_preds(t), norm)
    np.testing.assert_allclose(d.cpu().numpy(), G["dists_hm"], rtol=1e-6)
    fp = Ev.final_preds(o, center, scale, [64, 64], rot)
    np.testing.assert_allclose(Ev.calc_dists(fp, grnd_pts, normalizer, use_zero=True).cpu().numpy(), G["dists_origin"], rtol=1e-6)
    np.testing.assert_allclose(Ev.accuracy_origin_res(o, center, scale, [64, 64], grnd_pts, normalizer, rot).cpu().numpy(),
                               G["accuracy_origin_res"], rtol=1e-6)
    near = torch.from_numpy(G["near_pts"])
    np.testing.assert_allclose(Ev.per_person_pckh(o, t, center, scale, [64, 64], near, normalizer, rot).cpu().numpy(),
                               G["per_person_pckh"], rtol=1e-6)
    # dist_acc on one row, incl. the "no valid entry" case (ref:53-54)
    row = torch.from_numpy(G["dists_hm"][3]).to(DEV)
    assert abs(float(Ev.dist_acc(row)) - E.dist_acc(G["dists_hm"][3])) < 1e-6
    assert float(Ev.dist_acc(torch.full((5,), -1.0, device=DEV))) == -1.0


@pytest.mark.parametrize("shape", [(24, 16, 64, 64), (3, 5, 32, 48), (1, 1, 8, 8)])
def test_peaks_vs_oracle_seeded(shape):
    """bs-24 config-2 size and ragged shapes (H != W keeps the reference's divide-by-H quirk), random maps."""
    Ev = _mods()[0]
    N, J, H, W = shape
    x = synth.make_tensor("peaks", shape, seed=N + J, lo=-0.2, hi=1.0)
    x[0, 0] = -1.0
    p = Ev.get_preds(x.to(DEV)).cpu().numpy()
    assert np.array_equal(p, E.get_preds(x.numpy()))
    if H == W:
        r = np.random.Generator(np.random.PCG64(5))
        center = r.uniform(100, 800, size=(N, 2)); scale = r.uniform(0.7, 3.0, size=N); rot = r.uniform(-30, 30, size=N)
        rot[::2] = 0
        fp = Ev.final_preds(x.to(DEV), torch.from_numpy(center), torch.from_numpy(scale), [W, H], torch.from_numpy(rot))
        ref = E.final_preds(x.numpy(), center, scale, [W, H], rot)
        assert np.array_equal(fp.cpu().numpy(), ref)


def test_peaks_properties_at_full_size():
    """size-independent properties: a planted strict maximum is found at its (x, y); shifting the map moves the peak."""
    Ev = _mods()[0]
    N, J, H, W = 24, 16, 64, 64
    x = torch.rand(N, J, H, W, device=DEV) * 0.5
    ys = torch.randint(0, H, (N, J), device=DEV); xs = torch.randint(0, W, (N, J), device=DEV)
    x.view(N, J, -1).scatter_(2, (ys * W + xs).unsqueeze(2), 2.0)
    p = Ev.get_preds(x)
    assert torch.equal(p[..., 0], (xs + 1).float()) and torch.equal(p[..., 1], (ys + 1).float())
    acc = Ev.accuracy(x, x, list(range(16)))
    assert float(acc[0]) == 1.0                                   # a map against itself: every distance is 0


def test_flip_merge_and_helpers():
    _, HA, _ = _mods()
    out = eval_inputs()[0]
    out2 = synth.make_tensor("flip_out2", tuple(out.shape), seed=44, lo=0.0, hi=1.0)
    m = HA.flip_merge(out.to(DEV), out2.to(DEV))
    ref = E.flip_merge(out.numpy(), out2.numpy())
    assert np.array_equal(m.cpu().numpy(), ref)
    assert np.array_equal(m.cpu().numpy()[:, :, ::7, ::5], G["flip_merged_sample"])
    f = HA.flip_channels(out2.to(DEV))
    assert torch.equal(f.cpu(), torch.flip(out2, dims=[3]))
    assert torch.equal(HA.flip_channels(f), out2.to(DEV))          # involution
    s = out2.clone().to(DEV)
    r = HA.shuffle_channels_for_horizontal_flipping(s)
    assert r is s                                                   # in place, as the reference
    perm = list(range(16))
    for a, b in E.FLIP_PAIRS:
        perm[a], perm[b] = perm[b], perm[a]
    assert torch.equal(s.cpu(), out2[:, perm])
    m3 = HA.flip_merge(out[0].to(DEV), out2[0].to(DEV))             # 3-dim variant (ref:185-186,203-204)
    assert np.array_equal(m3.cpu().numpy(), ref[0])
    with pytest.raises(Exception):
        HA.flip_merge(out, out2)                                    # CPU tensors: no fallback


def test_agent_sampling_equals_numpy_choice_stream():
    agent = _mods()[2]
    ls = synth.make_tensor("agent_logits_s", (24, 7), seed=45, lo=-3, hi=3)
    lr = synth.make_tensor("agent_logits_r", (24, 7), seed=46, lo=-3, hi=3)
    np.random.seed(1234)                                            # the golden generator's seed
    ps, pr, si, ri = agent.sample_scale_rotation(ls.to(DEV), lr.to(DEV))
    assert si.is_cuda and si.dtype == torch.int64
    np.testing.assert_allclose(ps.cpu().numpy(), G["agent_probs_s"], rtol=2e-6)
    np.testing.assert_allclose(pr.cpu().numpy(), G["agent_probs_r"], rtol=2e-6)
    assert np.array_equal(si.cpu().numpy(), G["agent_idx_s"])       # == the reference's np.random.choice loop
    assert np.array_equal(ri.cpu().numpy(), G["agent_idx_r"])
    # distribution property at a larger size: empirical frequencies follow the probabilities
    logits = torch.tensor([[0.0, 1.0, 2.0, -1.0]], device=DEV).repeat(20000, 1)
    u = torch.from_numpy(np.random.Generator(np.random.PCG64(3)).random(20000))
    probs, idx = agent.softmax_sample(logits, u)
    freq = torch.bincount(idx, minlength=4).double() / 20000
    assert float((freq - probs[0].double().cpu().to(freq.device)).abs().max()) < 0.015
    # u -> 1 edge: searchsorted(side='right') semantics, last bin
    p1, i1 = agent.softmax_sample(logits[:2], torch.tensor([0.0, 1.0 - 1e-12], dtype=torch.float64))
    assert i1.tolist() == [0, 3]


def test_mask_mul_kernels():
    """ASN dropout (ref models/asn_stacked_hg.py:79-100): y = act(x) * nearest_upsample(mask), gx = g * mask."""
    N, H, W, C = 3, 16, 16, 8
    x = synth.make_tensor("mm_x", (N, H, W, C), seed=1).to(DEV)
    sc = (synth.make_tensor("mm_s", (C,), seed=2, lo=0.5, hi=1.5)).to(DEV)
    sh = synth.make_tensor("mm_t", (C,), seed=3, lo=-0.3, hi=0.3).to(DEV)
    mask = (synth.make_tensor("mm_m", (N, 4, 4), seed=4) > -0.5).float().to(DEV)
    y = torch.empty_like(x)
    call("mask_mul_fwd", ptr(x), ptr(sc), ptr(sh), 1, ptr(mask), N, H, W, C, 4, 4, ptr(y))
    up = mask.repeat_interleave(4, dim=1).repeat_interleave(4, dim=2).unsqueeze(3)
    ref = torch.relu(torch.addcmul(sh, x, sc)) * up
    assert torch.allclose(y, ref, rtol=0, atol=1e-6)
    call("mask_mul_fwd", ptr(x), 0, 0, 0, ptr(mask), N, H, W, C, 4, 4, ptr(y))
    assert torch.equal(y, x * up)
    g = synth.make_tensor("mm_g", (N, H, W, C), seed=5).to(DEV)
    gx = torch.ones_like(g)
    call("mask_mul_bwd", ptr(g), ptr(mask), N, H, W, C, 4, 4, ptr(gx), 1)
    assert torch.equal(gx, 1 + g * up)
    call("mask_mul_bwd", ptr(g), ptr(mask), N, H, W, C, 4, 4, ptr(gx), 0)
    assert torch.equal(gx, g * up)
    # 4x4 feature map: scale 1 (the neck, ref:92)
    x4 = synth.make_tensor("mm_x4", (N, 4, 4, C), seed=6).to(DEV)
    y4 = torch.empty_like(x4)
    call("mask_mul_fwd", ptr(x4), 0, 0, 0, ptr(mask), N, 4, 4, C, 4, 4, ptr(y4))
    assert torch.equal(y4, x4 * mask.unsqueeze(3))


def test_pts2heatmap_and_heatmap2pts():
    """GPU ground-truth rendering (pylib/HumanPts.py:36-48,82-116) bit-exact vs the oracle and the reference golden,
    incl. clipped blobs at every border, the accept/reject limits, non-integer points and sigma=2; heatmap2pts vs the
    oracle restatement and as the inverse of the rendering for interior points."""
    from pose_adv_aug_b200.pylib import HumanPts as HP
    g = np.load(os.path.join(ROOT, "tests", "golden", "humanpts_f32.npz"))
    pts = g["pts"]
    hm, vp = HP.pts2heatmap(torch.from_numpy(pts).to(DEV), [64, 64], sigma=1)
    assert tuple(hm.shape) == (5, 16, 64, 64) and hm.is_cuda
    assert np.array_equal(vp.cpu().numpy(), g["valid_pts"])
    assert np.array_equal(hm.cpu().numpy()[:, :, ::9], g["heatmap_rows"])
    for n in range(pts.shape[0]):
        ref, _ = E.pts2heatmap(pts[n].copy(), [64, 64], sigma=1)
        assert np.array_equal(hm[n].cpu().numpy(), torch.from_numpy(ref).float().numpy())
    h1, v1 = HP.pts2heatmap(torch.from_numpy(pts[0]).to(DEV), [64, 64], sigma=2)       # [J,2] form, 13x13 blob
    ref2, _ = E.pts2heatmap(pts[0].copy(), [64, 64], sigma=2)
    assert np.array_equal(h1.cpu().numpy(), torch.from_numpy(ref2).float().numpy())
    hr, _ = HP.pts2heatmap(torch.from_numpy(pts[:2]).to(DEV), [48, 80], sigma=1)        # H != W
    for n in range(2):
        ref, _ = E.pts2heatmap(pts[n].copy(), [48, 80], sigma=1)
        assert np.array_equal(hr[n].cpu().numpy(), torch.from_numpy(ref).float().numpy())
    # heatmap2pts
    p2 = HP.heatmap2pts(hm)
    assert np.array_equal(p2.cpu().numpy(), E.heatmap2pts(hm.cpu().numpy()))
    interior = (pts[..., 0] >= 4) & (pts[..., 0] <= 60) & (pts[..., 1] >= 4) & (pts[..., 1] <= 60)
    q = p2.cpu().numpy()
    assert np.array_equal(q[interior][:, 0], pts[interior][:, 0]) and np.array_equal(q[interior][:, 1], pts[interior][:, 1] + 0.5)
    out = eval_inputs()[0]
    assert np.array_equal(HP.heatmap2pts(out.to(DEV)).cpu().numpy(), E.heatmap2pts(out.numpy()))
    # a full config-2 batch of targets: every map sums to the blob mass or to 0
    big = torch.randint(4, 60, (24, 16, 2), device=DEV).float()
    hb, _ = HP.pts2heatmap(big, [64, 64])
    mass = float(HP.gaussian_blob(1).double().sum())
    assert torch.allclose(hb.double().sum(dim=(2, 3)), torch.full((24, 16), mass, dtype=torch.float64, device=DEV), rtol=1e-6)
