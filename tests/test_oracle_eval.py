"""oracle/eval_oracle.py (numpy restatement of pylib/Evaluation.py, the flip-test helpers and the agent sampler)
against tests/golden/eval_n6_f32.npz, which holds outputs of THE REFERENCE's own functions
(oracle/gen_golden_eval.py, run where /root/reference exists).  CPU only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import eval_oracle as E                      # noqa: E402
from oracle.gen_golden_eval import eval_inputs            # noqa: E402
from oracle import synth                                  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "eval_n6_f32.npz"))
IDX = E.MPII_IDXS


def _inputs():
    out, tgt, center, scale, rot, grnd_pts, normalizer = eval_inputs()
    return (out.numpy(), tgt.numpy(), center.numpy(), scale.numpy(), rot.numpy(), grnd_pts.numpy(), normalizer.numpy())


def test_get_preds_and_edge_cases():
    out, tgt = _inputs()[:2]
    p = E.get_preds(out)
    assert np.array_equal(p, G["get_preds_out"])
    assert np.array_equal(E.get_preds(tgt), G["get_preds_tgt"])
    assert tuple(p[0, 3]) == (0.0, 0.0) and tuple(p[1, 5]) == (0.0, 0.0)      # max <= 0 -> masked
    assert tuple(p[2, 7]) == (1.0, 1.0) and tuple(p[2, 8]) == (64.0, 64.0)    # corners, 1-based
    assert tuple(p[3, 9]) == (21.0, 11.0)                                     # tie: first maximum


def test_final_preds_dists_accuracy():
    out, tgt, center, scale, rot, grnd_pts, normalizer = _inputs()
    fp = E.final_preds(out, center, scale, [64, 64], rot)
    assert np.array_equal(fp, G["final_preds"])
    norm = np.ones(out.shape[0], dtype=np.float32) * 64 / 10
    d = E.calc_dists(E.get_preds(out), E.get_preds(tgt), norm)
    np.testing.assert_allclose(d, G["dists_hm"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(E.accuracy(out, tgt, IDX), G["accuracy"], rtol=1e-6)
    np.testing.assert_allclose(E.accuracy(out, tgt, [0, 3, 9], thr=0.2), G["accuracy_thr02"], rtol=1e-6)
    np.testing.assert_allclose(E.calc_dists(fp, grnd_pts, normalizer, use_zero=True), G["dists_origin"], rtol=1e-6)
    np.testing.assert_allclose(E.accuracy_origin_res(out, center, scale, [64, 64], grnd_pts, normalizer, rot),
                               G["accuracy_origin_res"], rtol=1e-6)
    np.testing.assert_allclose(E.per_person_pckh(out, tgt, center, scale, [64, 64], G["near_pts"], normalizer, rot),
                               G["per_person_pckh"], rtol=1e-6)


def test_flip_merge_and_agent_sampling():
    out = _inputs()[0]
    out2 = synth.make_tensor("flip_out2", tuple(out.shape), seed=44, lo=0.0, hi=1.0).numpy()
    m = E.flip_merge(out, out2)
    assert np.array_equal(m[:, :, ::7, ::5], G["flip_merged_sample"])
    assert abs(float(m.astype(np.float64).sum()) - float(G["flip_merged_sum"])) < 1e-6 * abs(float(G["flip_merged_sum"]))
    ls = synth.make_tensor("agent_logits_s", (24, 7), seed=45, lo=-3, hi=3).numpy()
    lr = synth.make_tensor("agent_logits_r", (24, 7), seed=46, lo=-3, hi=3).numpy()
    # the reference draws scale then rotation per sample from the global numpy RandomState: one uniform per draw
    np.random.seed(1234)
    u = np.random.random_sample(48).reshape(24, 2)
    ps, si = E.sample_agent(ls, u[:, 0])
    pr, ri = E.sample_agent(lr, u[:, 1])
    np.testing.assert_allclose(ps, G["agent_probs_s"], rtol=2e-6)
    np.testing.assert_allclose(pr, G["agent_probs_r"], rtol=2e-6)
    assert np.array_equal(si, G["agent_idx_s"]) and np.array_equal(ri, G["agent_idx_r"])


def test_dropout_mode_oracle_vs_reference_golden():
    """oracle hg_forward_dropout (models/asn_stacked_hg.py:79-136,172-190,308-322) vs the reference run recorded in
    tests/golden/dropout_s2_c32_n2_r256_f32.npz (same numpy seed -> same sampled cells)."""
    import torch
    from oracle import hg_oracle as O
    g = np.load(os.path.join(ROOT, "tests", "golden", "dropout_s2_c32_n2_r256_f32.npz"))
    S, C, N, R = 2, 32, 2, 256
    sd = synth.make_state_dict(O.hg_schema(S, 1, 16, C), seed=21)
    asd = synth.make_state_dict(O.asn_schema(C, C, is_dropout=True), seed=22)
    x = synth.make_images(N, R, seed=23)
    t = synth.make_heatmaps(N, R, 16, seed=24)
    torch.set_num_threads(8)
    with torch.no_grad():
        pm, _ = O.hg_forward_dropout(dict(sd), x, S, asd, training=True, is_half_hg=True)
    np.testing.assert_allclose(pm.numpy(), g["half_pred_mask"], rtol=1e-4, atol=1e-5)
    np.random.seed(4321)
    with torch.no_grad():
        (outs, pm2, indexes, masks), _ = O.hg_forward_dropout(dict(sd), x, S, asd, training=True)
    assert np.array_equal(indexes.numpy(), g["indexes"])
    assert float(masks.sum()) == N * 14 and masks.shape == (N, 1, 4, 4)
    loss = float(O.mse_loss(outs, t))
    assert abs(loss - float(g["loss"])) < 1e-4 * float(g["loss"])
    for i, o in enumerate(outs):
        ref = g["out%d_sub" % i]
        assert np.abs(o.numpy()[:, :, ::3, ::3] - ref).max() < 1e-3 * np.abs(ref).max()


def test_pts2heatmap_oracle_vs_reference_golden():
    """oracle pts2heatmap / draw_gaussian (pylib/HumanPts.py:36-48,82-116) vs the reference's own output."""
    import torch
    g = np.load(os.path.join(ROOT, "tests", "golden", "humanpts_f32.npz"))
    pts = g["pts"]
    for n in range(pts.shape[0]):
        hm, vp = E.pts2heatmap(pts[n].copy(), [64, 64], sigma=1)
        hm32 = torch.from_numpy(hm).float().numpy()
        assert np.array_equal(vp.astype(np.float32), g["valid_pts"][n])
        assert np.array_equal(hm32[:, ::9], g["heatmap_rows"][n])
        np.testing.assert_allclose(hm32.astype(np.float64).sum(axis=(1, 2)), g["heatmap_sum"][n], rtol=1e-12)
    hm2, _ = E.pts2heatmap(pts[0].copy(), [64, 64], sigma=2)
    np.testing.assert_allclose(hm2.sum(axis=(1, 2)), g["heatmap_sigma2_sum"], rtol=1e-12)
