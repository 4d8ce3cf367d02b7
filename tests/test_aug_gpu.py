"""-m gpu parity tests of the crop / rotate / resize augmentation (csrc/warp.cu through the C-ABI and the drop-in
pose_adv_aug_b200/pylib/HumanAug.py::crop / crop_batch; SURVEY 8f row N3, warp half).  Byte work: every comparison is
bit-exact -- against PIL for the two resampling kernels, against oracle/aug_oracle.py::crop on the golden cases (whose
hashes come from the reference's own crop, tests/golden/aug_crop.npz) and on seeded random cases."""
import ctypes
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hgk_testlib import DEV, call, lib                        # noqa: E402
from oracle import aug_oracle as A, synth                     # noqa: E402
from oracle.gen_golden_aug import CASES, case_inputs          # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "aug_crop.npz"))


def _aug():
    from pose_adv_aug_b200.pylib import HumanAug
    return HumanAug


def _resize_gpu(a, oh, ow, y_off=0, x_off=0, in_h=None, in_w=None):
    H = _aug()
    t = torch.from_numpy(a).to(DEV)
    in_h = a.shape[0] - y_off if in_h is None else in_h
    in_w = a.shape[1] - x_off if in_w is None else in_w
    return H._Aug(DEV).resize(t, a.shape[0], a.shape[1], y_off, x_off, in_h, in_w, oh, ow).cpu().numpy()


def test_resize_kernels_equal_pil():
    rng = np.random.default_rng(5)
    for (h, w, oh, ow) in [(300, 400, 256, 256), (700, 333, 256, 256), (100, 90, 256, 256), (720, 1280, 281, 500),
                           (256, 256, 256, 256), (513, 517, 171, 172), (37, 1200, 256, 256), (1080, 1920, 277, 492)]:
        a = (rng.random((h, w, 3)) * 255).astype(np.uint8)
        ref = np.array(Image.fromarray(a).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(ref, _resize_gpu(a, oh, ow)), (h, w, oh, ow)
    # sub-image (the padding removed after a rotation): same as resizing the slice
    a = (rng.random((400, 380, 3)) * 255).astype(np.uint8)
    ref = np.array(Image.fromarray(np.ascontiguousarray(a[41:-41, 41:-41])).resize((256, 256), Image.BILINEAR))
    assert np.array_equal(ref, _resize_gpu(a, 256, 256, 41, 41, 400 - 82, 380 - 82))
    # the tap tables themselves
    for n_in, n_out in [(300, 256), (1280, 500), (90, 256), (256, 256)]:
        ks = lib().aug_resample_ksize(n_in, n_out)
        b, k = A.pil_resample_coeffs(n_in, n_out)
        assert ks == k.shape[1]
        bd = torch.empty(n_out * 2, device=DEV, dtype=torch.int32)
        kd = torch.empty(n_out * ks, device=DEV, dtype=torch.int32)
        call("aug_resample_coeffs", n_in, n_out, ks, bd.data_ptr(), kd.data_ptr())
        assert np.array_equal(bd.cpu().numpy().reshape(n_out, 2), b) and np.array_equal(kd.cpu().numpy().reshape(n_out, ks), k)


def test_rotate_kernel_equals_pil():
    H = _aug()
    rng = np.random.default_rng(6)
    # (the multiples of 90 degrees are PIL's copy / transpose fast paths: the inverse-affine kernel returns the same bytes)
    for (h, w, ang) in [(300, 300, 17.3), (411, 411, -33.0), (200, 260, 5.5), (333, 333, 61.2), (128, 128, -0.7), (90, 64, 200.0),
                        (725, 725, 44.9), (200, 200, 90.0), (200, 260, 90.0), (301, 301, 270.0), (333, 332, 180.0), (200, 260, 360.0),
                        (120, 120, -90.0), (64, 90, 450.0)]:
        a = (rng.random((h, w, 3)) * 255).astype(np.uint8)
        ref = np.array(Image.fromarray(a).rotate(ang, resample=Image.BILINEAR))
        t = torch.from_numpy(a).to(DEV)
        out = torch.empty_like(t)
        m = (ctypes.c_double * 6)(*H._rotate_matrix(ang, w, h))
        call("aug_rotate", t.data_ptr(), h, w, ctypes.cast(m, ctypes.c_void_p).value, out.data_ptr())
        assert np.array_equal(ref, out.cpu().numpy()), (h, w, ang)


def test_byte_scaling_kernels():
    rng = np.random.default_rng(7)
    # whole float32 image (float32 arithmetic)
    a = rng.random((211, 307, 3)).astype(np.float32) * 0.9 + 0.03
    t = torch.from_numpy(a).to(DEV)
    mm = torch.empty(2, device=DEV, dtype=torch.float64)
    scr = torch.tensor([-1, 0, 0], device=DEV, dtype=torch.int32)      # idle state of the reduction scratch
    call("aug_minmax", t.data_ptr(), 0, 211, 307, 0, 211, 0, 307, 0, scr.data_ptr(), mm.data_ptr())
    assert scr.cpu().tolist() == [-1, 0, 0]
    assert mm.cpu().tolist() == [float(a.min()), float(a.max())]
    out = torch.empty(211, 307, 3, device=DEV, dtype=torch.uint8)
    call("aug_image_bytes_f32", t.data_ptr(), 211, 307, mm.data_ptr(), out.data_ptr())
    assert np.array_equal(out.cpu().numpy(), A.bytescale(a))
    # zero-padded window of a float32 and of a uint8 source (float64 arithmetic), region min / max joined with 0
    for src in (a, (a * 255).astype(np.uint8)):
        is_u8 = int(src.dtype == np.uint8)
        ts = torch.from_numpy(src).to(DEV)
        Hn, Wn = 150, 170
        new_y, new_x, old_y, old_x = (20, 150), (0, 160), (0, 130), (147, 307)
        win = np.zeros((Hn, Wn, 3))
        win[new_y[0]:new_y[1], new_x[0]:new_x[1]] = src[old_y[0]:old_y[1], old_x[0]:old_x[1]]
        call("aug_minmax", ts.data_ptr(), is_u8, 211, 307, old_y[0], old_y[1], old_x[0], old_x[1], 1, scr.data_ptr(), mm.data_ptr())
        assert mm.cpu().tolist() == [float(win.min()), float(win.max())]
        wb = torch.empty(Hn, Wn, 3, device=DEV, dtype=torch.uint8)
        call("aug_window_bytes", ts.data_ptr(), is_u8, 211, 307, old_y[0] - new_y[0], old_x[0] - new_x[0], new_y[0], new_y[1],
             new_x[0], new_x[1], mm.data_ptr(), Hn, Wn, wb.data_ptr())
        assert np.array_equal(wb.cpu().numpy(), A.bytescale(win))


def test_crop_equals_reference_golden():
    H = _aug()
    for k in range(len(CASES)):
        img, c, s, r = case_inputs(k)
        out = H.crop(torch.from_numpy(img).to(DEV), c, s, r, 256, 200)
        assert out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (256, 256, 3)
        o = out.cpu().numpy()
        assert np.array_equal(o[::16, ::16], G["sample%d" % k]), "case %d %s" % (k, CASES[k])
        assert hashlib.sha256(o.tobytes()).digest() == G["sha%d" % k].tobytes(), "case %d %s" % (k, CASES[k])


def test_crop_equals_oracle_on_random_cases_and_batch():
    H = _aug()
    rng = np.random.default_rng(321)
    imgs, cs, ss, rs, want = [], [], [], [], []
    for k in range(16):
        h, w = int(rng.integers(240, 900)), int(rng.integers(240, 1300))
        img = synth.make_photo(h, w, 200 + k)
        c = np.array([rng.uniform(0.05 * w, 0.95 * w), rng.uniform(0.05 * h, 0.95 * h)], dtype=np.float32)
        s = np.float32(rng.uniform(0.5, 4.5))
        r = 0.0 if rng.random() < 0.35 else float(rng.uniform(-65, 65))
        ref = A.crop(img, c, s, r, 256, 200)
        out = H.crop(torch.from_numpy(img).to(DEV), c, s, r, 256, 200)
        assert np.array_equal(out.cpu().numpy(), ref), (h, w, c, s, r)
        imgs.append(torch.from_numpy(img).to(DEV)); cs.append(c); ss.append(s); rs.append(r)
        want.append(A.im_to_torch_float(ref))
    # one nearly black image: im_to_torch divides by 255 only when the maximum exceeds 1
    dark = np.zeros((300, 300, 3), dtype=np.float32)
    ref = A.crop(dark, np.array([150, 150], dtype=np.float32), np.float32(1.0), 0, 256, 200)
    assert ref.max() == 0
    imgs.append(torch.from_numpy(dark).to(DEV)); cs.append([150, 150]); ss.append(1.0); rs.append(0.0)
    want.append(A.im_to_torch_float(ref))
    # the batched pipeline (one descriptor table, a dozen launches for the whole batch) and the per-image one: same bytes
    batch = H.crop_batch(imgs, np.array(cs), np.array(ss), np.array(rs), 256, 200)
    assert tuple(batch.shape) == (17, 3, 256, 256) and batch.dtype == torch.float32
    assert np.array_equal(batch.cpu().numpy(), np.stack(want))
    assert torch.equal(batch, H.crop_batch(imgs, np.array(cs), np.array(ss), np.array(rs), 256, 200, batched=False))
    assert torch.equal(batch, H.crop_batch(imgs, np.array(cs), np.array(ss), np.array(rs), 256, 200))       # scratch re-armed
    # a batch without any pre-shrink / rotation, and the golden cases as one batch
    sub = [k for k in range(17) if float(ss[k]) * 200 / 256 < 2 and rs[k] == 0.0]
    if sub:
        b2 = H.crop_batch([imgs[k] for k in sub], np.array([cs[k] for k in sub]), np.array([ss[k] for k in sub]),
                          np.array([rs[k] for k in sub]), 256, 200)
        assert np.array_equal(b2.cpu().numpy(), np.stack([want[k] for k in sub]))
    gi = [case_inputs(k) for k in range(len(CASES))]
    gb = H._crop_batch_u8([torch.from_numpy(g[0]).to(DEV) for g in gi], np.stack([g[1] for g in gi]),
                          np.array([g[2][0] for g in gi], dtype=np.float32), np.array([g[3] for g in gi], dtype=np.float64), 256, 200)
    for k in range(len(CASES)):
        assert hashlib.sha256(gb[k].cpu().numpy().tobytes()).digest() == G["sha%d" % k].tobytes(), "batched, case %d" % k


def test_crop_error_behaviour():
    H = _aug()
    from pose_adv_aug_b200 import HGKError
    img = synth.make_photo(300, 400, 1)
    with pytest.raises(HGKError):
        H.crop(torch.from_numpy(img), [200, 150], 1.0, 0, 256, 200)                # CPU tensor: no fallback
    t = torch.from_numpy(img).to(DEV)
    with pytest.raises(ValueError):
        H.crop(t, [5000, 5000], 1.0, 0, 256, 200)                                  # window off the image (numpy raises too)
    with pytest.raises(ValueError):
        H.crop(t.double(), [200, 150], 1.0, 0, 256, 200)
    # rotations by multiples of 90 degrees (PIL's fast paths) give the oracle's bytes too
    for r in (90.0, 180.0, -90.0, 360.0):
        c, s_ = np.array([200.0, 150.0], dtype=np.float32), np.float32(0.9)
        assert np.array_equal(H.crop(t, c, s_, r, 256, 200).cpu().numpy(), A.crop(img, c, s_, r, 256, 200)), r
    tiny = torch.from_numpy(synth.make_photo(8, 8, 3)).to(DEV)
    assert H.crop(tiny, [4, 4], 12.0, 0, 256, 200) is tiny                         # degenerate early return (ref :128-129)


def _annos(n, rng, sizes):
    out = []
    for k in range(n):
        h, w = sizes[k]
        pts = np.concatenate([rng.uniform(0.2 * w, 0.8 * w, (16, 1)), rng.uniform(0.2 * h, 0.8 * h, (16, 1)), np.ones((16, 1))], axis=1)
        pts[3, :2] = 0                       # an invisible joint
        pts[7] = [12.0, 33.0, 1]             # integer coordinates
        out.append({"joint_self": pts.tolist(), "objpos": [float(rng.uniform(0.3 * w, 0.7 * w)), float(rng.uniform(0.3 * h, 0.7 * h))],
                    "scale_provided": float(rng.uniform(0.7, 3.0)), "normalizer": float(rng.uniform(40, 120))})
    return out


def test_agent_batch_loader_equals_oracle():
    """load_batch_data without the DataLoader (agent.AgentBatchLoader): same random stream, same bytes as the restated
    AGENT.__getitem__ + gen_img_heatmap of the oracle -- images, heat-maps and every host-side quantity."""
    from pose_adv_aug_b200.agent import AgentBatchLoader
    rng = np.random.default_rng(77)
    sizes = [(int(rng.integers(300, 760)), int(rng.integers(400, 1300))) for _ in range(6)]
    photos = [np.ascontiguousarray(np.transpose(synth.make_photo(h, w, 400 + k), (2, 0, 1))) for k, (h, w) in enumerate(sizes)]
    annos = _annos(6, rng, sizes)
    img_index = [4, 0, 5, 2, 2, 1, 3, 0]
    si = [int(v) for v in rng.integers(0, 7, len(img_index))]
    ri = [int(v) for v in rng.integers(0, 7, len(img_index))]
    want = A.agent_batch(photos, annos, si, ri, img_index, np.random.RandomState(2024))
    loader = AgentBatchLoader([torch.from_numpy(p).to(DEV) for p in photos], annos)
    got = loader.load_batch(si, ri, torch.tensor(img_index), rng=np.random.RandomState(2024))
    assert got[0].is_cuda and got[1].is_cuda
    assert np.array_equal(got[0].cpu().numpy(), want[0])                      # images: bit-exact
    assert np.array_equal(got[1].cpu().numpy(), want[1])                      # heat-maps: bit-exact
    assert np.array_equal(got[2].numpy(), want[2]) and np.array_equal(got[3].numpy().reshape(-1), want[3])
    assert np.array_equal(got[4].numpy().reshape(-1), want[4].astype(np.float32))
    assert np.array_equal(got[5].numpy(), want[5]) and np.array_equal(got[6].numpy(), want[6])
    assert want[1].reshape(len(img_index), 16, -1).max(axis=2).min() == 0 and want[1].max() == 1      # blobs and an absent joint


def test_joint_train_iteration_vs_oracle():
    """One agent-augmented joint-train iteration end to end (ref joint-train-pose-s-r-agent.py:246-299): half-hourglass
    forward (hg.train(), agent.eval()) -> ASN scale / rotation distributions -> sampling -> the batch re-augmented with the
    sampled bins from resident photographs (load_batch_data) -> full train step -> PCK -- every stage against the oracle
    chain hg_oracle / eval_oracle / aug_oracle fed with the same random numbers."""
    from collections import OrderedDict
    from oracle import hg_oracle as O, eval_oracle as E
    from pose_adv_aug_b200 import HourglassTrainer, agent
    from pose_adv_aug_b200.models import asn_stacked_hg as M
    from pose_adv_aug_b200.pylib import Evaluation
    S, C, N, R = 2, 32, 3, 256
    sd = synth.make_state_dict(O.hg_schema(S, 1, 16, C), seed=301)
    asd = synth.make_state_dict(O.asn_schema(C, C, 7, 7, is_aug=True), seed=302)
    net = M.create_hg(S, 1, 16, C)
    net.load_state_dict(sd)
    asn = M.create_asn(C, C, 7, 7, is_aug=True)
    asn.load_state_dict(asd)
    net.to(DEV), asn.to(DEV)
    rng = np.random.default_rng(303)
    sizes = [(int(rng.integers(400, 700)), int(rng.integers(500, 900))) for _ in range(N)]
    photos = [np.ascontiguousarray(np.transpose(synth.make_photo(h, w, 310 + k), (2, 0, 1))) for k, (h, w) in enumerate(sizes)]
    annos = _annos(N, rng, sizes)
    loader = agent.AgentBatchLoader([torch.from_numpy(p).to(DEV) for p in photos], annos)
    img_index = list(range(N))
    x0 = synth.make_images(N, R, seed=304)
    # 1. half-hourglass + ASN forward (:250-251)
    net.train(); asn.eval()
    with torch.no_grad():
        ps, pr = net(x0.to(DEV), asn, is_half_hg=True, is_aug=True)
    (ps_o, pr_o), _ = O.hg_forward(OrderedDict((k, v.clone()) for k, v in sd.items()), x0, S, training=True, asn_sd=asd, is_half_hg=True)
    assert float((ps.cpu() - ps_o).abs().max() / ps_o.abs().max()) < 1e-3 and float((pr.cpu() - pr_o).abs().max() / pr_o.abs().max()) < 1e-3
    # 2. sampling (:252-271): the same uniforms give the same bins
    u = torch.from_numpy(np.random.RandomState(305).random_sample(2 * N).reshape(N, 2))
    _, _, si, ri = agent.sample_scale_rotation(ps, pr, uniforms=u)
    si_o = E.sample_agent(ps_o.detach().numpy(), u[:, 0].numpy())[1]
    ri_o = E.sample_agent(pr_o.detach().numpy(), u[:, 1].numpy())[1]
    assert si.cpu().tolist() == si_o.tolist() and ri.cpu().tolist() == ri_o.tolist()
    # 3. the re-augmented batch (load_batch_data, :425-450): byte-exact images and targets
    got = loader.load_batch(si.tolist(), ri.tolist(), img_index, rng=np.random.RandomState(306))
    want = A.agent_batch(photos, annos, si_o.tolist(), ri_o.tolist(), img_index, np.random.RandomState(306))
    assert np.array_equal(got[0].cpu().numpy(), want[0]) and np.array_equal(got[1].cpu().numpy(), want[1])
    # 4. the train step on it (:273-289) and 5. PCK of the last stack (:291-296)
    tr = HourglassTrainer(net, N, R, use_graph=True)
    loss = float(tr.step(got[0], got[1]))
    outs_o, loss_o, _, _ = O.train_step(OrderedDict((k, v.clone()) for k, v in sd.items()), torch.from_numpy(want[0]),
                                        torch.from_numpy(want[1]), S, 1)
    assert abs(loss - float(loss_o)) < 1e-3 * abs(float(loss_o)), (loss, float(loss_o))
    for h, o in zip(tr.heatmaps(), outs_o):
        assert float((h.cpu() - o).abs().max() / o.abs().max()) < 1e-3
    idx = list(range(16))
    acc = Evaluation.accuracy(tr.heatmaps()[-1], got[1], idx)
    acc_o = E.accuracy(outs_o[-1].detach().numpy(), want[1], idx)
    assert abs(float(acc[0]) - float(acc_o[0])) <= 0.1
