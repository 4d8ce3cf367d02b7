"""Drop-in for the reference `pylib/Evaluation.py` (same function names, argument meaning and return
values) with the heat-map work on the GPU.

The reference pulls `output[-1]` to the host every iteration (stack-hg.py:176-179, joint-train-pose-s-r-agent.py
:242,296: 6.3 MB D2H + a sync) and walks it with python double loops.  Here the heat-maps stay in HBM: one CTA per
(sample, joint) finds the peak (`hgk_heatmap_peaks`), one small kernel forms the distances and PCK accuracies
(`hgk_pck_accuracy`); the results are CUDA tensors of a few bytes (`acc[0].item()` is the only sync a caller needs).
The crop transforms (3x3 fp64 matrices per sample) are built on the host with numpy exactly as the reference does.

`ref:` = /root/reference/pylib/Evaluation.py.
"""
import numpy as np
import torch

from .._lib import get_lib, HGKError

__all__ = ["get_preds", "calc_dists", "dist_acc", "accuracy", "accuracy_origin_res", "per_person_pckh",
           "final_preds", "GetTransform", "TransformPts", "transform_preds"]

MPII_IDXS = [0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13, 14, 15]       # ref:86,109


def _cuda_f32(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HGKError("Evaluation.%s runs on CUDA tensors only (no CPU fallback); got %s"
                       % (what, t.device if isinstance(t, torch.Tensor) else type(t)))
    return t.contiguous().float()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _peaks(scores, mode, res=(0, 0), tinv=None):
    lib = get_lib()
    s = _cuda_f32(scores, "get_preds")
    assert s.dim() == 4, 'Score maps should be 4-dim'                # ref:10
    N, J, H, W = s.shape
    preds = torch.empty(N, J, 2, device=s.device, dtype=torch.float32)
    lib.check(lib.heatmap_peaks(s.data_ptr(), N, J, H, W, mode, int(res[0]), int(res[1]),
                                0 if tinv is None else tinv.data_ptr(), preds.data_ptr(), 0, _stream(s)),
              "hgk_heatmap_peaks")
    return preds


def get_preds(scores):
    """ref:6-23 -- [N,J,H,W] score maps -> float [N,J,2] 1-based (x, y) peak locations, (0,0) where max <= 0."""
    return _peaks(scores, 0)


def GetTransform(center, scale, rot, res, size):
    """Crop transform (original image -> res x res crop) as a 3x3 homogeneous matrix (host, fp64); the map of ref:211-237 in
    closed form: a similarity p -> k*(p - center) + res/2 with k = res / (size*scale), followed -- when rot != 0 -- by a
    rotation of -rot degrees about the crop centre (res/2, res/2).  With q = k*(p - center) the crop-centred point, the whole
    map is  p -> R(-rot) q + res/2,  i.e. linear part k*R and translation R*(k*(-center) ... ) + res/2."""
    h = size * scale
    k = float(res) / h
    half = res / 2
    # crop-centred image of the origin: k*(0 - center) = res*(-center/h + 1/2) - res/2
    ox = res * (-float(center[0]) / h + .5)
    oy = res * (-float(center[1]) / h + .5)
    if rot == 0:
        return np.array([[k, 0., ox], [0., k, oy], [0., 0., 1.]])
    ang = -rot * np.pi / 180
    sn, cs = np.sin(ang), np.cos(ang)
    qx, qy = ox - half, oy - half
    return np.array([[cs * k, -sn * k, (cs * qx - sn * qy) + half],
                     [sn * k, cs * k, (sn * qx + cs * qy) + half],
                     [0., 0., 1.]])


def TransformPts(pts, center, scale, rot, res, size, invert=0):
    """ref:239-247 -- 1-based [M,2] points through the crop transform (or its inverse), truncated to integers as the
    reference does (host numpy; kept for callers that transform ground-truth points)."""
    t = GetTransform(center, scale, rot, res, size)
    if invert:
        t = np.linalg.inv(t)
    p0 = np.asarray(pts, dtype=np.float64) - 1
    mapped = (p0[:, 0:1] * t[:2, 0] + p0[:, 1:2] * t[:2, 1]) + t[:2, 2]
    return mapped.astype(int) + 1


def transform_preds(coords, center, scale, res, rot):
    """ref:195-209 (host)."""
    c = TransformPts(coords.detach().cpu().numpy(), _np(center), float(_np(scale)), float(_np(rot)), res[0], size=200, invert=1)
    return torch.from_numpy(c)


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def final_preds(output, center, scale, res, rot):
    """ref:169-193 -- peaks + quarter-pixel shift + 0.5, mapped back through the inverse crop transform.
    center [N,2], scale [N], rot [N]: host (or CUDA) tensors as the data loader yields them."""
    out = _cuda_f32(output, "final_preds")
    N = out.shape[0]
    c, s, r = _np(center), _np(scale).reshape(-1), _np(rot).reshape(-1)
    tinv = np.empty((N, 6), dtype=np.float64)
    for i in range(N):
        t = np.linalg.inv(GetTransform(c[i], float(s[i]), float(r[i]), res[0], 200))
        tinv[i] = t[:2].reshape(-1)
    tinv_d = torch.from_numpy(tinv).to(out.device, non_blocking=True)
    return _peaks(out, 1, res, tinv_d)


def _dists(preds, target, normalize, boundary, thr=0.5, idxs=None):
    lib = get_lib()
    p = _cuda_f32(preds, "calc_dists")
    dev = p.device
    t = target.to(dev).contiguous().float() if isinstance(target, torch.Tensor) else torch.as_tensor(target, dtype=torch.float32, device=dev)
    nrm = normalize.to(dev).contiguous().float() if isinstance(normalize, torch.Tensor) else torch.as_tensor(normalize, dtype=torch.float32, device=dev)
    N, J = p.shape[0], p.shape[1]
    if tuple(t.shape) != (N, J, 2) or nrm.numel() != N:
        raise ValueError("calc_dists: preds %s, target %s, normalize %s do not match" % (tuple(p.shape), tuple(t.shape), tuple(nrm.shape)))
    dists = torch.empty(J, N, device=dev, dtype=torch.float32)
    acc, idx_d = None, None
    if idxs is not None:
        idx_list = [int(i) for i in (idxs.tolist() if isinstance(idxs, torch.Tensor) else idxs)]
        if any(i < 0 or i >= J for i in idx_list):
            raise IndexError("joint index out of range")
        idx_d = torch.tensor(idx_list, dtype=torch.int32, device=dev)
        acc = torch.zeros(len(idx_list) + 1, device=dev, dtype=torch.float32)
    lib.check(lib.pck_accuracy(p.data_ptr(), t.data_ptr(), nrm.data_ptr(), N, J, float(boundary), float(thr),
                               0 if idx_d is None else idx_d.data_ptr(), 0 if idx_d is None else idx_d.numel(),
                               dists.data_ptr(), 0 if acc is None else acc.data_ptr(), _stream(p)), "hgk_pck_accuracy")
    return dists, acc


def calc_dists(preds, target, normalize, use_zero=False):
    """ref:25-39 -> [J,N]: |pred - target| / normalize[n] where both target coordinates > boundary, else -1."""
    return _dists(preds, target, normalize, 0 if use_zero else 1)[0]


def dist_acc(dists, thr=0.5):
    """ref:41-54 -> 0-dim CUDA tensor (share of valid distances <= thr; -1 when none is valid)."""
    lib = get_lib()
    d = _cuda_f32(dists, "dist_acc").reshape(-1)
    out = torch.empty((), device=d.device, dtype=torch.float32)
    lib.check(lib.dist_acc(d.data_ptr(), d.numel(), float(thr), out.data_ptr(), _stream(d)), "hgk_dist_acc")
    return out


def accuracy(output, target, idxs, thr=0.5):
    """ref:56-80 -- PCK between the peaks of `output` and of the ground-truth heat-maps `target`, normalised by W/10.
    Returns a CUDA float tensor [len(idxs)+1]: [0] the average over joints with a valid value, then per joint."""
    preds = get_preds(output)
    gts = get_preds(target.to(output.device) if isinstance(target, torch.Tensor) else target)
    norm = torch.full((preds.size(0),), output.size(3) / 10.0, device=preds.device, dtype=torch.float32)
    return _dists(preds, gts, norm, 1, thr, idxs)[1]


def accuracy_origin_res(output, center, scale, res, grnd_pts, normalizers, rot):
    """ref:82-104 -- PCKh in original image coordinates."""
    pred_pts = final_preds(output, center, scale, res, rot)
    return _dists(pred_pts, grnd_pts, normalizers, 0, 0.5, MPII_IDXS)[1]


def per_person_pckh(output, grnd_heatmap, center, scale, res, grnd_pts, normalizers, rot, thr=0.5):
    """ref:106-167 -> CUDA float [N] per-sample PCKh over the 14 evaluated joints."""
    lib = get_lib()
    pred_pts = final_preds(output, center, scale, res, rot)
    dists = _dists(pred_pts, grnd_pts, normalizers, 0)[0]
    gt_preds = get_preds(grnd_heatmap.to(pred_pts.device))
    N, J = pred_pts.shape[0], pred_pts.shape[1]
    idx_d = torch.tensor(MPII_IDXS, dtype=torch.int32, device=pred_pts.device)
    acc = torch.empty(N, device=pred_pts.device, dtype=torch.float32)
    lib.check(lib.per_person_pckh(dists.data_ptr(), gt_preds.data_ptr(), N, J, idx_d.data_ptr(), idx_d.numel(), float(thr),
                                  acc.data_ptr(), _stream(acc)), "hgk_per_person_pckh")
    return acc
