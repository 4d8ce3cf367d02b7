"""The heat-map <-> point conversions of the reference `pylib/HumanPts.py`, on the GPU.

  pts2heatmap(pts, heatmap_shape, sigma=1)    ref pylib/HumanPts.py:36-48 (+ draw_gaussian :82-116)
  heatmap2pts(heatmap)                        ref pylib/HumanPts.py:118-137

The reference renders the ground-truth heat-maps on the host inside every dataset `__getitem__`
(data/mpii_for_mpii.py:151, data/joint_train_pose.py:172 ...) and ships them H2D (1.57 M floats per batch of 24).
Here the [N,J,2] points are the only thing that crosses PCIe (3 KB) and one CTA per map writes zeros + the blob
(`hgk_pts2heatmap`).  The blob table itself is computed on the host with numpy exactly as `draw_gaussian` does
(fp64 exp, then the reference's `.float()`), so the rendered maps are bit-identical to the reference's.
File I/O / PIL drawing helpers of the reference file are out of scope.
"""
import numpy as np
import torch

from .._lib import get_lib, HGKError


def gaussian_blob(sigma=1):
    """The `g` of draw_gaussian (ref:92-99): size = 2*ceil(3 sigma)+1, exp(-((x-x0)^2+(y-y0)^2)/tmp_size^2), fp64 -> fp32."""
    tmp_size = np.ceil(3 * sigma)
    size = 2 * tmp_size + 1
    x = np.arange(0, size, 1, float)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2
    g = np.exp(- ((x - x0) ** 2 + (y - y0) ** 2) / (tmp_size ** 2))
    return torch.from_numpy(g).float()


_blobs = {}


def pts2heatmap(pts, heatmap_shape, sigma=1):
    """pts: CUDA float tensor [J,2] (one sample, as the reference) or [N,J,2]; returns (heatmap [.., J, H, W], valid_pts)
    as CUDA float tensors."""
    if not isinstance(pts, torch.Tensor) or not pts.is_cuda:
        raise HGKError("HumanPts.pts2heatmap runs on CUDA tensors only (no CPU fallback)")
    lib = get_lib()
    p = pts.contiguous().float()
    if p.dim() not in (2, 3) or p.shape[-1] != 2:
        raise ValueError("pts must be [J,2] or [N,J,2]")
    H, W = int(heatmap_shape[0]), int(heatmap_shape[1])
    key = (float(sigma), p.device)
    if key not in _blobs:
        _blobs[key] = gaussian_blob(sigma).to(p.device)
    g = _blobs[key]
    M = p.numel() // 2
    heat = torch.empty(tuple(p.shape[:-1]) + (H, W), device=p.device, dtype=torch.float32)
    valid = torch.empty_like(p)
    lib.check(lib.pts2heatmap(p.data_ptr(), M, H, W, g.data_ptr(), g.shape[0], heat.data_ptr(), valid.data_ptr(),
                              torch.cuda.current_stream(p.device).cuda_stream), "hgk_pts2heatmap")
    return heat, valid


def heatmap2pts(heatmap):
    """ref:118-137 -- [B,J,H,W] -> [B,J,2]: (idx % W, floor(idx / W) + 0.5) of the maximum, zero where max <= 0."""
    if not isinstance(heatmap, torch.Tensor) or not heatmap.is_cuda:
        raise HGKError("HumanPts.heatmap2pts runs on CUDA tensors only (no CPU fallback)")
    lib = get_lib()
    h = heatmap.contiguous().float()
    B, J, H, W = h.shape
    pts = torch.empty(B, J, 2, device=h.device, dtype=torch.float32)
    lib.check(lib.heatmap_peaks(h.data_ptr(), B, J, H, W, 2, 0, 0, 0, pts.data_ptr(), 0,
                                torch.cuda.current_stream(h.device).cuda_stream), "hgk_heatmap_peaks")
    return pts
