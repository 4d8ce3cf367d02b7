"""The part of the reference `pylib/HumanAug.py` that sits on the validation path next to the network
(stack-hg.py:225-232): horizontal-flip test-time augmentation of the heat-maps, on the GPU.

  flip_channels(maps)                                ref pylib/HumanAug.py:199-210
  shuffle_channels_for_horizontal_flipping(maps)     ref pylib/HumanAug.py:179-197
  flip_merge(output1, output2)                       = (output1 + shuffle(flip(output2))) / 2 in ONE kernel

The reference does the flip with numpy on a host copy of the output; here the W-flip, the six left/right channel
swaps and the mean are the index arithmetic of a single pass (`hgk_flip_merge_nchw`).

and its image crop (the warp half of SURVEY 8f row N3): the function `load_batch_data`
(joint-train-pose-s-r-agent.py:425-450) reaches through a DataLoader for every batch of agent-sampled scales / rotations:

  crop(img, center, scale, rot, res, size)           ref pylib/HumanAug.py:117-175      (one image, uint8 H x W x 3 out)
  crop_batch(imgs, centers, scales, rots, res, size) = im_to_torch(crop(...)) per sample, stacked [N,3,res,res] float32
                                                       (ref data/joint_train_s_r_agent.py:198-204, utils/imutils.py:31-36)

Source images stay resident on the GPU as float32 H x W x 3 tensors in [0,1] (what `im_to_numpy(load_image(...))` hands
the reference); the pixel work -- scipy.misc's byte-scaling, PIL's two-pass bilinear resize and bilinear rotation -- runs as
byte-exact CUDA kernels (`csrc/warp.cu`), the crop geometry (a handful of scalars per image) on the host exactly as the
reference computes it with numpy.  No CPU fallback: CPU tensors raise.
"""
import ctypes
import math

import numpy as np
import torch

from .._lib import get_lib, HGKError

FLIP_PAIRS = [[1, 4], [0, 5], [12, 13], [11, 14], [10, 15], [2, 3]]      # ref:182


def _check(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HGKError("HumanAug.%s runs on CUDA tensors only (no CPU fallback)" % what)
    if t.dim() not in (3, 4):
        raise ValueError('tensor dimension is not right')                 # ref:188,206
    return t.contiguous().float()


def _launch(a, b, pairs, flip_w):
    """out = a is None ? f(b) : (a + f(b)) / 2,  f = channel swaps after an optional W-flip."""
    lib = get_lib()
    squeeze = b.dim() == 3
    b4 = b.unsqueeze(0) if squeeze else b
    a4 = None if a is None else (a.unsqueeze(0) if squeeze else a)
    if a4 is not None and a4.shape != b4.shape:
        raise ValueError("flip merge: shapes differ (%s vs %s)" % (tuple(a4.shape), tuple(b4.shape)))
    N, C, H, W = b4.shape
    flat = [int(v) for p in pairs for v in p]
    arr = (ctypes.c_int * max(len(flat), 1))(*flat)                      # host array, read during the call only
    out = torch.empty_like(b4)
    lib.check(lib.flip_merge_nchw(0 if a4 is None else a4.data_ptr(), b4.data_ptr(), N, C, H, W,
                                  ctypes.cast(arr, ctypes.c_void_p).value, len(pairs), int(flip_w), out.data_ptr(),
                                  torch.cuda.current_stream(b.device).cuda_stream), "hgk_flip_merge_nchw")
    return out[0] if squeeze else out


def flip_merge(output1, output2, pairs=FLIP_PAIRS):
    """(output1 + shuffle_channels_for_horizontal_flipping(flip_channels(output2))) / 2, stack-hg.py:229-232.
    output1: heat-maps of the images, output2: heat-maps of the W-flipped images; NCHW CUDA tensors."""
    return _launch(_check(output1, "flip_merge"), _check(output2, "flip_merge"), pairs, True)


def flip_channels(maps):
    """ref:199-210 -- horizontally flip every channel (returns a new float tensor)."""
    return _launch(None, _check(maps, "flip_channels"), [], True)


def shuffle_channels_for_horizontal_flipping(maps):
    """ref:179-197 -- swap the left/right joint channels IN PLACE (as the reference does) and return `maps`."""
    out = _launch(None, _check(maps, "shuffle_channels_for_horizontal_flipping"), FLIP_PAIRS, False)
    maps.copy_(out)
    return maps


# ---------------------------------------------------------------------------------------------------------------------
# crop (ref :117-175)
# ---------------------------------------------------------------------------------------------------------------------
def _window(ht, wd, center, scale, rot, res, size):
    """The integer crop window of ref :121-162 for an ht x wd source: (scale_factor, pre-shrunk size or None, window size,
    pasted range in the window, matching range in the source, rotation padding)."""
    f32 = np.float32
    c = np.asarray(center, dtype=np.float32).reshape(2)                   # the dataset hands float32 tensors (.numpy())
    s = f32(np.asarray(scale, dtype=np.float32).reshape(-1)[0])
    sf = float(s * f32(size)) / float(res)                                # :121
    pre = None
    if sf < 2:
        sf = 1
    else:
        if np.floor(max(ht, wd) / sf) < 2:                                # :126-129
            return None
        wh = (np.array([wd, ht]) * (1 / sf)).astype(int)                  # imresize(size=float): PIL size, truncated
        pre = (int(wh[1]), int(wh[0]))
        ht, wd = pre
    c = c / f32(sf)
    s = s / f32(sf)
    # GetTransform(center, scale, 0, res, size) (:10-21) has float32 entries; TransformSinglePts inverts it numerically and
    # truncates the mapped corners (:38-44)
    h = f32(size) * s
    t = np.zeros((3, 3))
    t[0, 0] = t[1, 1] = f32(res) / h
    t[0, 2] = f32(res) * (f32(-float(c[0])) / h + f32(.5))
    t[1, 2] = f32(res) * (f32(-float(c[1])) / h + f32(.5))
    t[2, 2] = 1
    inv = np.linalg.inv(t)
    ul = np.dot(inv, np.array([0, 0, 1.]))[:2].astype(int)
    br = np.dot(inv, np.array([res, res, 1.]))[:2].astype(int)
    if sf >= 2:
        br = br - (br - ul - res)                                         # :142-143
    pad = int(np.ceil(np.linalg.norm(br - ul) / 2 - float(br[1] - ul[1]) / 2))     # :146
    if not rot == 0:
        ul = ul - pad
        br = br + pad
    ulx, uly, brx, bry = int(ul[0]), int(ul[1]), int(br[0]), int(br[1])
    new_x = (max(0, -ulx), min(brx, wd) - ulx)
    new_y = (max(0, -uly), min(bry, ht) - uly)
    old_x = (max(0, ulx), min(wd, brx))
    old_y = (max(0, uly), min(ht, bry))
    return sf, pre, (bry - uly, brx - ulx), (new_y, new_x), (old_y, old_x), pad


def _rotate_matrix(angle, w, h):
    """The affine coefficients PIL's Image.rotate(angle) (expand = 0, centre = image centre) hands to its transform."""
    a = -math.radians(angle % 360.0)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    cx, cy = w / 2.0, h / 2.0
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return m


def GetTransform(center, scale, rot, res, size):
    """ref :10-36 -- image -> crop transform (3 x 3, float64 storage).  center / scale arrive as float32 arrays, so the
    similarity entries are float32 arithmetic; the rotation about the crop centre is composed in float64."""
    f32 = np.float32
    c = np.asarray(center, dtype=np.float32).reshape(2)
    h = f32(size) * f32(np.asarray(scale, dtype=np.float32).reshape(-1)[0])
    t = np.zeros((3, 3))
    t[0, 0] = t[1, 1] = f32(res) / h
    t[0, 2] = f32(res) * (f32(-float(c[0])) / h + f32(.5))
    t[1, 2] = f32(res) * (f32(-float(c[1])) / h + f32(.5))
    t[2, 2] = 1
    if not rot == 0:
        a = -rot * np.pi / 180                                # "to match direction of rotation from cropping"
        sn, cs = np.sin(a), np.cos(a)
        rot_mat = np.array([[cs, -sn, 0.], [sn, cs, 0.], [0., 0., 1.]])
        shift = np.eye(3)
        shift[0, 2] = shift[1, 2] = -res / 2
        back = shift.copy()
        back[:2, 2] *= -1
        t = np.dot(back, np.dot(rot_mat, np.dot(shift, t)))
    return t


def TransformPts(pts, center, scale, rot, res, size, invert=0):
    """ref :46-55 -- [M,2] points through the crop transform (or its inverse); float result, no truncation."""
    t = GetTransform(center, scale, rot, res, size)
    if invert:
        t = np.linalg.inv(t)
    p = np.asarray(pts)
    homog = np.concatenate((p, np.ones((p.shape[0], 1))), axis=1).T
    return np.dot(t, homog)[0:2, :].T


class _Aug(object):
    """Launch helper bound to one device / stream."""

    _scratch = {}          # per device: the three idle words of hgk_aug_minmax (stream-ordered re-use)

    def __init__(self, device):
        self.lib = get_lib()
        self.dev = device
        self.stream = torch.cuda.current_stream(device).cuda_stream
        key = (device.type, device.index)
        if key not in _Aug._scratch:
            _Aug._scratch[key] = torch.tensor([-1, 0, 0], device=device, dtype=torch.int32)
        self.scratch = _Aug._scratch[key]

    def u8(self, *shape):
        return torch.empty(shape, device=self.dev, dtype=torch.uint8)

    def minmax(self, img, is_u8, H, W, ys, xs, include_zero):
        out = torch.empty(2, device=self.dev, dtype=torch.float64)
        self.lib.check(self.lib.aug_minmax(img.data_ptr(), int(is_u8), H, W, ys[0], ys[1], xs[0], xs[1], int(include_zero),
                                           self.scratch.data_ptr(), out.data_ptr(), self.stream), "hgk_aug_minmax")
        return out

    def resize(self, src, SH, SW, y_off, x_off, in_h, in_w, out_h, out_w):
        """PIL Image.resize((out_w, out_h), BILINEAR) of the in_h x in_w sub-image of the uint8 image `src`."""
        lib = self.lib
        tabs = []
        for n_in, n_out in ((in_w, out_w), (in_h, out_h)):
            ks = lib.aug_resample_ksize(n_in, n_out)
            bounds = torch.empty(n_out * 2, device=self.dev, dtype=torch.int32)
            kk = torch.empty(n_out * ks, device=self.dev, dtype=torch.int32)
            lib.check(lib.aug_resample_coeffs(n_in, n_out, ks, bounds.data_ptr(), kk.data_ptr(), self.stream),
                      "hgk_aug_resample_coeffs")
            tabs.append((ks, bounds, kk))
        tmp = self.u8(in_h, out_w, 3)
        out = self.u8(out_h, out_w, 3)
        (kh, bh, kkh), (kv, bv, kkv) = tabs
        lib.check(lib.aug_resize_h(src.data_ptr(), SH, SW, y_off, x_off, in_h, in_w, out_w, bh.data_ptr(), kkh.data_ptr(), kh,
                                   tmp.data_ptr(), self.stream), "hgk_aug_resize_h")
        lib.check(lib.aug_resize_v(tmp.data_ptr(), in_h, out_w, out_h, bv.data_ptr(), kkv.data_ptr(), kv, out.data_ptr(),
                                   self.stream), "hgk_aug_resize_v")
        return out


def crop(img, center, scale, rot, res, size):
    """ref :117-175.  img: float32 CUDA tensor H x W x 3 in [0,1]; center (x, y), scale, rot (degrees): host numbers.
    Returns the res x res x 3 uint8 CUDA tensor the reference returns as a numpy array (or `img` itself in the degenerate
    early return of ref :128-129)."""
    if not isinstance(img, torch.Tensor) or not img.is_cuda:
        raise HGKError("HumanAug.crop runs on CUDA tensors only (no CPU fallback)")
    if img.dim() != 3 or img.shape[2] != 3 or img.dtype != torch.float32:
        raise ValueError("crop: expected a float32 H x W x 3 image, got %s %s" % (tuple(img.shape), img.dtype))
    img = img.contiguous()
    H, W = int(img.shape[0]), int(img.shape[1])
    if min(H, W) <= 4:
        raise ValueError("crop: image too small (scipy.misc.toimage would take a 3- or 4-pixel side for the channel axis)")
    rot = float(np.asarray(rot).reshape(-1)[0])
    win = _window(H, W, center, scale, rot, res, size)
    if win is None:
        return img
    sf, pre, (Hn, Wn), (new_y, new_x), (old_y, old_x), pad = win
    if Hn <= 4 or Wn <= 4 or old_y[1] <= old_y[0] or old_x[1] <= old_x[0] or \
            (new_y[1] - new_y[0], new_x[1] - new_x[0]) != (old_y[1] - old_y[0], old_x[1] - old_x[0]):
        # numpy would refuse the slice assignment of ref :164 (window entirely off the image)
        raise ValueError("crop: the crop window does not intersect the image")
    A = _Aug(img.device)
    lib, st = A.lib, A.stream
    src, src_u8, SH, SW = img, 0, H, W
    if pre is not None:                                                    # :131 imresize(img, 1 / scale_factor)
        mm = A.minmax(img, 0, H, W, (0, H), (0, W), False)
        b = A.u8(H, W, 3)
        lib.check(lib.aug_image_bytes_f32(img.data_ptr(), H, W, mm.data_ptr(), b.data_ptr(), st), "hgk_aug_image_bytes_f32")
        src = A.resize(b, H, W, 0, 0, H, W, pre[0], pre[1])
        src_u8, SH, SW = 1, pre[0], pre[1]
    # new_img = zeros; new_img[new] = img[old] (:156-164), then toimage's byte-scaling of that float64 array
    has_zero = (new_y[1] - new_y[0]) < Hn or (new_x[1] - new_x[0]) < Wn
    mm = A.minmax(src, src_u8, SH, SW, old_y, old_x, has_zero)
    wb = A.u8(Hn, Wn, 3)
    lib.check(lib.aug_window_bytes(src.data_ptr(), src_u8, SH, SW, old_y[0] - new_y[0], old_x[0] - new_x[0], new_y[0], new_y[1],
                                   new_x[0], new_x[1], mm.data_ptr(), Hn, Wn, wb.data_ptr(), st), "hgk_aug_window_bytes")
    off = 0
    if not rot == 0:                                                       # :166-170 imrotate + padding removed
        # (multiples of 90 degrees: PIL takes copy / transpose fast paths there; the generic inverse-affine kernel lands on
        #  exact pixel centres for them -- dx = dy = 0 -- and returns the same bytes, tests/test_aug_gpu.py)
        if pad <= 0 or Hn - 2 * pad <= 0 or Wn - 2 * pad <= 0:
            raise ValueError("crop: empty image after removing the rotation padding")
        m = (ctypes.c_double * 6)(*_rotate_matrix(rot, Wn, Hn))
        rb = A.u8(Hn, Wn, 3)
        lib.check(lib.aug_rotate(wb.data_ptr(), Hn, Wn, ctypes.cast(m, ctypes.c_void_p).value, rb.data_ptr(), st), "hgk_aug_rotate")
        wb, off = rb, pad
    return A.resize(wb, Hn, Wn, off, off, Hn - 2 * off, Wn - 2 * off, res, res)     # :175 imresize(new_img, (res, res))


_DESC_FIELDS = ("src", "H", "W", "pre_h", "pre_w", "o_bytes", "o_ptmp", "o_pout", "c_pw_b", "c_pw_k", "ks_pw", "c_ph_b", "c_ph_k",
                "ks_ph", "Hn", "Wn", "ny0", "ny1", "nx0", "nx1", "oy", "ox", "has_zero", "o_win", "rot", "o_rot", "pad", "in_h", "in_w",
                "o_ftmp", "c_fw_b", "c_fw_k", "ks_fw", "c_fh_b", "c_fh_k", "ks_fh", "out_index", "chw", "flip")      # struct AugDesc, csrc/warp.cu
_batch_scratch = {}


def _crop_batch_u8(imgs, centers, scales, rots, res, size, flips=None, gains=None):
    """All crops of a batch through hgk_aug_crop_batch: the host lays out one descriptor row per image (geometry as in
    `crop`, offsets of every intermediate into two scratch arenas), uploads the table once, and a dozen launches with
    blockIdx.y = image do the pixel work.  Returns the uint8 stack [N,res,res,3].
    flips / gains given: the images are the resident 3 x H x W tensors and every source pixel is read W-flipped (flips[k]),
    times gains[k][channel], clamped to [0, 1] -- the flip / colour augmentation of the agent's loader, never materialised."""
    chw = gains is not None
    lib = get_lib()
    n = len(imgs)
    dev = imgs[0].device
    assert lib.cdll.hgk_aug_desc_fields() == len(_DESC_FIELDS)
    rows, mats, keep = [], [], []
    ou8 = oi = 0

    def take_u8(nbytes):
        nonlocal ou8
        o = ou8
        ou8 += (int(nbytes) + 15) // 16 * 16
        return o

    def take_i(nints):
        nonlocal oi
        o = oi
        oi += int(nints)
        return o

    for k in range(n):
        img = imgs[k]
        if not isinstance(img, torch.Tensor) or not img.is_cuda:
            raise HGKError("HumanAug.crop_batch runs on CUDA tensors only (no CPU fallback)")
        if img.dim() != 3 or img.shape[0 if chw else 2] != 3 or img.dtype != torch.float32:
            raise ValueError("crop_batch: expected float32 %s images, got %s %s" % ("3 x H x W" if chw else "H x W x 3", tuple(img.shape), img.dtype))
        img = img.contiguous()
        keep.append(img)
        H, W = (int(img.shape[1]), int(img.shape[2])) if chw else (int(img.shape[0]), int(img.shape[1]))
        if min(H, W) <= 4:
            raise ValueError("crop_batch: image too small")
        rot = float(rots[k])
        win = _window(H, W, centers[k], scales[k], rot, res, size)
        if win is None:
            raise ValueError("crop_batch: sample %d hit the degenerate early return of crop (image smaller than 2 px after shrinking)" % k)
        sf, pre, (Hn, Wn), (new_y, new_x), (old_y, old_x), pad = win
        if Hn <= 4 or Wn <= 4 or old_y[1] <= old_y[0] or old_x[1] <= old_x[0] or \
                (new_y[1] - new_y[0], new_x[1] - new_x[0]) != (old_y[1] - old_y[0], old_x[1] - old_x[0]):
            raise ValueError("crop_batch: the crop window of sample %d does not intersect the image" % k)
        d = dict.fromkeys(_DESC_FIELDS, 0)
        d.update(src=img.data_ptr(), H=H, W=W, Hn=Hn, Wn=Wn, ny0=new_y[0], ny1=new_y[1], nx0=new_x[0], nx1=new_x[1],
                 oy=old_y[0] - new_y[0], ox=old_x[0] - new_x[0], out_index=k, chw=int(chw), flip=int(bool(flips[k])) if chw else 0,
                 has_zero=int((new_y[1] - new_y[0]) < Hn or (new_x[1] - new_x[0]) < Wn))
        if pre is not None:
            ks_w, ks_h = lib.aug_resample_ksize(W, pre[1]), lib.aug_resample_ksize(H, pre[0])
            d.update(pre_h=pre[0], pre_w=pre[1], o_bytes=take_u8(H * W * 3), o_ptmp=take_u8(H * pre[1] * 3),
                     o_pout=take_u8(pre[0] * pre[1] * 3), ks_pw=ks_w, ks_ph=ks_h,
                     c_pw_b=take_i(pre[1] * 2), c_pw_k=take_i(pre[1] * ks_w), c_ph_b=take_i(pre[0] * 2), c_ph_k=take_i(pre[0] * ks_h))
        d["o_win"] = take_u8(Hn * Wn * 3)
        off = 0
        m = [0.0] * 6
        if not rot == 0:
            if pad <= 0 or Hn - 2 * pad <= 0 or Wn - 2 * pad <= 0:
                raise ValueError("crop: empty image after removing the rotation padding")
            d.update(rot=1, o_rot=take_u8(Hn * Wn * 3), pad=pad)
            m = _rotate_matrix(rot, Wn, Hn)
            off = pad
        in_h, in_w = Hn - 2 * off, Wn - 2 * off
        ks_w, ks_h = lib.aug_resample_ksize(in_w, res), lib.aug_resample_ksize(in_h, res)
        d.update(in_h=in_h, in_w=in_w, o_ftmp=take_u8(in_h * res * 3), ks_fw=ks_w, ks_fh=ks_h,
                 c_fw_b=take_i(res * 2), c_fw_k=take_i(res * ks_w), c_fh_b=take_i(res * 2), c_fh_k=take_i(res * ks_h))
        rows.append([d[f] for f in _DESC_FIELDS])
        # (the gains travel as doubles and are narrowed to float32 on the device: torch multiplies a float32 tensor by the
        #  float32-rounded Python scalar)
        mats.append(list(m) + ([float(v) for v in gains[k]] if chw else [1.0, 1.0, 1.0]))
    desc_h = np.ascontiguousarray(np.array(rows, dtype=np.int64))
    desc_d = torch.from_numpy(desc_h).to(dev)
    mats_d = torch.from_numpy(np.array(mats, dtype=np.float64)).to(dev)
    arena = torch.empty(max(ou8, 16), device=dev, dtype=torch.uint8)
    arena_i = torch.empty(max(oi, 4), device=dev, dtype=torch.int32)
    minmax = torch.empty(n * 4, device=dev, dtype=torch.float64)
    key = (dev.type, dev.index)
    if key not in _batch_scratch or _batch_scratch[key].numel() < n * 6:
        _batch_scratch[key] = torch.tensor([-1, 0, 0] * (2 * max(n, 64)), device=dev, dtype=torch.int32)   # idle state, re-armed by the kernels
    stack = torch.empty(n, res, res, 3, device=dev, dtype=torch.uint8)
    lib.check(lib.aug_crop_batch(desc_h.ctypes.data, desc_d.data_ptr(), mats_d.data_ptr(), n, res, arena.data_ptr(),
                                 arena_i.data_ptr(), minmax.data_ptr(), _batch_scratch[key].data_ptr(), stack.data_ptr(),
                                 torch.cuda.current_stream(dev).cuda_stream), "hgk_aug_crop_batch")
    return stack


def crop_batch(imgs, centers, scales, rots, res=256, size=200, batched=True):
    """`inp = im_to_torch(crop(im_to_numpy(img), c, s, r, res, size)).float()` (ref data/joint_train_s_r_agent.py:200-203)
    for every sample of a batch: imgs = list of float32 CUDA H x W x 3 tensors (any sizes), centers [N,2], scales [N],
    rots [N] host arrays.  Returns [N,3,res,res] float32 on the GPU -- the `img` batch of load_batch_data.
    batched=True: one descriptor table and a dozen launches for the whole batch (hgk_aug_crop_batch); False: `crop` per image
    (the same bytes, ~10 launches per image)."""
    n = len(imgs)
    if n == 0:
        raise ValueError("crop_batch: empty batch")
    centers = np.asarray(centers, dtype=np.float32).reshape(n, 2)
    scales = np.asarray(scales, dtype=np.float32).reshape(n)
    rots = np.asarray(rots, dtype=np.float64).reshape(n)
    dev = imgs[0].device
    if batched:
        stack = _crop_batch_u8(imgs, centers, scales, rots, res, size)
    else:
        stack = torch.empty(n, res, res, 3, device=dev, dtype=torch.uint8)
        for k in range(n):
            c = crop(imgs[k], centers[k], scales[k], rots[k], res, size)
            if c.dtype != torch.uint8:
                raise ValueError("crop_batch: sample %d hit the degenerate early return of crop (image smaller than 2 px after shrinking)" % k)
            stack[k].copy_(c)
    out = torch.empty(n, 3, res, res, device=dev, dtype=torch.float32)
    lib = get_lib()
    lib.check(lib.aug_to_chw_float(stack.data_ptr(), n, res, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
              "hgk_aug_to_chw_float")
    return out


def crop_batch_resident(imgs_chw, flips, gains, centers, scales, rots, res=256, size=200):
    """crop_batch for the agent's loader: imgs_chw = the resident float32 3 x H x W CUDA images (what load_image returns);
    flips[k] / gains[k] = the horizontal flip and the three colour gains `AGENT.__getitem__` applies before the crop
    (ref data/joint_train_s_r_agent.py:160-168), evaluated on load instead of materialising a flipped, scaled, transposed copy
    of every image.  Returns [N,3,res,res] float32."""
    n = len(imgs_chw)
    if n == 0:
        raise ValueError("crop_batch_resident: empty batch")
    centers = np.asarray(centers, dtype=np.float32).reshape(n, 2)
    scales = np.asarray(scales, dtype=np.float32).reshape(n)
    rots = np.asarray(rots, dtype=np.float64).reshape(n)
    stack = _crop_batch_u8(imgs_chw, centers, scales, rots, res, size, flips=list(flips), gains=[list(g) for g in gains])
    dev = imgs_chw[0].device
    out = torch.empty(n, 3, res, res, device=dev, dtype=torch.float32)
    lib = get_lib()
    lib.check(lib.aug_to_chw_float(stack.data_ptr(), n, res, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
              "hgk_aug_to_chw_float")
    return out
