"""CPU oracle of the image crop / rotate / resize augmentation (warp half).

TEST INFRASTRUCTURE ONLY: imported by tests/, oracle/gen_golden_aug.py and nothing in the product package.

Restates
  * `crop(img, center, scale, rot, res, size)`                     ref pylib/HumanAug.py:117-175
    (called from `gen_img_heatmap`, ref data/joint_train_s_r_agent.py:200-204, with the agent's sampled scale / rotation,
     ref joint-train-pose-s-r-agent.py:425-450)
  * the four `scipy.misc` functions it calls.  `scipy.misc.{bytescale,toimage,fromimage,imresize,imrotate}` are a THIRD-PARTY
    DEPENDENCY THAT IS ABSENT HERE (removed in SciPy 1.3; this image has SciPy 1.18): the reference pins no version, its
    era (PyTorch 0.3, Python 2.7) is SciPy 0.19 - 1.0 and NumPy 1.13.  They are restated below from the published algorithm of
    `scipy/misc/pilutil.py` (SciPy 1.0.0) -- thin wrappers over PIL: `toimage` byte-scales a float array to 0..255 over its OWN
    min / max (a contrast stretch -- the reference inherits that quirk), `imresize` = `Image.resize(BILINEAR)`, `imrotate` =
    `Image.rotate(BILINEAR)`.  The PIL calls themselves are made for real (Pillow 12.2.0 in this image).

Pinned how: `oracle/make_ref.py` streams the reference's own `pylib/HumanAug.py` into `oracle/_ref/ref_humanaug.py`; run with
`install_scipy_misc_shim()` (this file's restated `scipy.misc`) the REFERENCE `crop` gives byte-identical images to `crop`
below on every golden case (tests/test_oracle_aug.py, live when /root/reference is present; tests/golden/aug_crop.npz otherwise).
What stays unpinned is the restated `scipy.misc` layer itself (no SciPy <= 1.2 exists here to run) and the Pillow version
(2017's Pillow 4.x and today's 12.2 share the two-pass fixed-point resize; `Image.rotate` rounds its matrix to 15 digits since
Pillow 5).

NumPy-1.x arithmetic is spelled out where NumPy 2 would differ (`bytescale` on a float32 array multiplies in float32).
"""
import math
import sys
import types

import numpy as np
from PIL import Image

_INTERP = {'nearest': 0, 'lanczos': 1, 'bilinear': 2, 'bicubic': 3, 'cubic': 3}      # pilutil.imresize / imrotate


# ---------------------------------------------------------------------------------------------------------------------
# scipy.misc (pilutil.py, SciPy 1.0.0) restated
# ---------------------------------------------------------------------------------------------------------------------
def bytescale(data, cmin=None, cmax=None, high=255, low=0):
    """pilutil.bytescale: uint8 passes through; anything else is stretched from [min, max] to [low, high], + 0.5, truncated.
    Arithmetic dtype as NumPy 1.x value-based casting gives it: a float32 array stays float32 (the float64 scalar `scale`
    is cast down), a float64 array is float64."""
    data = np.asarray(data)
    if data.dtype == np.uint8:
        return data
    if cmin is None:
        cmin = data.min()
    if cmax is None:
        cmax = data.max()
    cscale = cmax - cmin
    if cscale < 0:
        raise ValueError("`cmax` should be larger than `cmin`.")
    elif cscale == 0:
        cscale = 1
    scale = np.float64(high - low) / np.float64(cscale)
    dt = data.dtype if data.dtype in (np.float32, np.float64) else np.float64
    bytedata = (data.astype(dt) - dt.type(cmin)) * dt.type(scale) + dt.type(low)
    return (bytedata.clip(low, high) + dt.type(0.5)).astype(np.uint8)


def toimage(arr):
    """pilutil.toimage for the two shapes `crop` produces: H x W (mode 'L') and H x W x 3 (mode 'RGB')."""
    data = np.asarray(arr)
    if np.iscomplexobj(data):
        raise ValueError("Cannot convert a complex-valued array.")
    shape = list(data.shape)
    if len(shape) == 2:
        bytedata = bytescale(data)
        return Image.frombytes('L', (shape[1], shape[0]), bytedata.tobytes())
    if len(shape) != 3 or shape[2] != 3 or 3 in shape[:2] or 4 in shape[:2]:
        # (pilutil picks the FIRST axis of length 3 as the channel axis: a 3-pixel-high image would be transposed)
        raise ValueError("oracle restates toimage for H x W and H x W x 3 arrays with H, W > 4 only")
    bytedata = bytescale(data)
    return Image.frombytes('RGB', (shape[1], shape[0]), bytedata.tobytes())


def fromimage(im):
    return np.array(im)


def imresize(arr, size, interp='bilinear', mode=None):
    """pilutil.imresize: int = percent, float = fraction, tuple = (rows, cols)."""
    im = toimage(arr)
    ts = type(size)
    if np.issubdtype(ts, np.signedinteger):
        percent = size / 100.0
        size = tuple((np.array(im.size) * percent).astype(int))
    elif np.issubdtype(ts, np.floating):
        size = tuple((np.array(im.size) * size).astype(int))
    else:
        size = (size[1], size[0])
    imnew = im.resize(tuple(int(s) for s in size), resample=_INTERP[interp])
    return fromimage(imnew)


def imrotate(arr, angle, interp='bilinear'):
    arr = np.asarray(arr)
    im = toimage(arr)
    im = im.rotate(angle, resample=_INTERP[interp])
    return fromimage(im)


def install_scipy_misc_shim():
    """Make `import scipy.misc` inside the streamed reference file resolve the four functions above."""
    import scipy
    shim = types.ModuleType("scipy.misc")
    shim.bytescale, shim.toimage, shim.fromimage, shim.imresize, shim.imrotate = bytescale, toimage, fromimage, imresize, imrotate
    sys.modules["scipy.misc"] = shim
    scipy.misc = shim
    return shim


# ---------------------------------------------------------------------------------------------------------------------
# HumanAug.crop restated
# ---------------------------------------------------------------------------------------------------------------------
def _crop_to_image_matrix(center, scale, res, size):
    """Inverse of GetTransform(center, scale, 0, res, size) (ref :10-21,38-44): crop pixel -> source pixel, inverted
    numerically as the reference does (LAPACK), because the corners are TRUNCATED to integers afterwards.  `center` and
    `scale` reach the reference as float32 arrays (torch .numpy()), so the matrix entries are float32 arithmetic."""
    f32 = np.float32
    h = f32(size) * f32(scale)
    t = np.zeros((3, 3))
    t[0, 0] = f32(res) / h
    t[1, 1] = f32(res) / h
    t[0, 2] = f32(res) * (f32(-float(center[0])) / h + f32(.5))
    t[1, 2] = f32(res) * (f32(-float(center[1])) / h + f32(.5))
    t[2, 2] = 1
    return np.linalg.inv(t)


def crop_geometry(img_shape, center, scale, rot, res, size):
    """Everything `crop` decides before it touches pixels (ref :121-158).  Returns a dict, or None for the degenerate
    early return of ref :128-129.  center: 2 floats, scale: float -- both taken as float32 (the dataset's tensors)."""
    f32 = np.float32
    center = np.asarray(center, dtype=np.float32).reshape(2)
    scale = f32(np.asarray(scale, dtype=np.float32).reshape(-1)[0])
    scale_factor = float(scale * f32(size)) / float(res)                  # ref :121
    pre = None
    ht, wd = int(img_shape[0]), int(img_shape[1])
    if scale_factor < 2:                                                  # ref :123-124
        scale_factor = 1
    else:
        new_img_size = np.floor(max(ht, wd) / scale_factor)               # ref :126
        if new_img_size < 2:
            return None
        frac = 1 / scale_factor                                           # imresize(size=float): (W*f, H*f) truncated
        wh = (np.array([wd, ht]) * frac).astype(int)
        pre = (int(wh[1]), int(wh[0]))
        ht, wd = pre
    center = center / f32(scale_factor)                                   # ref :133-134 (float32 arrays / python scalar)
    scale = scale / f32(scale_factor)
    inv = _crop_to_image_matrix(center, scale, res, size)
    ul = np.dot(inv, np.array([0, 0, 1.]))[:2].astype(int)                # ref :137
    br = np.dot(inv, np.array([res, res, 1.]))[:2].astype(int)            # ref :139
    if scale_factor >= 2:                                                 # ref :142-143
        br = br - (br - ul - res)
    pad = int(np.ceil(np.linalg.norm(br - ul) / 2 - float(br[1] - ul[1]) / 2).astype(int))     # ref :146
    if not rot == 0:
        ul = ul - pad
        br = br + pad
    Hn, Wn = int(br[1] - ul[1]), int(br[0] - ul[0])
    new_x = (max(0, -int(ul[0])), min(int(br[0]), wd) - int(ul[0]))       # ref :158-162
    new_y = (max(0, -int(ul[1])), min(int(br[1]), ht) - int(ul[1]))
    old_x = (max(0, int(ul[0])), min(wd, int(br[0])))
    old_y = (max(0, int(ul[1])), min(ht, int(br[1])))
    return dict(pre=pre, Hn=Hn, Wn=Wn, new_x=new_x, new_y=new_y, old_x=old_x, old_y=old_y, pad=pad,
                rot=float(rot), scale_factor=scale_factor)


def crop(img, center, scale, rot, res, size):
    """img: H x W x 3 float array in [0, 1] (ref imutils.im_to_numpy of the loaded image); returns the res x res x 3 uint8
    crop around `center` of a `scale * size` pixel box, rotated by `rot` degrees (ref pylib/HumanAug.py:117-175)."""
    img = np.asarray(img)
    g = crop_geometry(img.shape, center, scale, rot, res, size)
    if g is None:
        return img
    if g["pre"] is not None:
        img = imresize(img, size=1 / g["scale_factor"], interp='bilinear')               # ref :131
        assert img.shape[:2] == g["pre"]
    new_shape = [g["Hn"], g["Wn"]] + ([img.shape[2]] if img.ndim > 2 else [])
    new_img = np.zeros(new_shape)                                                        # float64 (ref :156)
    (ny0, ny1), (nx0, nx1) = g["new_y"], g["new_x"]
    (oy0, oy1), (ox0, ox1) = g["old_y"], g["old_x"]
    new_img[ny0:ny1, nx0:nx1] = img[oy0:oy1, ox0:ox1]                                    # ref :164
    if not rot == 0:                                                                     # ref :166-170
        new_img = imrotate(new_img, rot, interp='bilinear')
        pad = g["pad"]
        new_img = new_img[pad:-pad, pad:-pad]
    return imresize(new_img, (res, res))                                                 # ref :175


def im_to_torch_float(img_u8):
    """utils/imutils.im_to_torch (ref data/joint_train_s_r_agent.py:203): HWC -> CHW float, divided by 255 only when the
    maximum exceeds 1."""
    a = np.transpose(np.asarray(img_u8), (2, 0, 1)).astype(np.float32)
    if a.max() > 1:
        a = a / np.float32(255)
    return a


# ---------------------------------------------------------------------------------------------------------------------
# the index-list path of AGENT.__getitem__ + gen_img_heatmap (ref data/joint_train_s_r_agent.py:100-175,196-217), i.e. what
# load_batch_data (ref joint-train-pose-s-r-agent.py:425-450) returns for one batch with num_workers = 0
# ---------------------------------------------------------------------------------------------------------------------
def get_transform(center, scale, rot, res, size):
    """HumanAug.GetTransform, ref pylib/HumanAug.py:10-36, for float32 `center` / `scale` arrays."""
    f32 = np.float32
    center = np.asarray(center, dtype=np.float32)
    h = f32(size) * f32(np.asarray(scale, dtype=np.float32).reshape(-1)[0])
    t = np.zeros((3, 3))
    t[0, 0] = f32(res) / h
    t[1, 1] = f32(res) / h
    t[0, 2] = f32(res) * (f32(-float(center[0])) / h + f32(.5))
    t[1, 2] = f32(res) * (f32(-float(center[1])) / h + f32(.5))
    t[2, 2] = 1
    if not rot == 0:
        rot = -rot
        rot_mat = np.zeros((3, 3))
        rot_rad = rot * np.pi / 180
        sn, cs = np.sin(rot_rad), np.cos(rot_rad)
        rot_mat[0, :2] = [cs, -sn]
        rot_mat[1, :2] = [sn, cs]
        rot_mat[2, 2] = 1
        t_mat = np.eye(3)
        t_mat[0, 2] = -res / 2
        t_mat[1, 2] = -res / 2
        t_inv = t_mat.copy()
        t_inv[:2, 2] *= -1
        t = np.dot(t_inv, np.dot(rot_mat, np.dot(t_mat, t)))
    return t


def transform_pts(pts, center, scale, rot, res, size):
    """HumanAug.TransformPts (forward), ref pylib/HumanAug.py:46-55."""
    t = get_transform(center, scale, rot, res, size)
    new_pt = np.concatenate((pts, np.ones((pts.shape[0], 1))), axis=1).T
    return np.dot(t, new_pt)[0:2, :].T


_MATCHED = ([0, 5], [1, 4], [2, 3], [10, 15], [11, 14], [12, 13])


def agent_batch(images_chw, annos, scale_index_list, rotation_index_list, img_index, rng, inp_res=256, out_res=64, std_size=200):
    """images_chw: list of float32 [3,H,W] arrays in [0,1].  Returns (img [N,3,res,res] f32, heatmap [N,16,out,out] f32,
    c [N,2] f32, s [N] f32, r [N], pts [N,16,2] f32, normalizer [N]) -- the random numbers are drawn from `rng`
    (a numpy RandomState) in the order __getitem__ draws them from the global state."""
    from oracle import eval_oracle as E
    f32 = np.float32
    scale_means, rotation_means = np.arange(-0.6, 0.61, 0.2), np.arange(-60, 61, 20)

    def small_gaussian(mean, var):                                          # ref :15-16
        return max(mean - var + 1e-3, min(mean + var, mean + rng.randn() * var))

    out = [[] for _ in range(7)]
    for k, i in enumerate(img_index):
        a = annos[i]
        img = np.array(images_chw[i], dtype=np.float32)
        pts = np.asarray(a['joint_self'], dtype=np.float32)[:, 0:2].copy()
        c = np.asarray(a['objpos'], dtype=np.float32).copy()
        s = f32(a['scale_provided'])
        c[1] = c[1] + f32(15) * s                                           # ref :122-124
        s = s * f32(1.25)
        normalizer = a['normalizer'] * 0.6
        scale_factor = small_gaussian(scale_means[scale_index_list[k]], 0.05)                  # ref :139-143
        r_aug = small_gaussian(rotation_means[rotation_index_list[k]], 5)
        s_aug = s * f32(2 ** scale_factor)
        if rng.random_sample() <= 0.5:                                      # ref :160-163
            img = img[:, :, ::-1].copy()
            width = img.shape[2]
            pts[:, 0] = f32(width) - pts[:, 0]
            for p0, p1 in _MATCHED:
                tmp = pts[p0].copy()
                pts[p0] = pts[p1]
                pts[p1] = tmp
            c[0] = f32(width) - c[0]
        for ch in range(3):                                                 # ref :166-168
            img[ch] = np.clip(img[ch] * f32(rng.uniform(0.6, 1.4)), 0, 1)
        inp = crop(np.transpose(img, (1, 2, 0)), c, np.array([s_aug], dtype=np.float32), r_aug, inp_res, std_size)
        inp = im_to_torch_float(inp)
        pts_aug = transform_pts(pts, c, np.array([s_aug], dtype=np.float32), r_aug, out_res, std_size)
        pts_aug[(pts[:, 0] <= 0) | (pts[:, 1] <= 0), :] = 0                 # ref :207-210
        heatmap, _ = E.pts2heatmap(pts_aug, [out_res, out_res], sigma=1)
        for lst, v in zip(out, (inp, heatmap.astype(np.float32), c, s_aug, r_aug, pts, normalizer)):
            lst.append(v)
    return tuple(np.stack(v) for v in out)


# ---------------------------------------------------------------------------------------------------------------------
# own restatement of the two PIL algorithms the CUDA kernels implement (checked against PIL itself in tests/test_oracle_aug.py)
# ---------------------------------------------------------------------------------------------------------------------
def pil_resample_coeffs(in_size, out_size):
    """Pillow `precompute_coeffs` + `normalize_coeffs_8bpc` for the BILINEAR filter (support 1) over the full box:
    per output index the first input index, the tap count and the 22-bit fixed-point taps."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = []
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            v = 1.0 - a if a < 1.0 else 0.0
            w.append(v)
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << 22)) if k < 0 else int(0.5 + k * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_resize_bilinear_u8(a, out_h, out_w):
    """Pillow ImagingResample on an 8-bit image: horizontal pass, then vertical pass, each rounding to uint8."""
    a = np.asarray(a, dtype=np.uint8)
    H, W = a.shape[:2]
    bh, kh = pil_resample_coeffs(W, out_w)
    tmp = np.zeros((H, out_w) + a.shape[2:], dtype=np.uint8)
    for xx in range(out_w):
        x0, n = bh[xx]
        acc = (a[:, x0:x0 + n].astype(np.int64) * kh[xx, :n].reshape((1, n) + (1,) * (a.ndim - 2))).sum(axis=1) + (1 << 21)
        tmp[:, xx] = np.clip(acc >> 22, 0, 255)
    bv, kv = pil_resample_coeffs(H, out_h)
    out = np.zeros((out_h, out_w) + a.shape[2:], dtype=np.uint8)
    for yy in range(out_h):
        y0, n = bv[yy]
        acc = (tmp[y0:y0 + n].astype(np.int64) * kv[yy, :n].reshape((n, 1) + (1,) * (a.ndim - 2))).sum(axis=0) + (1 << 21)
        out[yy] = np.clip(acc >> 22, 0, 255)
    return out


def pil_rotate_matrix(angle, w, h):
    """`Image.rotate` (expand = 0, centre = image centre): output pixel -> input pixel affine coefficients."""
    angle = angle % 360.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    cx, cy = w / 2.0, h / 2.0
    m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2]
    m[5] = m[3] * (-cx) + m[4] * (-cy) + 0.0
    m[2] += cx
    m[5] += cy
    return m


def pil_rotate_bilinear_u8(a, angle):
    """Pillow ImagingGenericTransform(affine_transform, bilinear_filter) on an 8-bit image, fill 0."""
    a = np.asarray(a, dtype=np.uint8)
    H, W = a.shape[:2]
    m = pil_rotate_matrix(angle, W, H)
    ys, xs = np.meshgrid(np.arange(H) + 0.5, np.arange(W) + 0.5, indexing="ij")
    xin = m[0] * xs + m[1] * ys + m[2]
    yin = m[3] * xs + m[4] * ys + m[5]
    ok = (xin >= 0.0) & (xin < W) & (yin >= 0.0) & (yin < H)
    xin = xin - 0.5
    yin = yin - 0.5
    x = np.where(xin < 0.0, np.floor(xin), np.trunc(xin)).astype(np.int64)
    y = np.where(yin < 0.0, np.floor(yin), np.trunc(yin)).astype(np.int64)
    dx = xin - x
    dy = yin - y
    x0 = np.clip(x, 0, W - 1)
    x1 = np.clip(x + 1, 0, W - 1)
    y0 = np.clip(y, 0, H - 1)
    af = a.astype(np.float64)
    ex = (slice(None), slice(None)) + (None,) * (a.ndim - 2)
    p00, p01 = af[y0, x0], af[y0, x1]
    v1 = p00 + (p01 - p00) * dx[ex]
    has = (y + 1 >= 0) & (y + 1 < H)
    y1 = np.clip(y + 1, 0, H - 1)
    p10, p11 = af[y1, x0], af[y1, x1]
    v2 = np.where(has[ex], p10 + (p11 - p10) * dx[ex], v1)
    v = v1 + (v2 - v1) * dy[ex]
    out = v.astype(np.uint8)
    out[~ok] = 0
    return out
