"""Deterministic synthetic weights / MPII-shaped batches (no dataset, no checkpoint: there is no network on the
bench box).  Used by bench.py for both arms, by __graft_entry__.smoke() and -- through the re-export
`oracle/synth.py` -- by the golden generators and the tests.  Pure numpy/torch host code: nothing here is on the
CUDA path and nothing here imports the oracle.

Everything is a pure function of (name, shape, seed) through numpy's PCG64 so the same
tensors can be rebuilt on any machine without shipping multi-megabyte weight files:
the reference model, the oracle and the CUDA path are all loaded from the same dict.

Distributions follow the reference initialiser (models/asn_stacked_hg.py:258-270,381-393):
conv / linear weight and bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)), BN gamma ~ U(0,1).
To make parity tests sensitive to every term we additionally perturb BN beta (reference: 0)
and the running statistics (reference: 0 / 1); a loaded checkpoint has such values anyway.
"""
import math
import zlib
from collections import OrderedDict

import numpy as np
import torch


def _rng(name, seed):
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def schema_of(module):
    """[(name, shape)] of a module's state_dict -- the same list oracle.hg_oracle.*_schema spells out by hand
    (tests/test_abi_cpu.py checks both against the reference's own state_dict)."""
    return [(k, tuple(v.shape)) for k, v in module.state_dict().items()]


def make_state_dict(schema, seed=0, dtype=torch.float32, perturb_bn=True):
    """schema: list of (name, shape): `schema_of(module)` or oracle.hg_oracle.*_schema."""
    sd = OrderedDict()
    fan = {}
    for name, shape in schema:
        if name.endswith(".weight") and len(shape) >= 2:
            fan[name[:-len(".weight")]] = int(np.prod(shape[1:]))
    for name, shape in schema:
        r = _rng(name, seed)
        base = name.rsplit(".", 1)[0]
        leaf = name.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            sd[name] = torch.zeros((), dtype=torch.long)
            continue
        if base in fan:                                   # conv / linear weight + bias
            stdv = 1.0 / math.sqrt(fan[base])
            a = r.uniform(-stdv, stdv, size=shape)
        elif leaf == "weight":                            # BN gamma
            a = r.uniform(0.0, 1.0, size=shape)
            if perturb_bn:
                a = 0.25 + 0.75 * a                       # keep away from 0 so grads stay informative
        elif leaf == "bias":                              # BN beta
            a = r.uniform(-0.2, 0.2, size=shape) if perturb_bn else np.zeros(shape)
        elif leaf == "running_mean":
            a = r.uniform(-0.1, 0.1, size=shape) if perturb_bn else np.zeros(shape)
        elif leaf == "running_var":
            a = r.uniform(0.5, 1.5, size=shape) if perturb_bn else np.ones(shape)
        else:
            raise KeyError(name)
        sd[name] = torch.from_numpy(np.asarray(a, dtype=np.float64)).to(dtype)
    return sd


def make_images(n, res, seed=0, dtype=torch.float32):
    """Images U[0,1) fp32 NCHW [n,3,res,res] (the reference feeds un-normalised [0,1] crops,
    data/mpii_for_mpii.py:141 has color_normalize commented out)."""
    r = _rng("images", seed)
    return torch.from_numpy(r.random((n, 3, res, res))).to(dtype)


def make_heatmaps(n, res, num_classes=16, seed=0, dtype=torch.float32, p_absent=0.15):
    """MPII-shaped targets [n,num_classes,res/4,res/4]: zeros plus one 7x7 blob
    exp(-(dx^2+dy^2)/9) per present joint (what pylib/HumanPts.py:84,94-99 draw_gaussian
    renders for sigma=1: tmp_size=3, denominator tmp_size^2), ~15 % joints absent (:41-43)."""
    r = _rng("heatmaps", seed)
    h = res // 4
    t = np.zeros((n, num_classes, h, h), dtype=np.float64)
    ax = np.arange(-3, 4)
    blob = np.exp(-(ax[None, :] ** 2 + ax[:, None] ** 2) / 9.0)
    for i in range(n):
        for j in range(num_classes):
            if r.random() < p_absent:
                continue
            cy, cx = int(r.integers(0, h)), int(r.integers(0, h))
            y0, y1 = max(cy - 3, 0), min(cy + 4, h)
            x0, x1 = max(cx - 3, 0), min(cx + 4, h)
            t[i, j, y0:y1, x0:x1] = blob[y0 - cy + 3:y1 - cy + 3, x0 - cx + 3:x1 - cx + 3]
    return torch.from_numpy(t).to(dtype)


def make_tensor(name, shape, seed=0, lo=-1.0, hi=1.0, dtype=torch.float32):
    """Generic named uniform tensor for per-kernel tests."""
    r = _rng(name, seed)
    return torch.from_numpy(r.uniform(lo, hi, size=shape)).to(dtype)


def make_photo(h, w, seed=0):
    """A synthetic 'photograph' for the crop / rotate / resize augmentation: H x W x 3 float32 in [0,1] with smooth
    large-scale structure (a bilinearly up-sampled random 1/16-resolution field) plus fine noise, numpy only (the same
    bytes on every machine)."""
    r = np.random.Generator(np.random.PCG64(zlib.crc32(b"photo") ^ (seed * 7919 + h * 31 + w)))
    gh, gw = h // 16 + 2, w // 16 + 2
    g = r.random((gh, gw, 3))
    ys = np.linspace(0.0, gh - 1.001, h)
    xs = np.linspace(0.0, gw - 1.001, w)
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    rows = g[y0] * (1.0 - fy) + g[y0 + 1] * fy
    img = rows[:, x0] * (1.0 - fx) + rows[:, x0 + 1] * fx
    img = img * r.uniform(0.55, 1.0) + r.uniform(0.0, 0.1) + r.normal(0.0, 0.02, size=img.shape)
    return np.clip(img, 0.0, 1.0).astype(np.float32)
