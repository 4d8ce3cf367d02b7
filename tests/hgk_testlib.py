"""Helpers for the -m gpu parity tests: call libhgk entry points on torch CUDA tensors and
build float64 CPU references with plain torch ops (the oracle's building blocks)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import synth          # noqa: E402

DEV = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")


def lib():
    from pose_adv_aug_b200._lib import get_lib
    return get_lib()


def stream():
    return torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else 0


def ptr(t):
    return 0 if t is None else t.data_ptr()


def call(name, *args):
    L = lib()
    rc = getattr(L, name)(*args, stream())
    assert rc == 0, "%s -> %d: %s" % (name, rc, L.last_error())


def rnd(name, shape, lo=-1.0, hi=1.0, seed=0):
    return synth.make_tensor(name, shape, seed=seed, lo=lo, hi=hi, dtype=torch.float64)


def dev32(t):
    return t.to(torch.float32).to(DEV).contiguous()


def nhwc(t):
    """NCHW (cpu, any dtype) -> NHWC fp32 device tensor."""
    return dev32(t.permute(0, 2, 3, 1))


def from_nhwc(t):
    """NHWC device tensor -> NCHW float64 cpu."""
    return t.detach().cpu().double().permute(0, 3, 1, 2).contiguous()


def relerr(got, ref):
    got = got.detach().cpu().double()
    ref = ref.detach().cpu().double()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def affine_act(x, scale, shift, relu):
    """reference of a virtual activation on an NCHW fp64 tensor"""
    if scale is None:
        return x
    y = x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    return F.relu(y) if relu else y


def pack_w(w, mode):
    """OIHW fp64 cpu -> packed fp32 device: mode 0 [tap][I][O], mode 1 [tap][O][I]."""
    O, I, kh, kw = w.shape
    t = w.reshape(O, I, kh * kw)
    if mode == 0:
        return dev32(t.permute(2, 1, 0))
    return dev32(t.permute(2, 0, 1))
