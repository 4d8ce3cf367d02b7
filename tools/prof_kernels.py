"""Launches the big-layer kernels of the config-2 step in isolation (for `ncu --set full`): every kernel runs
`reps` times on rotating buffers (> L2) so that each profiled launch sees cold inputs.
usage: python tools/prof_kernels.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from hgk_testlib import *

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N, H, W = 24, 64, 64
NB = 4                                              # rotating buffer sets: 4 x (50 + 100 MB) > 126 MB L2


def pack(Co, Ci, k, mode):
    w = torch.randn(Co, Ci, k, k, dtype=torch.float64) * 0.05
    src = dev32(w.reshape(-1)); dst = torch.zeros(2 * w.numel(), device=DEV)
    Nn, K = (Co, Ci) if mode != 1 else (Ci, Co)
    table = torch.tensor([[0, 0, w.numel(), Nn, K, k * k, mode, Nn]], dtype=torch.long, device=DEV)
    call("pack_weights_tc", ptr(src), ptr(dst), ptr(table), 1)
    return dst[:w.numel()], dst[w.numel():]


bufs = {}
def buf(C, i):
    key = (C, i % NB)
    if key not in bufs:
        bufs[key] = torch.randn(N, H, W, C, device=DEV)
    return bufs[key]


def vec(C, lo=0.5, hi=1.5):
    return torch.rand(C, device=DEV) * (hi - lo) + lo


def fwd(Ci, Co, k, i):
    # the shipping forward products: TF32 + 2xBF16 (weight pack mode 2, hgk_conv_tc_x2_nhwc); PROF_X2=0: 3xTF32
    x2 = os.environ.get("PROF_X2", "1") == "1"
    hi, lo = pack(Co, Ci, k, 2 if x2 else 0)
    x, y = buf(Ci, i), torch.empty(N, H, W, Co, device=DEV)
    sc, sh, b = vec(Ci), vec(Ci, -0.3, 0.3), vec(Co, -0.1, 0.1)
    s1 = torch.zeros(Co, device=DEV, dtype=torch.float64); s2 = torch.zeros_like(s1)
    call("conv_tc_x2_nhwc" if x2 else "conv_tc_nhwc", ptr(x), ptr(sc), ptr(sh), 1, N, H, W, Ci, ptr(hi), ptr(lo), k, ptr(b), Co, 0, 0, 0, 0,
         ptr(y), 0, ptr(s1), ptr(s2))


def dgrad(Ci, Co, k, i):            # conv Ci -> Co; gradient Co -> Ci with the fused BN reduction
    hi, _ = pack(Co, Ci, k, 1)
    dz, gx, bz = buf(Co, i), torch.empty(N, H, W, Ci, device=DEV), buf(Ci, i + 1)
    v = [vec(Ci) for _ in range(4)]
    s1 = torch.zeros(Ci, device=DEV, dtype=torch.float64); s2 = torch.zeros_like(s1)
    call("conv_tc_dgrad_bnstats_nhwc", ptr(dz), N, H, W, Co, ptr(hi), 0, k, Ci, 0, ptr(gx), 0, ptr(bz), ptr(v[0]), ptr(v[1]), 1,
         ptr(v[2]), ptr(v[3]), ptr(s1), ptr(s2))


def wgrad(Ci, Co, k, i):
    x, dz = buf(Ci, i), buf(Co, i + 2)
    sc, sh = vec(Ci), vec(Ci, -0.3, 0.3)
    gw = torch.zeros(k * k, Co, Ci, device=DEV); gb = torch.zeros(Co, device=DEV)
    call("conv_wgrad_tc_nhwc", ptr(x), ptr(sc), ptr(sh), 1, N, H, W, Ci, ptr(dz), Co, k, ptr(gw), ptr(gb))


for i in range(reps):
    fwd(128, 128, 3, i); fwd(128, 256, 1, i); fwd(256, 128, 1, i)
    dgrad(128, 128, 3, i); dgrad(256, 128, 1, i); dgrad(128, 256, 1, i)
    wgrad(128, 128, 3, i); wgrad(128, 256, 1, i); wgrad(256, 128, 1, i)
torch.cuda.synchronize()
print("done")
