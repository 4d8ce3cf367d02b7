"""Stem forward micro-benchmark at the headline shape (24 x 3 x 256 x 256): the FFMA kernel against the space-to-depth
tensor-core form, with and without the BatchNorm-statistics epilogue.  python tools/bench_stem.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pose_adv_aug_b200 import get_lib          # noqa: E402

lib = get_lib()
dev = torch.device("cuda", 0)
N, H, W = 24, 256, 256
H2, W2 = H // 2, W // 2
st = torch.cuda.current_stream().cuda_stream
img = [torch.rand(N, 3, H, W, device=dev) for _ in range(4)]
w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
b = torch.randn(64, device=dev) * 0.1
y = [torch.empty(N, H2, W2, 64, device=dev) for _ in range(4)]
xs = [torch.empty(N, H2, W2, 16, device=dev) for _ in range(4)]
ws = torch.zeros(64, 32, 4, 4, device=dev)
ssum = torch.zeros(64, device=dev, dtype=torch.float64)
ssq = torch.zeros(64, device=dev, dtype=torch.float64)
vec = [torch.ones(64, device=dev) for _ in range(8)]
ticket = torch.zeros(1, device=dev, dtype=torch.int32)
dst = torch.zeros(2 * ws.numel(), device=dev)
lib.check(lib.stem_s2d_weight(w.data_ptr(), 64, ws.data_ptr(), st))
table = torch.tensor([[0, 0, ws.numel(), 64, 32, 16, 0, 64]], dtype=torch.long, device=dev)
lib.check(lib.pack_weights_tc(ws.data_ptr(), dst.data_ptr(), table.data_ptr(), 1, st))
hi, lo = dst[:ws.numel()], dst[ws.numel():]
for k in range(4):
    lib.check(lib.stem_s2d_image(img[k].data_ptr(), N, H, W, xs[k].data_ptr(), st))


def timeit(name, fn, reps=40):
    for k in range(4):
        fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        fn(r % 4)
    e1.record()
    torch.cuda.synchronize()
    print("%-60s %8.1f us" % (name, e0.elapsed_time(e1) / reps * 1e3))


timeit("stem_conv7_fwd (FFMA, stats)", lambda k: lib.check(lib.stem_conv7_fwd(
    img[k].data_ptr(), N, H, W, w.data_ptr(), b.data_ptr(), 64, y[k].data_ptr(), ssum.data_ptr(), ssq.data_ptr(), st)))
timeit("stem_s2d_image", lambda k: lib.check(lib.stem_s2d_image(img[k].data_ptr(), N, H, W, xs[k].data_ptr(), st)))
timeit("stem_s2d_weight", lambda k: lib.check(lib.stem_s2d_weight(w.data_ptr(), 64, ws.data_ptr(), st)))
base = lambda k: [xs[k].data_ptr(), 0, 0, 0, N, H2, W2, 16, hi.data_ptr(), lo.data_ptr(), 4, b.data_ptr(), 64, 0, 0, 0, 0, y[k].data_ptr(), 0]
timeit("conv k4 3xTF32, no statistics", lambda k: lib.check(lib.conv_tc_nhwc(*(base(k) + [0, 0, st]))))
timeit("conv k4 3xTF32, statistics", lambda k: lib.check(lib.conv_tc_nhwc(*(base(k) + [ssum.data_ptr(), ssq.data_ptr(), st]))))
timeit("conv k4 3xTF32, statistics + BatchNorm finaliser", lambda k: lib.check(lib.conv_tc_bn_nhwc(*(base(k) + [
    ssum.data_ptr(), ssq.data_ptr(), vec[0].data_ptr(), vec[1].data_ptr(), 1e-5, 0.1, vec[2].data_ptr(), vec[3].data_ptr(),
    vec[4].data_ptr(), vec[5].data_ptr(), vec[6].data_ptr(), vec[7].data_ptr(), ticket.data_ptr(), st]))))
