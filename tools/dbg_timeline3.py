"""Who waits for whom inside the persistent tile kernel (conv_tc3_kernel): per-role barrier-wait cycles of every CTA
(T3_WAIT / T3_DBG in csrc/conv_tc3.cu), plus the kernel time cold-cache.
usage (GPU box): python tools/dbg_timeline3.py [fwd|dgrad|all]"""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from hgk_testlib import *
L = lib()
L.cdll.hgk_debug_set_timeline.argtypes = [ctypes.c_void_p]
buf = torch.zeros(256, 32, dtype=torch.long, device=DEV)
SLOTS = {0: "T total", 1: "T wait raw", 12: "T transform", 13: "T fence+arrive", 2: "M total", 3: "M wait afull", 4: "M wait bfull",
         5: "M wait accempty", 6: "LA total", 7: "LA wait aempty", 8: "LB total", 9: "LB wait bempty", 10: "E0 total",
         11: "E0 wait accfull", 16: "E0 tmem->patch", 17: "CTA total"}


def report(tag, ms):
    b = buf.cpu().double()
    live = b[:, 2] > 0
    r = b[live]
    span = float(r[:, 15].max() - r[:, 14].min()) / 1e3
    mhz = float((r[:, 17] / (r[:, 15] - r[:, 14]).clamp_min(1.0)).median()) * 1e3
    print('%s: events %.1f us, kernel span (globaltimer) %.1f us, start spread %.1f us, %d CTAs, SM clock ~%.0f MHz' % (
        tag, ms * 1e3, span, float(r[:, 14].max() - r[:, 14].min()) / 1e3, int(live.sum()), mhz))
    for i, nme in SLOTS.items():
        col = r[:, i] / mhz      # us
        print('   %-18s mean %8.2f us   min %8.2f   max %8.2f' % (nme, float(col.mean()), float(col.min()), float(col.max())))


def pack(w, mode, BN):
    src = dev32(w.reshape(-1)); dst = torch.zeros(2 * w.numel(), device=DEV)
    O, I, k, _ = w.shape
    Nn, K = (O, I) if mode == 0 else (I, O)
    table = torch.tensor([[0, 0, w.numel(), Nn, K, k * k, mode, BN]], dtype=torch.long, device=DEV)
    call("pack_weights_tc", ptr(src), ptr(dst), ptr(table), 1)
    return dst[:w.numel()], dst[w.numel():]


WARM = False


def timed(fn):
    ms = 0.0
    for rep in range(3):
        buf.zero_()
        if not WARM:
            flush = torch.empty(64 << 20, device=DEV).fill_(1.0); del flush
        L.cdll.hgk_debug_set_timeline(buf.data_ptr() if rep == 2 else 0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    L.cdll.hgk_debug_set_timeline(0)
    return ms


def run_fwd(N, H, W, Ci, Co, k, split=1, stats=1, res=0):
    x = torch.randn(N, H, W, Ci, device=DEV); y = torch.empty(N, H, W, Co, device=DEV)
    sc = torch.rand(Ci, device=DEV) + 0.5; sh = torch.randn(Ci, device=DEV) * 0.1
    w = torch.randn(Co, Ci, k, k, dtype=torch.float64) * 0.1
    hi, lo = pack(w, 0, Co)
    s1 = torch.zeros(Co, dtype=torch.float64, device=DEV); s2 = torch.zeros(Co, dtype=torch.float64, device=DEV)
    bias = torch.randn(Co, device=DEV)
    r = torch.randn(N, H, W, Co, device=DEV) if res else None
    ms = timed(lambda: call("conv_tc_nhwc", ptr(x), ptr(sc), ptr(sh), 1, N, H, W, Ci, ptr(hi), ptr(lo) if split else 0, k, ptr(bias), Co,
                            ptr(r), 0, 0, 0, ptr(y), 0, ptr(s1) if stats else 0, ptr(s2) if stats else 0))
    report('fwd %dx%d %d->%d k%d split=%d stats=%d res=%d' % (H, W, Ci, Co, k, split, stats, res), ms)


def run_dgrad(N, H, W, Cin, Cout, k, red=1, acc=0):
    """bnapply data gradient: input g [.., Cin] (Cin = forward Cout), output [.., Cout]"""
    g = torch.randn(N, H, W, Cin, device=DEV); gz = torch.randn(N, H, W, Cin, device=DEV)
    v = lambda C: torch.rand(C, device=DEV) + 0.5
    gs, gt, gm, cA, cB, cC = v(Cin), v(Cin), v(Cin), v(Cin), v(Cin) * 0.01, v(Cin) * 0.01
    dz = torch.empty_like(g)
    w = torch.randn(Cin, Cout, k, k, dtype=torch.float64) * 0.1      # forward weight [O=Cin][I=Cout]
    hi, _ = pack(w, 1, Cout)
    dy = torch.zeros(N, H, W, Cout, device=DEV)
    bz = torch.randn(N, H, W, Cout, device=DEV)
    bs, bt, bm, bi = v(Cout), v(Cout), v(Cout), v(Cout)
    sg = torch.zeros(Cout, dtype=torch.float64, device=DEV); sgx = torch.zeros(Cout, dtype=torch.float64, device=DEV)
    gamma = v(Cout); dgam = torch.zeros(Cout, device=DEV); dbet = torch.zeros(Cout, device=DEV)
    oA, oB, oC = v(Cout), v(Cout), v(Cout)
    tick = torch.zeros(1, dtype=torch.int32, device=DEV)
    if red:
        tail = [ptr(bz), ptr(bs), ptr(bt), 1, ptr(bm), ptr(bi), ptr(sg), ptr(sgx), ptr(gamma), 1, ptr(dgam), ptr(dbet), ptr(oA), ptr(oB),
                ptr(oC), ptr(tick)]
    else:
        tail = [0] * 16
    ms = timed(lambda: call("conv_tc_dgrad_bnapply_nhwc", ptr(g), ptr(gz), ptr(gs), ptr(gt), 1, ptr(gm), ptr(cA), ptr(cB), ptr(cC), ptr(dz),
                            N, H, W, Cin, ptr(hi), k, Cout, 0, ptr(dy), acc, *tail))
    report('dgrad+bnapply %dx%d %d->%d k%d red=%d acc=%d' % (H, W, Cin, Cout, k, red, acc), ms)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "warm":         # everything L2-resident (no flush, small batch): separates HBM latency from the rest
    WARM = True
    run_fwd(6, 64, 64, 128, 256, 1, res=1)
    run_fwd(6, 64, 64, 256, 128, 1)
    WARM = False
    run_fwd(6, 64, 64, 128, 256, 1, res=1)
    run_fwd(6, 64, 64, 256, 128, 1)
if which == "one":          # a single layer (ncu capture)
    run_fwd(24, 64, 64, 128, 256, 1, res=1)
if which == "one3":
    run_fwd(24, 64, 64, 128, 128, 3)
if which == "oned":
    run_dgrad(24, 64, 64, 256, 128, 1, red=1)
if which in ("fwd", "all"):
    run_fwd(24, 64, 64, 128, 256, 1, res=1)
    run_fwd(24, 64, 64, 256, 128, 1)
    run_fwd(24, 64, 64, 128, 128, 3)
    run_fwd(24, 32, 32, 128, 128, 3)
    run_fwd(24, 16, 16, 128, 128, 3)
    run_fwd(24, 8, 8, 256, 128, 1)
if which in ("dgrad", "all"):
    run_dgrad(24, 64, 64, 256, 128, 1, red=1)
    run_dgrad(24, 64, 64, 128, 256, 1, red=0, acc=1)
    run_dgrad(24, 64, 64, 128, 128, 3, red=1)
