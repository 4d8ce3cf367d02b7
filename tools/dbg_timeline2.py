"""Per-CTA phase timeline of the image-tile tcgen05 kernel (conv_tc2_kernel): globaltimer stamps written by
thread 0 / the MMA issuer of the first 512 CTAs (developer diagnostics, HGK_STAMP in csrc/tc_common.cuh).
usage (GPU box): python tools/dbg_timeline2.py"""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from hgk_testlib import *
L = lib()
L.cdll.hgk_debug_set_timeline.argtypes = [ctypes.c_void_p]
buf = torch.zeros(512, 16, dtype=torch.long, device=DEV)
NAMES = {1: 'alloc', 2: 'loads', 3: 'store0', 8: 'fullA0', 9: 'fullB0', 4: 'prodend', 5: 'mmadone', 6: 'epiend'}


def run(N, H, W, Ci, Co, k, split, stats):
    x = torch.randn(N, H, W, Ci, device=DEV); y = torch.empty(N, H, W, Co, device=DEV)
    sc = torch.rand(Ci, device=DEV) + 0.5; sh = torch.randn(Ci, device=DEV) * 0.1
    w = torch.randn(Co, Ci, k, k, dtype=torch.float64) * 0.1
    src = dev32(w.reshape(-1)); dst = torch.zeros(2 * w.numel(), device=DEV)
    table = torch.tensor([[0, 0, w.numel(), Co, Ci, k * k, 0, Co]], dtype=torch.long, device=DEV)
    call("pack_weights_tc", ptr(src), ptr(dst), ptr(table), 1)
    hi, lo = dst[:w.numel()], dst[w.numel():]
    s1 = torch.zeros(Co, dtype=torch.float64, device=DEV); s2 = torch.zeros(Co, dtype=torch.float64, device=DEV)
    bias = torch.randn(Co, device=DEV)
    for rep in range(3):
        buf.zero_()
        flush = torch.empty(64 << 20, device=DEV).fill_(1.0); del flush
        L.cdll.hgk_debug_set_timeline(buf.data_ptr() if rep == 2 else 0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        call("conv_tc_nhwc", ptr(x), ptr(sc), ptr(sh), 1, N, H, W, Ci, ptr(hi), ptr(lo) if split else 0, k, ptr(bias), Co,
             0, 0, 0, 0, ptr(y), 0, ptr(s1) if stats else 0, ptr(s2) if stats else 0)
        e1.record(); torch.cuda.synchronize()
    L.cdll.hgk_debug_set_timeline(0)
    b = buf.cpu().double()[:256]
    live = b[:, 0] > 0
    t0 = b[live, 0].min()
    print('conv %dx%d %d->%d k%d split=%d stats=%d: %.1f us total, %d CTAs stamped' %
          (H, W, Ci, Co, k, split, stats, e0.elapsed_time(e1) * 1e3, int(live.sum())))
    r = b[live]
    order = (1, 2, 3, 8, 9, 4, 5, 6)
    print('  mean offsets from CTA start (us): ' + ' '.join('%s=%.2f' % (NAMES[i], float(((r[:, i] - r[:, 0]) / 1e3).mean())) for i in order))
    print('  phases (us): prologue=%.2f mainloop(first store->mma done)=%.2f epilogue=%.2f lifetime=%.2f' % (
        float(((r[:, 3] - r[:, 0]) / 1e3).mean()), float(((r[:, 5] - r[:, 3]) / 1e3).mean()),
        float(((r[:, 6] - r[:, 5]) / 1e3).mean()), float(((r[:, 6] - r[:, 0]) / 1e3).mean())))
    starts = ((r[:, 0] - t0) / 1e3)
    ends = ((r[:, 6] - t0) / 1e3)
    print('  CTA start times (us) pct 0/25/50/75/100: ' + ' '.join('%.1f' % float(torch.quantile(starts, q)) for q in (0, .25, .5, .75, 1)))
    print('  last stamped CTA end: %.1f us' % float(ends.max()))
    tr = buf.cpu()[256:263].double()
    base = tr[tr > 0].min()
    KC = Ci // 16
    kinds = ['P pass ea', 'P arrived', 'M got fa', 'M got fb', 'M issued', 'W eb free', 'P stored']
    print('  CTA 0 per-chunk trace (SM cycles from the first event; chunk = 16 input channels):')
    for k in (0, 6, 1, 5, 2, 3, 4):
        print('   %-10s ' % kinds[k] + ' '.join('%6d' % int(tr[k, c] - base) if tr[k, c] > 0 else '     -' for c in range(min(KC, 16))))
    te = buf.cpu()[264:268].double()
    ek = ['E start', 'E tmem->stg', 'E rows done', 'E stats done']
    print('  CTA 0 epilogue trace (SM cycles from the first event; column = Cout chunk of 128 x pixel half):')
    for k in range(4):
        print('   %-12s ' % ek[k] + ' '.join('%6d' % int(te[k, c] - base) if te[k, c] > 0 else '     -' for c in range(4)))
    for cta in (0, 1, 200):
        if cta < r.shape[0]:
            q = r[cta]
            print('  cta %3d start@%7.1f ' % (cta, float((q[0] - t0) / 1e3)) + ' '.join('%s=%.1f' % (NAMES[i], float((q[i] - q[0]) / 1e3)) for i in order))


if os.environ.get("TL_SMALL"):
    # the 8x8 / 4x4 rungs (conv_tc_kernel, 128 linear pixels per CTA): trace columns are pipeline iterations (32 channels x tap)
    def run_small(N, H, W, Ci, Co, k):
        global buf
        x = torch.randn(N, H, W, Ci, device=DEV); y = torch.empty(N, H, W, Co, device=DEV)
        sc = torch.rand(Ci, device=DEV) + 0.5; sh = torch.randn(Ci, device=DEV) * 0.1
        w = torch.randn(Co, Ci, k, k, dtype=torch.float64) * 0.1
        src = dev32(w.reshape(-1)); dst = torch.zeros(2 * w.numel(), device=DEV)
        table = torch.tensor([[0, 0, w.numel(), Co, Ci, k * k, 0, Co]], dtype=torch.long, device=DEV)
        call("pack_weights_tc", ptr(src), ptr(dst), ptr(table), 1)
        hi, lo = dst[:w.numel()], dst[w.numel():]
        s1 = torch.zeros(Co, dtype=torch.float64, device=DEV); s2 = torch.zeros(Co, dtype=torch.float64, device=DEV)
        for rep in range(3):
            buf.zero_()
            L.cdll.hgk_debug_set_timeline(buf.data_ptr() if rep == 2 else 0)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            call("conv_tc_nhwc", ptr(x), ptr(sc), ptr(sh), 1, N, H, W, Ci, ptr(hi), ptr(lo), k, 0, Co, 0, 0, 0, 0, ptr(y), 0, ptr(s1), ptr(s2))
            e1.record(); torch.cuda.synchronize()
        L.cdll.hgk_debug_set_timeline(0)
        tr = buf.cpu()[256:263].double()
        base = tr[tr > 0].min()
        kinds = ['P pass em', 'P arrived', 'M got fa', 'M got fb', 'M issued', '-', 'P stored']
        print('conv_tc_kernel %dx%d %d->%d k%d (warm L2): %.1f us; CTA 0 per-iteration trace (SM cycles):' % (H, W, Ci, Co, k, e0.elapsed_time(e1) * 1e3))
        for kk in (0, 6, 1, 2, 3, 4):
            print('   %-10s ' % kinds[kk] + ' '.join('%6d' % int(tr[kk, c] - base) if tr[kk, c] > 0 else '     -' for c in range(16)))
    run_small(24, 4, 4, 128, 128, 3)
    run_small(24, 8, 8, 128, 128, 3)
    run_small(24, 4, 4, 256, 128, 1)
    sys.exit(0)
run(24, 64, 64, 128, 256, 1, 1, 1)
run(24, 64, 64, 256, 128, 1, 1, 1)
run(24, 64, 64, 128, 128, 3, 1, 1)
run(24, 64, 64, 256, 128, 1, 0, 0)
run(24, 32, 32, 128, 256, 1, 1, 1)
run(24, 16, 16, 128, 128, 3, 1, 1)
