import sys, ctypes; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from hgk_testlib import *
L = lib()
L.cdll.hgk_debug_set_timeline.argtypes=[ctypes.c_void_p]
buf = torch.zeros(512,16, dtype=torch.long, device=DEV)
def run(N,H,W,Ci,Co,k,split):
    x = torch.randn(N,H,W,Ci, device=DEV); y = torch.empty(N,H,W,Co, device=DEV)
    w = torch.randn(Co,Ci,k,k, dtype=torch.float64)*0.1
    src = dev32(w.reshape(-1)); dst = torch.zeros(2*w.numel(), device=DEV)
    table = torch.tensor([[0,0,w.numel(),Co,Ci,k*k,0,Co]], dtype=torch.long, device=DEV)
    call("pack_weights_tc", ptr(src), ptr(dst), ptr(table), 1)
    hi, lo = dst[:w.numel()], dst[w.numel():]
    for rep in range(3):
        buf.zero_()
        flush = torch.empty(64 << 20, device=DEV).fill_(1.0); del flush      # 256 MB write: evicts the weights from L2
        L.cdll.hgk_debug_set_timeline(buf.data_ptr() if rep==2 else 0)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        call("conv_tc_nhwc", ptr(x),0,0,0,N,H,W,Ci,ptr(hi),ptr(lo) if split else 0,k,0,Co,0,0,0,0,ptr(y),0,0,0)
        e1.record(); torch.cuda.synchronize()
    L.cdll.hgk_debug_set_timeline(0)
    b = buf.cpu()
    t0 = b[:,0][b[:,0]>0].min()
    print('conv %dx%d %d->%d k%d split=%d: %.1f us total' % (H,W,Ci,Co,k,split, e0.elapsed_time(e1)*1e3))
    names=['start','alloc+sync','loads issued','stage0 stored','producer loop end','mma done','epi1 done','epi2 done','issuer fullA0','issuer fullB0']
    for cta in ((0, 1, 2) if N*H*W <= 4096 else (0, 1, 147, 148, 300)):
        r = b[cta]
        print('  cta %3d start@%7.1fus ' % (cta, (r[0]-t0)/1e3) + ' '.join('%s=%.1f' % (names[i].split()[0], (r[i]-r[0])/1e3) for i in (1,2,3,8,9,4,5,6,7)))
import os
if os.environ.get("TL_SMALL"):
    # cold weights: flush L2 between repetitions by touching a large buffer
    run(24,8,8,128,128,3,1)
    run(24,4,4,128,128,3,1)
    run(24,8,8,256,128,1,1)
    run(24,4,4,128,256,1,1)
else:
    run(24,64,64,128,256,1,0)
    run(24,64,64,256,128,1,1)
    run(24,64,64,128,128,3,1)
