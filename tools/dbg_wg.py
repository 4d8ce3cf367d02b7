import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from hgk_testlib import *
N,H,W,Ci,Co,k = 2,16,16,64,128,1
x = torch.ones(N,H,W,Ci, device=DEV); dz = torch.ones(N,H,W,Co, device=DEV)
gw = torch.zeros(k*k,Co,Ci, device=DEV); gb = torch.zeros(Co, device=DEV)
call("conv_wgrad_tc_nhwc", ptr(x),0,0,0,N,H,W,Ci,ptr(dz),Co,k,ptr(gw),ptr(gb))
torch.cuda.synchronize()
print('gw sum', float(gw.sum()), 'expected', N*H*W*Co*Ci, 'min/max', float(gw.min()), float(gw.max()))
print('gb', gb[:8].tolist(), 'expected', N*H*W)
# random
x = torch.randn(N,H,W,Ci, device=DEV); dz = torch.randn(N,H,W,Co, device=DEV)
gw.zero_()
call("conv_wgrad_tc_nhwc", ptr(x),0,0,0,N,H,W,Ci,ptr(dz),Co,k,ptr(gw),0)
torch.cuda.synchronize()
ref = torch.einsum('nhwo,nhwi->oi', dz.double(), x.double()).float()
print('rand relerr', float((gw[0]-ref).abs().max()/ref.abs().max()), float(gw.abs().max()), float(ref.abs().max()))
print(gw[0,:4,:4]); print(ref[:4,:4])
