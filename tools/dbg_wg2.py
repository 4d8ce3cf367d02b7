"""Developer check of the MN-major weight-gradient kernel on one small problem: prints error structure."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from hgk_testlib import *
N, H, W, Ci, Co, k = [int(v) for v in sys.argv[1:7]] if len(sys.argv) > 6 else (1, 8, 8, 128, 128, 1)
torch.manual_seed(0)
x = torch.randn(N, H, W, Ci, device=DEV)
dz = torch.randn(N, H, W, Co, device=DEV)
gw = torch.zeros(k * k, Co, Ci, device=DEV)
gb = torch.zeros(Co, device=DEV)
call("conv_wgrad_tc_nhwc", ptr(x), 0, 0, 0, N, H, W, Ci, ptr(dz), Co, k, ptr(gw), ptr(gb))
torch.cuda.synchronize()
xr = x.permute(0, 3, 1, 2).double().cpu().requires_grad_(False)
w = torch.zeros(Co, Ci, k, k, dtype=torch.float64, requires_grad=True)
y = torch.nn.functional.conv2d(xr, w, None, padding=k // 2)
y.backward(dz.permute(0, 3, 1, 2).double().cpu())
ref = w.grad.reshape(Co, Ci, k * k).permute(2, 0, 1)
got = gw.cpu().double()
print("env", {k_: v for k_, v in os.environ.items() if k_.startswith("HGK_")})
print("bias err", float((gb.cpu().double() - dz.double().cpu().sum(dim=(0, 1, 2))).abs().max()), "bias max", float(gb.abs().max()))
print("got absmax %.4g nonzero frac %.4f ref absmax %.4g" % (float(got.abs().max()), float((got != 0).double().mean()), float(ref.abs().max())))
print("relerr", float((got - ref).abs().max() / ref.abs().max()))
for t in range(k * k):
    e = (got[t] - ref[t]).abs()
    print(" tap", t, "err max %.3g" % float(e.max()), "rows ok", int((e.max(dim=1).values < 1e-2 * ref.abs().max()).sum()), "/", Co,
          "cols ok", int((e.max(dim=0).values < 1e-2 * ref.abs().max()).sum()), "/", Ci)
# does got match ref under simple permutations?
g0, r0 = got[0], ref[0]
print("corr with ref:", float((g0 * r0).sum() / (g0.norm() * r0.norm() + 1e-30)), " with ref^T (if square):",
      float((g0 * r0.t()).sum() / (g0.norm() * r0.norm() + 1e-30)) if Co == Ci else None)
print("got[0,:3,:6]\n", g0[:3, :6], "\nref[0,:3,:6]\n", r0[:3, :6])
