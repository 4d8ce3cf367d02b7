"""Drop-in for the reference `pylib/Criterion.py` (same names, argument meaning, return value):
each loss is one fused reduction kernel forward and one elementwise kernel backward.

  weighted_sigmoid_crossentropy(pred, gt, weight)   ref pylib/Criterion.py:4-10
  weighted_L2(pred, gt, weight)                      ref pylib/Criterion.py:12-18
"""
import torch

from .._lib import get_lib, HGKError


class _Criterion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, pred, gt, weight):
        for t in (pred, gt, weight):
            if not t.is_cuda:
                raise HGKError("Criterion kernels run on CUDA only (no CPU fallback)")
        if pred.shape != gt.shape:
            raise ValueError("pred and gt must have the same shape")
        lib = get_lib()
        p = pred.contiguous().float()
        g = gt.contiguous().float()
        w = weight.expand_as(pred).contiguous().float()
        acc = torch.zeros(1, device=pred.device, dtype=torch.float64)
        out = torch.empty((), device=pred.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(pred.device).cuda_stream
        lib.check(lib.criterion_fwd(kind, p.data_ptr(), g.data_ptr(), w.data_ptr(), p.numel(), acc.data_ptr(), stream),
                  "hgk_criterion_fwd")
        lib.check(lib.f64_to_f32(acc.data_ptr(), out.data_ptr(), 1, 1.0, stream), "hgk_f64_to_f32")
        ctx.kind = kind
        ctx.save_for_backward(p, g, w)
        return out

    @staticmethod
    def backward(ctx, gout):
        p, g, w = ctx.saved_tensors
        lib = get_lib()
        go = gout.contiguous().float().reshape(1)
        dp = torch.empty_like(p)
        stream = torch.cuda.current_stream(p.device).cuda_stream
        lib.check(lib.criterion_bwd(ctx.kind, p.data_ptr(), g.data_ptr(), w.data_ptr(), p.numel(), go.data_ptr(),
                                    dp.data_ptr(), stream), "hgk_criterion_bwd")
        return None, dp, None, None


def weighted_sigmoid_crossentropy(pred, gt, weight):
    # pred: torch.sigmoid(output); gt: 0/1 label map; weight: w=1 for gt==0, w>1 for gt==1
    return _Criterion.apply(1, pred, gt, weight)


def weighted_L2(pred, gt, weight):
    # pred: net output; gt: [0,1] heatmap; weight: w=1 for gt==0, w>1 for gt>0
    return _Criterion.apply(0, pred, gt, weight)
