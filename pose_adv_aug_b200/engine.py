"""Host-side plan builder / executor for the B200 hourglass path.

The reference executes its graph op by op through torch-0.3 autograd (one THCUNN / cuDNN
launch per nn.Module call).  Here a network forward is *planned once* per input shape into a
static list of C-ABI kernel launches (include/hgk.h) over pre-allocated NHWC buffers; the
backward list is derived from the same tape.  Activations that follow a BatchNorm are kept
"virtual": the stored tensor is the pre-BN convolution output z and the consumers apply
relu(z*scale+shift) on load (boundary shift).

PyTorch is used for device memory, streams and autograd glue only; every arithmetic op of the
path is a libhgk kernel.
"""
import ctypes
import os

import torch

from ._lib import get_lib, HGKError

BN_EPS = 1e-5        # nn.BatchNorm2d defaults used by the reference (models/asn_stacked_hg.py:19)
BN_MOMENTUM = 0.1
_ALIGN = 32          # elements; keeps every parameter 128-byte aligned inside the flat buffers


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------
# Flat parameter / gradient storage
# ------------------------------------------------------------------------------------------
class ParamStore(object):
    """All parameters of a root module as views into ONE flat fp32 buffer (+ one flat gradient
    buffer, + flat BN running-stat buffers).  This is what the flat RMSprop kernel and the single
    NCCL all-reduce operate on (the reference's nn.DataParallel reduce_add + ~2000 optimizer
    launches, stack-hg.py:49,165)."""

    def __init__(self, module, device):
        self.device = device
        self.params = []
        seen = set()
        for p in module.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = max(off, _ALIGN)
        self.flat = torch.zeros(self.numel, device=device, dtype=torch.float32)
        self.grad = torch.zeros(self.numel, device=device, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data.to(device=device, dtype=torch.float32))
                old_grad = p.grad
                p.data = view
                gview = self.grad[o:o + p.numel()].view(p.shape)
                if old_grad is not None:
                    gview.copy_(old_grad.to(device=device, dtype=torch.float32))
                p.grad = gview
        # buffers: fp32 running statistics and int64 num_batches_tracked
        self.fbufs, self.ibufs = [], []
        seen = set()
        for name, b in module.named_buffers():
            if id(b) in seen:
                continue
            seen.add(id(b))
            (self.fbufs if b.dtype.is_floating_point else self.ibufs).append(b)
        foff, self.fboff = 0, []
        for b in self.fbufs:
            self.fboff.append(foff)
            foff += (b.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.fbuf_flat = torch.zeros(max(foff, _ALIGN), device=device, dtype=torch.float32)
        self.ibuf_flat = torch.zeros(max(len(self.ibufs), 1), device=device, dtype=torch.long)
        with torch.no_grad():
            for b, o in zip(self.fbufs, self.fboff):
                view = self.fbuf_flat[o:o + b.numel()].view(b.shape)
                view.copy_(b.data.to(device=device, dtype=torch.float32))
                b.data = view
            for i, b in enumerate(self.ibufs):
                view = self.ibuf_flat[i:i + 1].view(b.shape)
                view.copy_(b.data.to(device))
                b.data = view
        self.index = dict((id(p), i) for i, p in enumerate(self.params))
        self._ptrs = [p.data_ptr() for p in self.params]
        self._bptrs = [b.data_ptr() for b in self.fbufs + self.ibufs]
        self.version = 0

    def grad_ptr(self, p_index):
        return self.grad.data_ptr() + 4 * self.offsets[p_index]

    def valid(self):
        """True while every parameter / buffer is still the view this store created."""
        for p, ptr in zip(self.params, self._ptrs):
            if p.data_ptr() != ptr or p.dtype != torch.float32:
                return False
        for b, ptr in zip(self.fbufs + self.ibufs, self._bptrs):
            if b.data_ptr() != ptr:
                return False
        return True

    def attach_grads(self):
        """Re-attach .grad views dropped by zero_grad(set_to_none=True) (zeroing their storage)."""
        missing = [i for i, p in enumerate(self.params)
                   if p.grad is None or p.grad.data_ptr() != self.grad_ptr(i)]
        if not missing:
            return
        if len(missing) == len(self.params):
            self.grad.zero_()
        for i in missing:
            p, o = self.params[i], self.offsets[i]
            gview = self.grad[o:o + p.numel()].view(p.shape)
            if len(missing) != len(self.params):
                if p.grad is not None:
                    gview.copy_(p.grad)
                else:
                    gview.zero_()
            p.grad = gview

    def index_of(self, p):
        return self.index[id(p)]


# ------------------------------------------------------------------------------------------
# Plan
# ------------------------------------------------------------------------------------------
class T(object):
    """A tensor node of the plan: NHWC buffer z [N,H,W,C] + optional on-load affine/ReLU."""
    __slots__ = ("z", "N", "H", "W", "C", "scale", "shift", "relu", "needs_grad", "grad", "contribs",
                 "producer", "name", "bn", "n_cons", "n_done", "bwd_stats_fused", "bwd_fin_fused")

    def __init__(self, z, N, H, W, C, scale=None, shift=None, relu=False, needs_grad=False, name=""):
        self.z, self.N, self.H, self.W, self.C = z, N, H, W, C
        self.scale, self.shift, self.relu = scale, shift, relu
        self.needs_grad = needs_grad
        self.grad = None          # buffer holding d loss / d (activated value) once initialised
        self.contribs = []        # pending (buffer, donatable) gradient contributions
        self.producer = None
        self.name = name
        self.bn = None            # BNRec whose (scale, shift) this tensor carries
        self.n_cons = 0           # number of ops that consume this tensor
        self.n_done = 0           # ... of which have emitted their gradient contribution (backward plan construction)
        self.bwd_stats_fused = False   # the BN-backward reduction was produced by the consumer's data-gradient kernel
        self.bwd_fin_fused = False     # ... and so were dgamma / dbeta / cA / cB / cC (last-CTA finaliser)

    def use(self):
        self.n_cons += 1
        return self

    @property
    def P(self):
        return self.N * self.H * self.W

    def act_args(self):
        return [_ptr(self.z), _ptr(self.scale), _ptr(self.shift), int(self.relu)]


class BNRec(object):
    __slots__ = ("gamma", "beta", "rmean", "rvar", "C", "scale", "shift", "mean", "invstd", "sum", "sq",
                 "sum_g", "sum_gx", "cA", "cB", "cC", "training")


class Plan(object):
    """Static launch lists (forward, backward) for one module / input shape / mode."""

    VEC_CAP = 1 << 22      # fp32 per-channel vectors (scale, shift, mean, invstd, cA, cB, cC)
    STAT_CAP = 1 << 20     # fp64 per-channel accumulators

    def __init__(self, stores, device, training, need_grad, conv_path=0, precise_grads=False):
        self.lib = get_lib()
        self.stores = list(stores) if isinstance(stores, (list, tuple)) else [stores]
        self.device = device
        self.training = training       # current BN mode; builders may switch it between sub-modules
        self.need_grad = need_grad
        self.conv_path = conv_path
        self.fwd = []              # [fn, args(list), name]
        self.bwd = []
        self.tape = []
        self.keep = []             # tensors kept alive
        self.vec = torch.zeros(self.VEC_CAP, device=device, dtype=torch.float32)
        self.vec_used = 0
        self.stat_f = torch.zeros(self.STAT_CAP, device=device, dtype=torch.float64)
        self.stat_f_used = 0
        self.stat_b = torch.zeros(self.STAT_CAP if need_grad else 1, device=device, dtype=torch.float64)
        self.stat_b_used = 0
        self.use_tc = conv_path != 1   # tcgen05 convolutions where the shape is covered
        # gradients: plain TF32 operands by default (error below the fp32-vs-fp64 noise floor of the
        # whole net); precise_grads -> 3xTF32 data gradients + fp32 SIMT weight gradients
        self.precise_grads = precise_grads
        self.fuse_bn_bwd = os.environ.get("HGK_FUSE_BN_BWD", "1") == "1"   # fold BN-backward reductions into single-consumer data-gradient epilogues
        # BatchNorm finalisers (scale/shift/running stats; dgamma/dbeta/cA/cB/cC) run by the last CTA of the kernel that
        # produced the sums instead of a C-length launch of their own
        self.fuse_bn_fin = os.environ.get("HGK_FUSE_BN_FIN", "1") == "1"
        # BatchNorm-backward apply evaluated on load by the image-tile data-gradient kernel (no bn_bwd_apply launch)
        self.fuse_bn_apply = os.environ.get("HGK_FUSE_BN_APPLY", "1") == "1"
        self.tickets = torch.zeros(8192, device=device, dtype=torch.int32)   # last-CTA counters (re-armed by the kernels)
        self.tickets_used = 0
        self.tc_entries = []       # (src param, mode, BN, hi offset, lo offset or -1)
        self.tc_used = 0
        self.tc_buf = None
        self.tc_table = None
        self.tc_launch = None
        self.wg_entries = []       # (param, scratch offset): 3x3 weight grads accumulated tap-major
        self.wg_used = 0
        self.wg_buf = None
        self.wg_table = None
        self.pack_entries = []     # (src param, dst offset, O, I, taps, mode)
        self.pack_used = 0
        self.pack_buf = None
        self.pack_table = None
        self.pack_launch = None
        self.dyn = {}              # name -> list of (launch, arg index) patched at run time
        self.inputs = []
        self.outputs = []          # (kind, tensor node or torch tensor)
        self.bn_recs = []
        self.nbt_bufs = []
        self.nbt_flat = []
        self.fwd_count = 0
        self.bytes_alloc = 0
        self.finished = False
        self.low_recs = set()      # id() of forward launches built inside a low_scope() (off the critical path)
        self.x2_ok = True          # TF32 + 2xBF16 forward products allowed (the model clears it for deep stacks, see asn_stacked_hg)
        self.pre = []              # launches that run BEFORE the weight repack (derived weights: Plan.head_comb)
        self.aux_grad = {}         # id(derived weight / bias tensor) -> its gradient tensor (not part of any flat store)
        self.aux_zero = []         # ... which are zeroed at the start of every backward pass
        self.no_pack_dep = set()   # id() of launches that read neither packed-weight arena (see _NO_PACK_DEP)
        self.rw_override = {}      # id(launch) -> (read pointers, written pointers) for launches that take arena base pointers
        self.after = {}            # id(launch) -> id(earlier launch) it must additionally wait for (scheduling-only edge)

    # ---- allocation helpers ----
    def buf(self, *shape):
        t = torch.empty(*shape, device=self.device, dtype=torch.float32)
        self.bytes_alloc += t.numel() * 4
        self.keep.append(t)
        return t

    def vec_alloc(self, C):
        n = (C + 3) // 4 * 4
        if self.vec_used + n > self.VEC_CAP:
            raise HGKError("plan vector arena exhausted")
        v = self.vec[self.vec_used:self.vec_used + C]
        self.vec_used += n
        return v

    def _stat_alloc(self, which, C):
        n = (C + 1) // 2 * 2
        if which == "f":
            if self.stat_f_used + n > self.STAT_CAP:
                raise HGKError("plan statistics arena exhausted")
            v = self.stat_f[self.stat_f_used:self.stat_f_used + C]
            self.stat_f_used += n
        else:
            if self.stat_b_used + n > self.STAT_CAP:
                raise HGKError("plan statistics arena exhausted")
            v = self.stat_b[self.stat_b_used:self.stat_b_used + C]
            self.stat_b_used += n
        return v

    def ticket_alloc(self):
        if self.tickets_used >= self.tickets.numel():
            raise HGKError("plan ticket arena exhausted")
        ptr = self.tickets.data_ptr() + 4 * self.tickets_used
        self.tickets_used += 1
        return ptr

    def defer_scope(self, anchor_rec):
        """Context manager: the forward launches appended inside are low priority AND may not start before `anchor_rec`
        (an earlier launch) has finished -- a scheduling-only edge with which the big skip-branch kernels are held back
        until the down/up chain reaches its small rungs, so that they run UNDER that latency chain instead of before it."""
        plan = self

        class _Scope(object):
            def __enter__(self_):
                self_.start = len(plan.fwd)

            def __exit__(self_, *exc):
                for rec in plan.fwd[self_.start:]:
                    plan.low_recs.add(id(rec))
                    if anchor_rec is not None:
                        plan.after[id(rec)] = id(anchor_rec)
                return False
        return _Scope()

    def low_scope(self):
        """Context manager: forward launches appended inside are marked as off the critical path (the hourglass skip
        branches: ready early, needed late).  The multi-stream scheduler confines them to the low-priority streams so
        that they fill the SMs the 16x16 / 8x8 / 4x4 rungs of the down/up chain leave idle instead of racing ahead."""
        plan = self

        class _Scope(object):
            def __enter__(self_):
                self_.start = len(plan.fwd)

            def __exit__(self_, *exc):
                for rec in plan.fwd[self_.start:]:
                    plan.low_recs.add(id(rec))
                return False
        return _Scope()

    def launch(self, lst, fn_name, *args):
        fn = getattr(self.lib, fn_name)
        rec = [fn, list(args), fn_name]
        lst.append(rec)
        return rec

    def dynamic(self, key, rec, idx):
        self.dyn.setdefault(key, []).append((rec, idx))

    # ---- parameters ----
    def param_ptr(self, p):
        return p.data_ptr()

    def param_grad_ptr(self, p):
        g = self.aux_grad.get(id(p))
        if g is not None:
            return g.data_ptr()
        for st in self.stores:
            i = st.index.get(id(p))
            if i is not None:
                return st.grad_ptr(i)
        raise KeyError("parameter is not part of any flat store of this plan")

    def packed_weight(self, w, mode):
        """Device pointer (resolved at finish()) of the repacked copy of OIHW weight `w`.
        mode 0: [tap][I][O]  (forward operand), mode 1: [tap][O][I] (data-gradient operand)."""
        O, I = w.shape[0], w.shape[1]
        taps = w.shape[2] * w.shape[3]
        for e in self.pack_entries:
            if e[0] is w and e[5] == mode:
                return e[1]
        off = self.pack_used
        self.pack_used += (O * I * taps + _ALIGN - 1) // _ALIGN * _ALIGN
        self.pack_entries.append((w, off, O, I, taps, mode))
        return off

    def wgrad_scratch(self, w):
        off = self.wg_used
        self.wg_used += (w.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.wg_entries.append((w, off))
        return _WgRef(off)

    def packed_weight_tc(self, w, mode, need_lo):
        """(hi, lo) references of the UMMA-layout copy of OIHW weight `w` (see hgk_pack_weights_tc).
        mode 0: forward operand (N=O, K=I); mode 1: data-gradient operand (N=I, K=O, taps flipped); mode 2: forward operand
        whose "lo" buffer holds the bf16 cross-term operands of the TF32 + 2xBF16 kernel (hgk_conv_tc_bn_x2_nhwc)."""
        O, I = w.shape[0], w.shape[1]
        BN = O if mode != 1 else I
        n = (w.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        for e in self.tc_entries:
            if e[0] is w and e[1] == mode:
                if need_lo and e[4] < 0:
                    e[4] = self.tc_used
                    self.tc_used += n
                return _TcRef(e[3]), (_TcRef(e[4]) if need_lo else 0)
        hi = self.tc_used
        self.tc_used += n
        lo = -1
        if need_lo:
            lo = self.tc_used
            self.tc_used += n
        self.tc_entries.append([w, mode, BN, hi, lo])
        return _TcRef(hi), (_TcRef(lo) if need_lo else 0)

    # ---- BN record ----
    def bn_rec(self, bn):
        r = BNRec()
        r.gamma, r.beta, r.rmean, r.rvar = bn.weight, bn.bias, bn.running_mean, bn.running_var
        r.training = self.training
        C = r.C = bn.weight.numel()
        r.scale, r.shift = self.vec_alloc(C), self.vec_alloc(C)
        r.mean, r.invstd = self.vec_alloc(C), self.vec_alloc(C)
        if self.training:
            r.sum, r.sq = self._stat_alloc("f", C), self._stat_alloc("f", C)
        else:
            r.sum = r.sq = None
        if self.need_grad:
            r.sum_g, r.sum_gx = self._stat_alloc("b", C), self._stat_alloc("b", C)
            r.cA, r.cB, r.cC = self.vec_alloc(C), self.vec_alloc(C), self.vec_alloc(C)
        self.bn_recs.append(r)
        if getattr(bn, "num_batches_tracked", None) is not None and self.training:
            self.nbt_bufs.append(bn.num_batches_tracked)
        if not self.training:
            self.launch(self.fwd, "bn_eval_prepare", _ptr(r.gamma), _ptr(r.beta), _ptr(r.rmean), _ptr(r.rvar),
                        BN_EPS, _ptr(r.scale), _ptr(r.shift), _ptr(r.mean), _ptr(r.invstd), C)
        return r

    # ---- gradient bookkeeping (resolved at plan-build time) ----
    def contribute(self, t, buf, donatable):
        if t is not None and t.needs_grad:
            t.contribs.append((buf, donatable))
            t.n_done += 1

    def grad_target(self, t, allow_res=True):
        """(buffer, accumulate flag, residual buffer or None) for a kernel about to write d/d t."""
        res = None
        t.n_done += 1
        if t.grad is None:
            for i, (b, don) in enumerate(t.contribs):
                if don:
                    t.grad = b
                    t.contribs.pop(i)
                    break
            if t.grad is not None:
                acc = 1
            else:
                t.grad = self.buf(t.N, t.H, t.W, t.C)
                acc = 0
        else:
            acc = 1
        if allow_res and t.contribs:
            res = t.contribs.pop(0)[0]
        return t.grad, acc, res

    def finalize_grad(self, t):
        """Flush pending contributions; returns the complete gradient buffer of t (or None)."""
        if t.grad is None:
            for i, (b, don) in enumerate(t.contribs):
                if don:
                    t.grad = b
                    t.contribs.pop(i)
                    break
        n = t.P * t.C
        while t.contribs:
            b, _ = t.contribs.pop(0)
            if t.grad is None:
                t.grad = self.buf(t.N, t.H, t.W, t.C)
                self.launch(self.bwd, "add_into", _ptr(b), _ptr(t.grad), n, 0)
            else:
                self.launch(self.bwd, "add_into", _ptr(b), _ptr(t.grad), n, 1)
        return t.grad

    # ---- ops ----
    def input_image(self, N, H, W):
        """Raw NCHW [N,3,H,W] user tensor (pointer patched per call); only the stem reads it."""
        t = T(None, N, H, W, 3, name="image")
        self.inputs.append(("image", t))
        return t

    def input_nchw(self, N, C, H, W, needs_grad):
        """NCHW feature tensor at the module boundary -> NHWC plan tensor."""
        z = self.buf(N, H, W, C)
        t = T(z, N, H, W, C, needs_grad=needs_grad and self.need_grad, name="in%d" % len(self.inputs))
        key = "in%d" % len(self.inputs)
        rec = self.launch(self.fwd, "nchw_to_nhwc", 0, N, C, H, W, _ptr(z))
        self.dynamic(key, rec, 0)
        self.inputs.append((key, t))
        op = _InputOp(self, t, key)
        t.producer = op
        self.tape.append(op)
        return t

    def stem(self, img, conv, bn):
        op = _StemOp(self, img, conv, bn)
        self.tape.append(op)
        return op.out

    def conv(self, x, conv, bn=None, res=None, relu=True):
        op = _ConvOp(self, x, conv, bn, res, relu)
        self.tape.append(op)
        return op.out

    def head_comb(self, forth_conv, in_conv, out_conv):
        """`forth_conv(y) + in_conv(out_conv(y))` (models/asn_stacked_hg.py:332-334) folded into ONE C->C convolution:
        returns a conv-like object (weight Wc = Wf + Wi Wo, bias bc = bf + bi + Wi bo, recomputed by a launch in front of the
        weight repack) to be passed to `conv(y, comb, res=x)`; the backward redistributes dWc / dbc exactly (csrc/heads.cu)."""
        op = _HeadCombOp(self, forth_conv, in_conv, out_conv)
        self.tape.append(op)
        return op

    def maxpool(self, x):
        op = _PoolOp(self, x)
        self.tape.append(op)
        return op.out

    def add(self, a, b, upsample_a=False):
        op = _AddOp(self, a, b, upsample_a)
        self.tape.append(op)
        return op.out

    def avgpool(self, x, k):
        op = _AvgPoolOp(self, x, k)
        self.tape.append(op)
        return op.out

    def linear(self, x, fc):
        op = _LinearOp(self, x, fc)
        self.tape.append(op)
        return op.out

    def materialize(self, x):
        """relu(bn(z)) written out as a plain NHWC tensor (AvgPool2d(1): the 1x1 window of the pooling kernel)."""
        return self.avgpool(x, 1) if x.scale is not None else x

    def input_plain(self, N, H, W, C, key):
        """A no-gradient NHWC input read in place through a per-call pointer (e.g. the [N,1,4,4] dropout masks, whose
        NCHW and NHWC layouts coincide because C == 1)."""
        t = T(None, N, H, W, C, name=key)
        self.inputs.append((key, t))
        return t

    def dropout(self, x, mask):
        """x * nearest_upsample(mask) (models/asn_stacked_hg.py:79-100); mask: [N,MH,MW,1] plan tensor, no gradient."""
        op = _DropoutOp(self, x, mask)
        self.tape.append(op)
        return op.out

    def sample_mask(self, pred):
        """Host sampling of the dropout cells (models/asn_stacked_hg.py:102-136) between two launches of the forward
        list: returns the [N,MH,MW,1] mask tensor (no gradient); the chosen indexes are left in `plan.mask_indexes`."""
        op = _SampleMaskOp(self, pred)
        self.tape.append(op)
        return op.out

    def output_plain(self, x):
        """A plain plan tensor whose NHWC buffer IS the NCHW result (C == 1 or H == W == 1), e.g. the ASN dropout head's
        [N,1,4,4] mask logits."""
        op = _OutputOp(self, x, len(self.outputs), rows=True)
        self.tape.append(op)
        self.outputs.append(op)
        return op

    def detach(self, x):
        """Same values, no gradient flow (x.detach(), models/asn_stacked_hg.py:161)."""
        x.use()      # a detached reader still means "more than one consumer" for fusion decisions
        return T(x.z, x.N, x.H, x.W, x.C, x.scale, x.shift, x.relu, needs_grad=False, name=x.name + ".detached")

    def target_nchw(self, N, C, H, W):
        """NCHW ground-truth heat-maps (pointer patched per call) -> NHWC, no gradient."""
        z = self.buf(N, H, W, C)
        rec = self.launch(self.fwd, "nchw_to_nhwc", 0, N, C, H, W, _ptr(z))
        self.dynamic("target", rec, 0)
        t = T(z, N, H, W, C, name="target")
        self.inputs.append(("target", t))
        return t

    def mse_loss(self, o, target, loss_acc, gscale=1.0):
        """loss_acc (fp64 scalar) += mean((o - target)^2); the gradient 2(o-t)/numel is produced in
        the same pass (inline loss of stack-hg.py:156-159)."""
        op = _MSEOp(self, o, target, loss_acc, gscale)
        self.tape.append(op)
        return op

    def output_nchw(self, x, no_grad=False):
        """Materialise T(x) as a contiguous NCHW tensor handed to the caller."""
        op = _OutputOp(self, x, len(self.outputs), no_grad=no_grad)
        self.tape.append(op)
        self.outputs.append(op)
        return op

    def output_rows(self, x):
        """[N,1,1,C] plan tensor returned as an [N,C] torch tensor (ASN logits)."""
        op = _OutputOp(self, x, len(self.outputs), rows=True)
        self.tape.append(op)
        self.outputs.append(op)
        return op

    # ---- finish / run ----
    def finish(self, grad_splits=None):
        """Close the plan: emit the backward list and resolve the scratch / packed-weight references.

        `grad_splits` (sorted element offsets into the flat gradient buffer, at parameter-slot boundaries; or a callable
        plan -> offsets evaluated once the backward list exists) cuts the 3x3
        weight-gradient unpack into one launch per gradient bucket, placed right behind the last weight-gradient kernel of
        that bucket, so that a bucket of the flat buffer is final long before the end of the backward pass (the bucketed
        all-reduce of HourglassTrainer overlaps the remaining backward)."""
        if self.need_grad:
            for op in reversed(self.tape):
                op.emit_bwd()
        if callable(grad_splits):          # chosen from the emitted backward list (trainer.plan_bucket_splits)
            grad_splits = grad_splits(self)
        self.grad_splits = grad_splits
        if self.wg_entries:
            import bisect
            self.wg_buf = torch.zeros(self.wg_used, device=self.device, dtype=torch.float32)
            self.bytes_alloc += self.wg_used * 4
            gbase = min(st.grad.data_ptr() for st in self.stores)
            wb = self.wg_buf.data_ptr()
            groups = {}
            for (w, off) in self.wg_entries:
                d = self.param_grad_ptr(w) - gbase
                assert d >= 0 and d % 4 == 0
                g = bisect.bisect_right(grad_splits, d // 4) if grad_splits else 0
                groups.setdefault(g, []).append([off, d // 4, w.shape[0], w.shape[1], w.shape[2] * w.shape[3]])
            order = sorted(groups)
            rows = [r for g in order for r in groups[g]]
            self.wg_table = torch.tensor(rows, dtype=torch.long, device=self.device)
            place, start = [], 0
            for g in order:
                offs = set(r[0] for r in groups[g])
                last = max(i for i, rec in enumerate(self.bwd)
                           if any(isinstance(a, _WgRef) and a.off in offs for a in rec[1]))
                rec = [self.lib.unpack_add_grads,
                       [wb, gbase, self.wg_table.data_ptr() + start * 5 * 8, len(groups[g])], "unpack_add_grads"]
                if grad_splits:
                    # precise dependencies instead of "ordered against everything" (schedule_streams)
                    self.rw_override[id(rec)] = ([wb + 4 * r[0] for r in groups[g]], [gbase + 4 * r[1] for r in groups[g]])
                place.append((last if grad_splits else len(self.bwd) - 1, rec))
                start += len(groups[g])
            for last, rec in sorted(place, key=lambda t: -t[0]):
                self.bwd.insert(last + 1, rec)
            for rec in self.bwd:
                rec[1] = [wb + 4 * a.off if isinstance(a, _WgRef) else a for a in rec[1]]
        if self.pack_entries:
            self.pack_buf = torch.empty(self.pack_used, device=self.device, dtype=torch.float32)
            self.bytes_alloc += self.pack_used * 4
            base = min(st.flat.data_ptr() for st in self.stores)
            rows = []
            for (w, off, O, I, taps, mode) in self.pack_entries:
                d = w.data_ptr() - base            # signed: derived weights (Plan.head_comb) live outside the flat stores
                assert d % 4 == 0
                rows.append([d // 4, off, O, I, taps, mode])
            self.pack_table = torch.tensor(rows, dtype=torch.long, device=self.device)
            self.pack_launch = [self.lib.pack_weights,
                                [base, self.pack_buf.data_ptr(), self.pack_table.data_ptr(), len(rows)], "pack_weights"]
            pb = self.pack_buf.data_ptr()
            # real read / write sets of the repack (schedule_streams): it reads the weights (the derived ones are written by
            # the launches of self.pre) and writes exactly the packed blocks the convolutions point at -- launches that read
            # no packed weight (the FFMA stem, the target transpose) are then free to run next to it
            self.rw_override[id(self.pack_launch)] = ([e[0].data_ptr() for e in self.pack_entries],
                                                      [pb + 4 * e[1] for e in self.pack_entries])
            for lst in (self.fwd, self.bwd):
                for rec in lst:
                    rec[1] = [pb + 4 * a.off if isinstance(a, _PackRef) else a for a in rec[1]]
        if self.tc_entries:
            self.tc_buf = torch.empty(self.tc_used, device=self.device, dtype=torch.float32)
            self.bytes_alloc += self.tc_used * 4
            base = min(st.flat.data_ptr() for st in self.stores)
            rows = []
            for (w, mode, BN, hi, lo) in self.tc_entries:
                d = w.data_ptr() - base            # signed, see above
                assert d % 4 == 0
                O, I, taps = w.shape[0], w.shape[1], w.shape[2] * w.shape[3]
                N, K = (O, I) if mode != 1 else (I, O)
                rows.append([d // 4, hi, lo, N, K, taps, mode, BN])
            self.tc_table = torch.tensor(rows, dtype=torch.long, device=self.device)
            self.tc_launch = [self.lib.pack_weights_tc,
                              [base, self.tc_buf.data_ptr(), self.tc_table.data_ptr(), len(rows)], "pack_weights_tc"]
            tb = self.tc_buf.data_ptr()
            self.rw_override[id(self.tc_launch)] = ([e[0].data_ptr() for e in self.tc_entries],
                                                    [tb + 4 * o for e in self.tc_entries for o in (e[3], e[4]) if o >= 0])
            for lst in (self.fwd, self.bwd):
                for rec in lst:
                    rec[1] = [tb + 4 * a.off if isinstance(a, _TcRef) else a for a in rec[1]]
        # num_batches_tracked: one flat add per store when every BN of that store is in training mode
        ids = set(id(b) for b in self.nbt_bufs)
        for st in self.stores:
            if st.ibufs and all(id(b) in ids for b in st.ibufs):
                self.nbt_flat.append(st.ibuf_flat)
                mine = set(id(b) for b in st.ibufs)
                self.nbt_bufs = [b for b in self.nbt_bufs if id(b) not in mine]
        self.finished = True

    def step_launches(self):
        """head + forward + backward in ISSUE order for the dependency-scheduled step: the forward launches in front of the
        first one that reads packed weights (target transpose, FFMA stem) are issued BEFORE the two weight repacks.  The
        hardware dispatches grids of one priority in launch order, so the persistent stem grid goes first and the repack
        blocks fill the SMs next to it, instead of the stem waiting ~60 us behind them."""
        packs = [r for r in (self.pack_launch, self.tc_launch) if r is not None]
        k = 0
        while k < len(self.fwd) and (self.fwd[k][2] in _NO_PACK_DEP or id(self.fwd[k]) in self.no_pack_dep):
            k += 1
        return list(self.pre) + self.fwd[:k] + packs + self.fwd[k:] + self.bwd

    def head_launches(self):
        """The launches in front of the forward list: derived weights, then the two weight repacks."""
        lst = list(self.pre)
        if self.pack_launch is not None:
            lst.append(self.pack_launch)
        if self.tc_launch is not None:
            lst.append(self.tc_launch)
        return lst

    def _run(self, lst, stream):
        arm = self.lib.pdl_arm
        prev_kernel = False
        for fn, args, name in lst:
            # programmatic dependent launch is only legal behind another kernel launch of the same stream (hgk.h)
            arm(1 if prev_kernel else 0)           # (always set: a launch that does not consume the flag must not leak it)
            prev_kernel = name != "host_sample_mask"
            rc = fn(*args, stream)
            if rc != 0:
                raise HGKError("%s failed (%d): %s" % (name, rc, self.lib.last_error()))

    def patch(self, key, ptr):
        for rec, idx in self.dyn.get(key, ()):
            rec[1][idx] = ptr

    def run_forward(self, inputs):
        """inputs: list of contiguous fp32 CUDA tensors in the order they were declared."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for (key, t), x in zip(self.inputs, inputs):
            self.patch(key, x.data_ptr())
        if self.stat_f_used:
            self.stat_f[:self.stat_f_used].zero_()
        self._run(self.head_launches(), stream)
        self._run(self.fwd, stream)
        for f in self.nbt_flat:
            f.add_(1)
        for b in self.nbt_bufs:
            b.add_(1)
        self.fwd_count += 1
        return [op.result for op in self.outputs]

    def run_backward(self, inputs, gouts):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for (key, t), x in zip(self.inputs, inputs):
            self.patch(key, x.data_ptr())
        if self.stat_b_used:
            self.stat_b[:self.stat_b_used].zero_()
        if self.wg_buf is not None:
            self.wg_buf.zero_()
        for a in self.aux_zero:
            a.zero_()
        held = []
        for op, g in zip(self.outputs, gouts):
            if op.no_grad:
                continue
            if g is None:
                op.gsrc.zero_()
                self.patch("gout%d" % op.index, op.gsrc.data_ptr())
            else:
                g = g.contiguous()
                if g.dtype != torch.float32:
                    g = g.float()
                held.append(g)
                self.patch("gout%d" % op.index, g.data_ptr())
        self._run(self.bwd, stream)
        return [op.gin for op in self.tape if isinstance(op, _InputOp)]


# argument positions written by each entry point (everything else is read-only); used by schedule_streams
_WRITES = {
    "conv_nhwc": (17, 19, 20), "conv_tc_nhwc": (17, 19, 20), "conv_tc_x2_nhwc": (17, 19, 20), "conv_tc_dgrad_bnstats_nhwc": (10, 18, 19),
    "conv_tc_bn_nhwc": (17, 19, 20, 25, 26, 27, 28, 29, 30, 31), "conv_tc_bn_x2_nhwc": (17, 19, 20, 25, 26, 27, 28, 29, 30, 31),
    "conv_tc_dgrad_bnfin_nhwc": (10, 18, 19, 22, 23, 24, 25, 26, 27),
    "bn_bwd_reduce_fin": (9, 10, 13, 14, 15, 16, 17, 18),
    "conv_tc_dgrad_bnapply_nhwc": (9, 18, 26, 27, 30, 31, 32, 33, 34, 35),
    "conv_wgrad_nhwc": (11, 15), "conv_wgrad_tc_nhwc": (11, 12),
    "bn_finalize": (7, 8, 9, 10, 11, 12), "bn_eval_prepare": (5, 6, 7, 8),
    "bn_bwd_reduce": (9, 10), "bn_bwd_finalize": (7, 8, 9, 10, 11), "bn_bwd_apply": (0,),
    "maxpool2_fwd": (8,), "maxpool2_bwd": (9,), "add_fwd": (13,), "upsample2_bwd": (5,), "add_into": (1,),
    "maxpool2_bwd_bnred": (9, 13, 14, 17, 18, 19, 20, 21, 22), "upsample2_bwd_bnred": (5, 13, 14, 17, 18, 19, 20, 21, 22),
    "nchw_to_nhwc": (5,), "nhwc_to_nchw": (8,),
    "stem_s2d_image": (4,), "stem_s2d_weight": (2,),
    "stem_conv7_fwd": (7, 8, 9), "stem_conv7_wgrad": (6, 7), "stem_conv7_wgrad_bnapply": (14, 15),
    "head_combine_fwd": (6, 7), "head_combine_bwd": (4, 5, 6, 7, 8, 9),
    "mse_fwd_bwd": (5, 7), "avgpool_fwd": (9,), "avgpool_bwd": (6,), "linear_fwd": (6,), "linear_bwd": (6, 7, 8),
}


# entry points that read neither packed-weight arena: they need not wait for the weight repack at the head of the step
# (the stem convolution reads the raw OIHW weights, the target transpose no weights at all)
_NO_PACK_DEP = ("stem_conv7_fwd", "stem_s2d_image", "nchw_to_nhwc")


def schedule_streams(launches, n_streams=4, barrier_names=("pack_weights", "pack_weights_tc", "unpack_add_grads"),
                     low_names=(), n_low=0, low_ids=(), after=None, rw_override=None, no_pack_ids=()):
    """Assign each launch of a static list to one of `n_streams` streams.

    Dependencies are derived from the launch arguments (device pointers; `_WRITES` says which positions
    an entry point writes): read-after-write, write-after-write and write-after-read orderings are kept,
    read-read sharing is free.  Launches named in `barrier_names` take base pointers of whole arenas and are
    ordered against everything (unless `rw_override` gives their real read / write pointer sets).  Returns (stream index per launch, for each launch the launches on OTHER
    streams it has to wait for).  The hourglass has coarse branch parallelism (skip residuals vs the down/up
    chain, weight- vs data-gradients): this lets the latency-bound 4x4 / 8x8 / 16x16 layers run under the
    large 64x64 ones inside one CUDA graph.

    `low_names` / `n_low`: launches with these entry-point names (the weight gradients: nothing on the critical path
    waits for them) are confined to the LAST `n_low` streams, which the caller creates with a lower priority, so that
    a data-gradient kernel whose CTAs are waiting for SMs is dispatched before the next weight-gradient CTA."""
    PTR_MIN = 1 << 32
    last_write = {}                    # ptr -> launch index of the last writer
    readers = {}                       # ptr -> launch indices that read it since that write
    tail = [-1] * n_streams            # last launch index per stream
    stream_of, cross = [], []
    barrier = -1
    lru = list(range(n_streams))
    waited = [[-1] * n_streams for _ in range(n_streams)]   # waited[k][kd]: newest launch of stream kd that k already waits for
    index_of = {}
    for i, (fn, args, name) in enumerate(launches):
        index_of[id(launches[i])] = i
        wpos = _WRITES.get(name)
        rd, wr = [], []
        override = rw_override.get(id(launches[i])) if rw_override else None
        if override is not None:            # a launch over arena base pointers whose real read / write sets are known
            rd, wr = list(override[0]), list(override[1])
        else:
            for j, a in enumerate(args):
                if isinstance(a, int) and a >= PTR_MIN:
                    (wr if (wpos is None or j in wpos) else rd).append(a)
        is_barrier = name in barrier_names and override is None
        if is_barrier:
            deps = set(t for t in tail if t >= 0)
        else:
            deps = set()
            for p in rd:
                if p in last_write:
                    deps.add(last_write[p])
            for p in wr:
                if p in last_write:
                    deps.add(last_write[p])
                deps.update(readers.get(p, ()))
            if barrier >= 0 and not ((name in _NO_PACK_DEP or id(launches[i]) in no_pack_ids)
                                     and launches[barrier][2] in ("pack_weights", "pack_weights_tc")):
                deps.add(barrier)
            if after:
                anchor = index_of.get(after.get(id(launches[i])))
                if anchor is not None and anchor < i:
                    deps.add(anchor)
        deps.discard(i)
        if n_low > 0:
            is_low = name in low_names or id(launches[i]) in low_ids
            pool = range(n_streams - n_low, n_streams) if is_low else range(0, n_streams - n_low)
        else:
            pool = range(n_streams)
        k = None
        for d in sorted(deps, reverse=True):       # continue on the stream of the most recent dependency
            if tail[stream_of[d]] == d and stream_of[d] in pool:
                k = stream_of[d]
                break
        if k is None:
            cand = [q for q in lru if q in pool]
            free = [q for q in cand if tail[q] < 0]
            k = free[0] if free else cand[0]
        lru.remove(k)
        lru.append(k)
        best = {}
        for d in deps:
            kd = stream_of[d]
            if kd != k and d > waited[k][kd]:
                best[kd] = max(best.get(kd, -1), d)
        for kd, d in best.items():
            waited[k][kd] = d
        cross.append(sorted(best.values()))
        stream_of.append(k)
        tail[k] = i
        for p in rd:
            readers.setdefault(p, []).append(i)
        for p in wr:
            last_write[p] = i
            readers[p] = []
        if is_barrier:
            barrier = i
    return stream_of, cross


class _PackRef(object):
    __slots__ = ("off",)

    def __init__(self, off):
        self.off = off


class _TcRef(object):
    __slots__ = ("off",)

    def __init__(self, off):
        self.off = off


class _WgRef(object):
    __slots__ = ("off",)

    def __init__(self, off):
        self.off = off


class _InputOp(object):
    def __init__(self, plan, t, key):
        self.plan, self.t, self.key = plan, t, key
        self.gin = None

    def emit_bwd(self):
        p, t = self.plan, self.t
        if not t.needs_grad:
            return
        g = p.finalize_grad(t)
        if g is None:
            return
        self.gin = torch.empty(t.N, t.C, t.H, t.W, device=p.device, dtype=torch.float32)
        p.launch(p.bwd, "nhwc_to_nchw", _ptr(g), 0, 0, 0, t.N, t.H, t.W, t.C, _ptr(self.gin))


class _StemOp(object):
    """conv1 7x7 s2 + bn1 (+relu on load), models/asn_stacked_hg.py:283-285."""

    def __init__(self, plan, img, conv, bn):
        self.plan, self.img, self.conv = plan, img, conv
        p = plan
        N, H, W = img.N, img.H, img.W
        Cout = conv.weight.shape[0]
        if H % 2 or W % 2:
            raise ValueError("stem needs even H, W (got %dx%d)" % (H, W))
        if tuple(conv.weight.shape[1:]) != (3, 7, 7) or Cout != 64:
            raise ValueError("stem conv must be Conv2d(3,64,7,stride=2,padding=3)")
        z = p.buf(N, H // 2, W // 2, Cout)
        self.bn = r = p.bn_rec(bn)
        H2, W2 = H // 2, W // 2
        fin_fused = False
        if (p.use_tc and os.environ.get("HGK_STEM_TC", "0") == "1" and H2 % 16 == 0 and W2 % 16 == 0
                and N * H2 * W2 * 64 < (1 << 32)):
            # HGK_STEM_TC=1 (off by default: measured on the config-2 step the launch itself is faster -- 205 us FFMA against
            # 14.5 + 139 us -- but the step is 0.06 ms SLOWER, 10.15 vs 10.08 ms, profiles/r6_stem_ab.txt).
            # tensor cores: space-to-depth turns the 7x7 stride-2 convolution into 4x4 taps over the 16-channel half-resolution
            # image (csrc/stem.cu), which the image-tile tcgen05 kernel runs as ksize = 4 with the usual fp32-class products,
            # bias, batch statistics and (when fused) the BatchNorm finaliser in its epilogue.  The rearranged weight is a
            # derived tensor rebuilt in front of the weight repack every step, like the folded head (Plan.head_comb).
            xs = p.buf(N, H2, W2, 16)
            self.ws = torch.zeros(Cout, 32, 4, 4, device=p.device, dtype=torch.float32)
            rec = p.launch(p.fwd, "stem_s2d_image", 0, N, H, W, _ptr(xs))
            p.launch(p.pre, "stem_s2d_weight", p.param_ptr(conv.weight), Cout, _ptr(self.ws))
            # its own (one-entry) repack into its own buffer: the stem is the first kernel of the step and must not wait for the
            # repack of all the other weights (the FFMA stem reads the raw weights and never did)
            nw = self.ws.numel()
            self.wpk = torch.zeros(2 * nw, device=p.device, dtype=torch.float32)
            self.wtab = torch.tensor([[0, 0, nw, Cout, 32, 16, 0, Cout]], dtype=torch.long, device=p.device)
            rec_pk = p.launch(p.pre, "pack_weights_tc", _ptr(self.ws), _ptr(self.wpk), _ptr(self.wtab), 1)
            hi, lo = _ptr(self.wpk), _ptr(self.wpk) + 4 * nw
            p.rw_override[id(rec_pk)] = ([_ptr(self.ws)], [hi, lo])
            # 3xTF32 products here even where the other layers take TF32 + 2xBF16: an error in the first layer is amplified by
            # every BatchNorm / ReLU of the network behind it (measured on the headline shape: whole-net gradient 7.8e-3 of
            # fp64 with bf16 cross terms in the stem, against the 2 x 3.2e-3 + 1e-3 gate), and the stem is not tensor-bound
            x2 = False
            stats = r.training
            base = [_ptr(xs), 0, 0, 0, N, H2, W2, 16, hi, lo, 4, p.param_ptr(conv.bias), Cout, 0, 0, 0, 0]
            if stats and p.fuse_bn_fin:
                crec = p.launch(p.fwd, "conv_tc_bn_x2_nhwc" if x2 else "conv_tc_bn_nhwc",
                                *(base + [_ptr(z), 0, _ptr(r.sum), _ptr(r.sq), _ptr(r.gamma), _ptr(r.beta), BN_EPS, BN_MOMENTUM,
                                          _ptr(r.rmean), _ptr(r.rvar), _ptr(r.scale), _ptr(r.shift), _ptr(r.mean), _ptr(r.invstd),
                                          p.ticket_alloc()]))
                fin_fused = True
            else:
                crec = p.launch(p.fwd, "conv_tc_x2_nhwc" if x2 else "conv_tc_nhwc",
                                *(base + [_ptr(z), 0, _ptr(r.sum) if stats else 0, _ptr(r.sq) if stats else 0]))
            p.no_pack_dep.add(id(crec))
        else:
            rec = p.launch(p.fwd, "stem_conv7_fwd", 0, N, H, W, p.param_ptr(conv.weight), p.param_ptr(conv.bias), Cout,
                           _ptr(z), _ptr(r.sum), _ptr(r.sq))
        p.dynamic("image", rec, 0)
        self.out = T(z, N, H // 2, W // 2, Cout, r.scale, r.shift, True, needs_grad=p.need_grad, name="stem")
        self.out.producer = self
        self.out.bn = r
        if r.training and not fin_fused:
            p.launch(p.fwd, "bn_finalize", _ptr(r.sum), _ptr(r.sq), self.out.P, _ptr(r.gamma), _ptr(r.beta), BN_EPS,
                     BN_MOMENTUM, _ptr(r.rmean), _ptr(r.rvar), _ptr(r.scale), _ptr(r.shift), _ptr(r.mean),
                     _ptr(r.invstd), Cout)

    def emit_bwd(self):
        p, o, r = self.plan, self.out, self.bn
        g = p.finalize_grad(o)
        if g is None:
            return
        if p.fuse_bn_apply:
            # BatchNorm-backward apply evaluated on load by the weight-gradient kernel (it is the only reader of dz: the stem
            # has no data gradient)
            _emit_bn_bwd(p, o, r, g, with_apply=False)
            rec = p.launch(p.bwd, "stem_conv7_wgrad_bnapply", 0, self.img.N, self.img.H, self.img.W, _ptr(g), _ptr(o.z),
                           _ptr(r.scale), _ptr(r.shift), int(o.relu), _ptr(r.mean), _ptr(r.cA), _ptr(r.cB), _ptr(r.cC), o.C,
                           p.param_grad_ptr(self.conv.weight), p.param_grad_ptr(self.conv.bias))
        else:
            _emit_bn_bwd(p, o, r, g)
            rec = p.launch(p.bwd, "stem_conv7_wgrad", 0, self.img.N, self.img.H, self.img.W, _ptr(g), o.C,
                           p.param_grad_ptr(self.conv.weight), p.param_grad_ptr(self.conv.bias))
        p.dynamic("image", rec, 0)


def _emit_bn_bwd(p, o, r, g, with_apply=True):
    """dY (buffer g, gradient w.r.t. relu(bn(z))) -> dz in place (with_apply=False: only the reduction / finaliser; the
    caller's data-gradient kernel evaluates the apply on load)."""
    fin_args = [_ptr(r.gamma), int(r.training), p.param_grad_ptr(r.gamma), p.param_grad_ptr(r.beta), _ptr(r.cA), _ptr(r.cB),
                _ptr(r.cC)]
    fin_done = o.bwd_fin_fused
    if not o.bwd_stats_fused:      # otherwise the consumer's data-gradient kernel already accumulated sum_g / sum_gx
        red_args = [_ptr(g), _ptr(o.z), _ptr(r.scale), _ptr(r.shift), int(o.relu), _ptr(r.mean), _ptr(r.invstd), o.P, o.C,
                    _ptr(r.sum_g), _ptr(r.sum_gx)]
        if p.fuse_bn_fin:
            p.launch(p.bwd, "bn_bwd_reduce_fin", *(red_args + fin_args + [p.ticket_alloc()]))
            fin_done = True
        else:
            p.launch(p.bwd, "bn_bwd_reduce", *red_args)
    if not fin_done:
        p.launch(p.bwd, "bn_bwd_finalize", _ptr(r.sum_g), _ptr(r.sum_gx), o.P, _ptr(r.gamma), _ptr(r.mean), _ptr(r.invstd),
                 int(r.training), p.param_grad_ptr(r.gamma), p.param_grad_ptr(r.beta), _ptr(r.cA), _ptr(r.cB), _ptr(r.cC), o.C)
    if with_apply:
        p.launch(p.bwd, "bn_bwd_apply", _ptr(g), _ptr(o.z), _ptr(r.scale), _ptr(r.shift), int(o.relu), _ptr(r.mean), _ptr(r.cA),
                 _ptr(r.cB), _ptr(r.cC), o.P, o.C)


class _ConvOp(object):
    """nn.Conv2d 1x1 / 3x3(p1) [+ `out += shortcut`] [+ nn.BatchNorm2d + ReLU applied by the consumers]."""

    def __init__(self, plan, x, conv, bn, res, relu):
        self.plan, self.x, self.conv, self.res = plan, x, conv, res
        x.use()
        if res is not None:
            res.use()
        p = plan
        w = conv.weight
        Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
        if k not in (1, 3) or w.shape[3] != k:
            raise ValueError("only 1x1 and 3x3 convolutions are supported (got %dx%d)" % (w.shape[2], w.shape[3]))
        if Cin != x.C:
            raise ValueError("conv expects %d input channels, got %d" % (Cin, x.C))
        if Cin % 4 or Cout % 4:
            raise ValueError("channel counts must be multiples of 4 (Cin=%d, Cout=%d)" % (Cin, Cout))
        if res is not None and (res.C != Cout or res.P != x.P):
            raise ValueError("shortcut shape mismatch")
        self.k, self.Cin, self.Cout = k, Cin, Cout
        z = p.buf(x.N, x.H, x.W, Cout)
        self.bn = r = p.bn_rec(bn) if bn is not None else None
        stats = r is not None and r.training
        ra = res.act_args() if res is not None else [0, 0, 0, 0]
        fin_fused = False
        if p.use_tc and p.lib.conv_tc_supported(Cin, Cout, k):
            # tcgen05 tensor cores, error-compensated 3xTF32 (fp32-class accuracy)
            # 3x3 layers on the image-tile kernel: TF32 + 2xBF16 products (4 instead of 6 tensor-core instructions per product
            # group, same fp32-class accuracy class; csrc/conv_tc2.cu)
            x2 = (p.x2_ok and not p.precise_grads                      # (PRECISE_GRADS: the all-3xTF32 / fp32 mode)
                  and p.lib.conv_tc_x2_supported(x.N, x.H, x.W, Cin, Cout, k))
            hi, lo = p.packed_weight_tc(w, 2 if x2 else 0, True)
            base = x.act_args() + [x.N, x.H, x.W, Cin, hi, lo, k, p.param_ptr(conv.bias), Cout] + ra
            if stats and p.fuse_bn_fin:
                # the last CTA of the convolution finalises the BatchNorm (scale/shift, running statistics)
                p.launch(p.fwd, "conv_tc_bn_x2_nhwc" if x2 else "conv_tc_bn_nhwc", *(base + [_ptr(z), 0, _ptr(r.sum), _ptr(r.sq), _ptr(r.gamma), _ptr(r.beta),
                                                             BN_EPS, BN_MOMENTUM, _ptr(r.rmean), _ptr(r.rvar), _ptr(r.scale),
                                                             _ptr(r.shift), _ptr(r.mean), _ptr(r.invstd), p.ticket_alloc()]))
                fin_fused = True
            else:
                p.launch(p.fwd, "conv_tc_x2_nhwc" if x2 else "conv_tc_nhwc",
                         *(base + [_ptr(z), 0, _ptr(r.sum) if stats else 0, _ptr(r.sq) if stats else 0]))
        else:
            p.launch(p.fwd, "conv_nhwc", *(x.act_args() + [x.N, x.H, x.W, Cin, _PackRef(p.packed_weight(w, 0)), k, 0,
                                                           p.param_ptr(conv.bias), Cout] + ra +
                                           [_ptr(z), 0, _ptr(r.sum) if stats else 0, _ptr(r.sq) if stats else 0, 0]))
        if r is not None:
            self.out = T(z, x.N, x.H, x.W, Cout, r.scale, r.shift, relu, needs_grad=p.need_grad, name="conv")
            self.out.bn = r
            if r.training and not fin_fused:
                p.launch(p.fwd, "bn_finalize", _ptr(r.sum), _ptr(r.sq), self.out.P, _ptr(r.gamma), _ptr(r.beta), BN_EPS,
                         BN_MOMENTUM, _ptr(r.rmean), _ptr(r.rvar), _ptr(r.scale), _ptr(r.shift), _ptr(r.mean),
                         _ptr(r.invstd), Cout)
        else:
            self.out = T(z, x.N, x.H, x.W, Cout, needs_grad=p.need_grad, name="conv")
        self.out.producer = self

    def emit_bwd(self):
        p, o, x = self.plan, self.out, self.x
        g = p.finalize_grad(o)
        if g is None:
            return
        w = self.conv.weight
        k, Cin, Cout = self.k, self.Cin, self.Cout
        dgrad_tc = x.needs_grad and p.use_tc and p.lib.conv_tc_supported(Cout, Cin, k)
        # BatchNorm-backward apply on load: the data-gradient kernel reads (g, z), forms dz on the fly and writes it once
        # to `dzb` for the weight gradient and the shortcut -- no bn_bwd_apply pass over the tensor
        ap = (self.bn is not None and dgrad_tc and p.fuse_bn_apply and not p.precise_grads
              and p.lib.conv_tc_bnapply_supported(x.N, x.H, x.W, Cout, Cin, k))
        if self.bn is not None:
            _emit_bn_bwd(p, o, self.bn, g, with_apply=not ap)
        dzb = p.buf(o.N, o.H, o.W, o.C) if ap else g

        def emit_wgrad():
            if p.use_tc and not p.precise_grads and p.lib.conv_wgrad_tc_supported(Cin, Cout, k):
                # tensor cores; 1x1: tap-major == OIHW, accumulate straight into .grad; 3x3: via the tap-major scratch
                dst = p.param_grad_ptr(w) if k == 1 else p.wgrad_scratch(w)
                p.launch(p.bwd, "conv_wgrad_tc_nhwc", *(x.act_args() + [x.N, x.H, x.W, Cin, _ptr(dzb), Cout, k, dst,
                                                                        p.param_grad_ptr(self.conv.bias)]))
            else:
                # fp32 SIMT kernel, weight / bias gradient straight into the OIHW .grad views
                p.launch(p.bwd, "conv_wgrad_nhwc", *(x.act_args() + [x.N, x.H, x.W, Cin, _ptr(dzb), Cout, k,
                                                                     p.param_grad_ptr(w), Cin * k * k, k * k, 1,
                                                                     p.param_grad_ptr(self.conv.bias)]))
        if not ap:
            emit_wgrad()
        # data gradient: same kernel on dz with the [tap][Cout][Cin] weights, taps flipped
        if x.needs_grad:
            gx, acc, extra = p.grad_target(x)
            if dgrad_tc:
                hi, lo = p.packed_weight_tc(w, 1, p.precise_grads)   # plain TF32 is enough for gradients
                fuse_red = (x.bn is not None and x.n_done == x.n_cons and not x.contribs and p.fuse_bn_bwd
                            and not p.precise_grads)
                # fuse_red: every consumer of x has contributed and this launch folds the rest in (accumulate / extra): it
                # produces the COMPLETE dL/d relu(bn(z)) of the previous layer, so that BatchNorm's backward
                # reduction (sum g, sum g*xhat) is fused into the epilogue
                r = x.bn
                red = ([_ptr(x.z), _ptr(x.scale), _ptr(x.shift), int(x.relu), _ptr(r.mean), _ptr(r.invstd), _ptr(r.sum_g),
                        _ptr(r.sum_gx)] if fuse_red else None)
                fin = ([_ptr(r.gamma), int(r.training), p.param_grad_ptr(r.gamma), p.param_grad_ptr(r.beta), _ptr(r.cA),
                        _ptr(r.cB), _ptr(r.cC)] if fuse_red else None)
                if ap:
                    ro = self.bn
                    head = [_ptr(g), _ptr(o.z), _ptr(ro.scale), _ptr(ro.shift), int(o.relu), _ptr(ro.mean), _ptr(ro.cA),
                            _ptr(ro.cB), _ptr(ro.cC), _ptr(dzb), x.N, x.H, x.W, Cout, hi, k, Cin, _ptr(extra), _ptr(gx), acc]
                    if fuse_red:
                        p.launch(p.bwd, "conv_tc_dgrad_bnapply_nhwc", *(head + red + fin + [p.ticket_alloc()]))
                        x.bwd_stats_fused = x.bwd_fin_fused = True
                    else:
                        p.launch(p.bwd, "conv_tc_dgrad_bnapply_nhwc", *(head + [0] * 16))
                elif fuse_red:
                    args = [_ptr(g), x.N, x.H, x.W, Cout, hi, lo, k, Cin, _ptr(extra), _ptr(gx), acc] + red
                    if p.fuse_bn_fin:
                        # ... and its finaliser (dgamma, dbeta, cA/cB/cC) runs in the last CTA
                        p.launch(p.bwd, "conv_tc_dgrad_bnfin_nhwc", *(args + fin + [p.ticket_alloc()]))
                        x.bwd_fin_fused = True
                    else:
                        p.launch(p.bwd, "conv_tc_dgrad_bnstats_nhwc", *args)
                    x.bwd_stats_fused = True
                else:
                    p.launch(p.bwd, "conv_tc_nhwc", _ptr(g), 0, 0, 0, x.N, x.H, x.W, Cout, hi, lo, k, 0, Cin,
                             _ptr(extra), 0, 0, 0, _ptr(gx), acc, 0, 0)
            else:
                wref = p.param_ptr(w) if k == 1 else _PackRef(p.packed_weight(w, 1))
                p.launch(p.bwd, "conv_nhwc", _ptr(g), 0, 0, 0, x.N, x.H, x.W, Cout, wref, k, 1, 0, Cin,
                         _ptr(extra), 0, 0, 0, _ptr(gx), acc, 0, 0, 0)
        if ap:
            emit_wgrad()          # reads the dz written by the data-gradient kernel
        # shortcut: d/d res = dz (donated: this op never touches the buffer again)
        if self.res is not None:
            p.contribute(self.res, dzb, True)


class _HeadCombOp(object):
    """Derived 1x1 convolution Wc = Wf + Wi Wo, bc = bf + bi + Wi bo (see Plan.head_comb); quacks like nn.Conv2d for
    _ConvOp (.weight, .bias).  Sits in the tape BEFORE the convolution that uses it, so that its backward (which turns the
    accumulated dWc / dbc into the gradients of the three real layers) is emitted AFTER that convolution's weight gradient."""

    def __init__(self, plan, forth_conv, in_conv, out_conv):
        self.plan, self.forth, self.inc, self.outc = plan, forth_conv, in_conv, out_conv
        C, J = forth_conv.weight.shape[0], out_conv.weight.shape[0]
        if (tuple(forth_conv.weight.shape) != (C, C, 1, 1) or tuple(in_conv.weight.shape) != (C, J, 1, 1)
                or tuple(out_conv.weight.shape) != (J, C, 1, 1)):
            raise ValueError("head_comb expects 1x1 convolutions C->C, J->C and C->J")
        self.C, self.J = C, J
        dev = plan.device
        self.weight = torch.empty(C, C, 1, 1, device=dev, dtype=torch.float32)
        self.bias = torch.empty(C, device=dev, dtype=torch.float32)
        plan.keep += [self.weight, self.bias]
        plan.launch(plan.pre, "head_combine_fwd", _ptr(forth_conv.weight), _ptr(forth_conv.bias), _ptr(in_conv.weight),
                    _ptr(in_conv.bias), _ptr(out_conv.weight), _ptr(out_conv.bias), _ptr(self.weight), _ptr(self.bias), C, J)
        if plan.need_grad:
            self.dweight, self.dbias = torch.zeros_like(self.weight), torch.zeros_like(self.bias)
            plan.aux_grad[id(self.weight)] = self.dweight
            plan.aux_grad[id(self.bias)] = self.dbias
            plan.aux_zero += [self.dweight, self.dbias]

    def emit_bwd(self):
        p = self.plan
        g = lambda t: p.param_grad_ptr(t) if t is not None else 0
        p.launch(p.bwd, "head_combine_bwd", _ptr(self.dweight), _ptr(self.dbias), _ptr(self.inc.weight), _ptr(self.outc.weight),
                 g(self.forth.weight), g(self.forth.bias), g(self.inc.weight), g(self.inc.bias), g(self.outc.weight),
                 g(self.outc.bias), self.C, self.J)


def _last_contribution(p, t):
    """True when the launch about to be emitted writes the LAST contribution to the gradient of the BatchNorm-carrying tensor t
    (every consumer has contributed, nothing is pending) and the BatchNorm-backward reduction can ride on it."""
    return (t.bn is not None and t.scale is not None and t.n_done == t.n_cons and not t.contribs and p.fuse_bn_bwd
            and p.fuse_bn_fin and t.C % 4 == 0 and 256 % (t.C // 4) == 0)


def _bn_red_args(p, t):
    r = t.bn
    return [_ptr(r.mean), _ptr(r.invstd), _ptr(r.sum_g), _ptr(r.sum_gx), _ptr(r.gamma), int(r.training),
            p.param_grad_ptr(r.gamma), p.param_grad_ptr(r.beta), _ptr(r.cA), _ptr(r.cB), _ptr(r.cC), p.ticket_alloc()]


class _PoolOp(object):
    """nn.MaxPool2d(2, 2) of a virtual activation, models/asn_stacked_hg.py:69,142-154,287."""

    def __init__(self, plan, x):
        self.plan, self.x = plan, x
        x.use()
        if x.H % 2 or x.W % 2:
            raise ValueError("max-pool input must have even H, W (got %dx%d)" % (x.H, x.W))
        z = plan.buf(x.N, x.H // 2, x.W // 2, x.C)
        plan.launch(plan.fwd, "maxpool2_fwd", *(x.act_args() + [x.N, x.H, x.W, x.C, _ptr(z)]))
        self.out = T(z, x.N, x.H // 2, x.W // 2, x.C, needs_grad=x.needs_grad, name="pool")
        self.out.producer = self

    def emit_bwd(self):
        p, x = self.plan, self.x
        g = p.finalize_grad(self.out)
        if g is None or not x.needs_grad:
            return
        gx, acc, _ = p.grad_target(x, allow_res=False)
        if _last_contribution(p, x):
            # this launch completes dL/d relu(bn(z)) of the pooled tensor: its BatchNorm-backward reduction + finaliser ride along
            p.launch(p.bwd, "maxpool2_bwd_bnred", *(x.act_args() + [x.N, x.H, x.W, x.C, _ptr(g), _ptr(gx), acc] + _bn_red_args(p, x)))
            x.bwd_stats_fused = x.bwd_fin_fused = True
        else:
            p.launch(p.bwd, "maxpool2_bwd", *(x.act_args() + [x.N, x.H, x.W, x.C, _ptr(g), _ptr(gx), acc]))


class _AddOp(object):
    """`upsample(x) + skip` (models/asn_stacked_hg.py:193-203) or a plain add (:408-417, :334)."""

    def __init__(self, plan, a, b, up):
        self.plan, self.a, self.b, self.up = plan, a, b, up
        a.use()
        b.use()
        if a.C != b.C or a.N != b.N or (a.H * (2 if up else 1), a.W * (2 if up else 1)) != (b.H, b.W):
            raise ValueError("add: shape mismatch")
        z = plan.buf(b.N, b.H, b.W, b.C)
        plan.launch(plan.fwd, "add_fwd", *(a.act_args() + [int(up)] + b.act_args() + [b.N, b.H, b.W, b.C, _ptr(z)]))
        self.out = T(z, b.N, b.H, b.W, b.C, needs_grad=a.needs_grad or b.needs_grad, name="add")
        self.out.producer = self

    def emit_bwd(self):
        p, a, b = self.plan, self.a, self.b
        g = p.finalize_grad(self.out)
        if g is None:
            return
        if a.needs_grad:
            ga, acc, _ = p.grad_target(a, allow_res=False)
            if self.up and _last_contribution(p, a):
                p.launch(p.bwd, "upsample2_bwd_bnred", _ptr(g), b.N, b.H, b.W, b.C, _ptr(ga), acc, _ptr(a.z), _ptr(a.scale),
                         _ptr(a.shift), int(a.relu), *_bn_red_args(p, a))
                a.bwd_stats_fused = a.bwd_fin_fused = True
            elif self.up:
                p.launch(p.bwd, "upsample2_bwd", _ptr(g), b.N, b.H, b.W, b.C, _ptr(ga), acc)
            else:
                p.launch(p.bwd, "add_into", _ptr(g), _ptr(ga), a.P * a.C, acc)
        p.contribute(b, g, True)


class _AvgPoolOp(object):
    """nn.AvgPool2d(k), models/asn_stacked_hg.py:375,431."""

    def __init__(self, plan, x, k):
        self.plan, self.x, self.k = plan, x, k
        x.use()
        if x.H < k or x.W < k:
            raise ValueError("avg-pool kernel %d larger than input %dx%d" % (k, x.H, x.W))
        z = plan.buf(x.N, x.H // k, x.W // k, x.C)
        plan.launch(plan.fwd, "avgpool_fwd", *(x.act_args() + [x.N, x.H, x.W, x.C, k, _ptr(z)]))
        self.out = T(z, x.N, x.H // k, x.W // k, x.C, needs_grad=x.needs_grad, name="avgpool")
        self.out.producer = self

    def emit_bwd(self):
        p, x = self.plan, self.x
        g = p.finalize_grad(self.out)
        if g is None or not x.needs_grad:
            return
        gx, acc, _ = p.grad_target(x, allow_res=False)
        p.launch(p.bwd, "avgpool_bwd", _ptr(g), x.N, x.H, x.W, x.C, self.k, _ptr(gx), acc)


class _LinearOp(object):
    """nn.Linear on [N, C] rows (fc_scale / fc_rotation, models/asn_stacked_hg.py:376-377,434-435)."""

    def __init__(self, plan, x, fc):
        self.plan, self.x, self.fc = plan, x, fc
        x.use()
        if x.scale is not None:
            raise ValueError("linear expects a plain (materialised) tensor")
        Nout, K = fc.weight.shape[0], fc.weight[0].numel()      # nn.Linear [Nout,K] or a 1x1 nn.Conv2d [Nout,K,1,1]
        if K != x.C:
            raise ValueError("linear expects %d features, got %d" % (K, x.C))
        z = plan.buf(x.N, x.H, x.W, Nout)
        plan.launch(plan.fwd, "linear_fwd", _ptr(x.z), plan.param_ptr(fc.weight), plan.param_ptr(fc.bias), x.P, K, Nout,
                    _ptr(z))
        self.out = T(z, x.N, x.H, x.W, Nout, needs_grad=plan.need_grad, name="linear")
        self.out.producer = self

    def emit_bwd(self):
        p, x, fc = self.plan, self.x, self.fc
        g = p.finalize_grad(self.out)
        if g is None:
            return
        Nout, K = fc.weight.shape[0], fc.weight[0].numel()
        gx = 0
        if x.needs_grad:
            tmp = p.buf(x.N, x.H, x.W, K)
            gx = _ptr(tmp)
            p.contribute(x, tmp, True)
        p.launch(p.bwd, "linear_bwd", _ptr(x.z), p.param_ptr(fc.weight), _ptr(g), x.P, K, Nout, gx,
                 p.param_grad_ptr(fc.weight), p.param_grad_ptr(fc.bias))


class _DropoutOp(object):
    """`_Hourglass._dropout` (models/asn_stacked_hg.py:79-100): multiply the activated tensor by the nearest-upsampled
    [N,1,4,4] mask.  The product is materialised (its consumers read a plain tensor)."""

    def __init__(self, plan, x, mask):
        self.plan, self.x, self.mask = plan, x, mask
        x.use()
        if mask.C != 1 or mask.N != x.N or x.H % mask.H or x.W % mask.W or x.H // mask.H != x.W // mask.W:
            raise ValueError("dropout: mask [%d,%d,%d,%d] does not tile the %dx%d feature map"
                             % (mask.N, mask.H, mask.W, mask.C, x.H, x.W))
        z = plan.buf(x.N, x.H, x.W, x.C)
        rec = plan.launch(plan.fwd, "mask_mul_fwd", *(x.act_args() + [_ptr(mask.z), x.N, x.H, x.W, x.C, mask.H, mask.W, _ptr(z)]))
        if mask.z is None:
            plan.dynamic(mask.name, rec, 4)
        self.out = T(z, x.N, x.H, x.W, x.C, needs_grad=x.needs_grad, name="dropout")
        self.out.producer = self

    def emit_bwd(self):
        p, x, m = self.plan, self.x, self.mask
        g = p.finalize_grad(self.out)
        if g is None or not x.needs_grad:
            return
        gx, acc, _ = p.grad_target(x, allow_res=False)
        rec = p.launch(p.bwd, "mask_mul_bwd", _ptr(g), _ptr(m.z), x.N, x.H, x.W, x.C, m.H, m.W, _ptr(gx), acc)
        if m.z is None:
            p.dynamic(m.name, rec, 1)


class _SampleMaskOp(object):
    """`_Hourglass._sample_mask` (models/asn_stacked_hg.py:102-136): softmax over the MH*MW mask logits of every sample,
    then `np.random.choice(MH*MW, 2, p=probs[i], replace=False)` on the HOST, exactly as the reference (same numpy
    RandomState stream); the two chosen cells are zeroed in an all-ones mask.  Runs between two kernel launches of the
    forward list (one stream synchronisation, as the reference's `.cpu()`); the softmax itself is a libhgk kernel."""

    DROPOUT_NUM = 2

    def __init__(self, plan, pred):
        import numpy as np
        self.plan, self.pred = plan, pred
        pred.use()
        if pred.C != 1 or pred.scale is not None or pred.H != pred.W:
            raise ValueError("sample_mask expects plain [N,H,W,1] mask logits with H == W")
        N, K = pred.N, pred.H * pred.W
        dev = plan.device
        self.probs = torch.empty(N, K, device=dev, dtype=torch.float32)
        self.mask = plan.buf(N, pred.H, pred.W, 1)
        self.host_probs = torch.empty(N, K, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.empty(N, K)
        self.host_mask = torch.empty(N, pred.H, pred.W, 1, dtype=torch.float32)
        if dev.type == "cuda":
            self.host_mask = self.host_mask.pin_memory()
        plan.mask_indexes = torch.zeros(N, self.DROPOUT_NUM, dtype=torch.long)
        plan.launch(plan.fwd, "softmax_sample", _ptr(pred.z), N, K, 0, _ptr(self.probs), 0)
        lib = plan.lib

        def host_sample(stream):
            self.host_probs.copy_(self.probs, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            probs = self.host_probs.numpy()
            self.host_mask.fill_(1.0)
            hm = self.host_mask.view(N, K)
            for i in range(N):
                idx = np.random.choice(K, self.DROPOUT_NUM, p=probs[i], replace=False)      # ref:122
                for j in range(len(idx)):
                    hm[i, int(idx[j])] = 0.0                                                 # ref:124-128
                    plan.mask_indexes[i, j] = int(idx[j])                                    # ref:131
            self.mask.copy_(self.host_mask, non_blocking=True)
            return 0
        plan.fwd.append([host_sample, [], "host_sample_mask"])
        self.out = T(self.mask, N, pred.H, pred.W, 1, name="dropout_mask")

    def emit_bwd(self):
        pass


class _MSEOp(object):
    def __init__(self, plan, o, target, loss_acc, gscale):
        self.plan, self.o = plan, o
        o.use()
        if o.scale is not None or (o.N, o.H, o.W, o.C) != (target.N, target.H, target.W, target.C):
            raise ValueError("mse_loss expects a plain output and a target of the same shape")
        n = o.P * o.C
        self.g = plan.buf(o.N, o.H, o.W, o.C) if plan.need_grad else None
        plan.launch(plan.fwd, "mse_fwd_bwd", _ptr(o.z), _ptr(target.z), n, 1.0 / n, gscale, _ptr(self.g), 0,
                    _ptr(loss_acc))

    def emit_bwd(self):
        if self.g is not None:
            self.plan.contribute(self.o, self.g, True)


class _OutputOp(object):
    def __init__(self, plan, x, index, rows=False, no_grad=False):
        self.plan, self.x, self.index, self.rows, self.no_grad = plan, x, index, rows, no_grad
        if not no_grad:
            x.use()
        p = plan
        if rows:
            if x.scale is not None or not ((x.H == 1 and x.W == 1) or x.C == 1):
                raise ValueError("row output expects a plain [N,1,1,C] or [N,H,W,1] tensor")
            self.result = x.z.view(x.N, x.C) if (x.H == 1 and x.W == 1) else x.z.view(x.N, 1, x.H, x.W)
            self.gsrc = torch.zeros_like(self.result) if (p.need_grad and not no_grad) else None
        else:
            self.result = torch.empty(x.N, x.C, x.H, x.W, device=p.device, dtype=torch.float32)
            p.bytes_alloc += self.result.numel() * 4
            p.launch(p.fwd, "nhwc_to_nchw", *(x.act_args() + [x.N, x.H, x.W, x.C, _ptr(self.result)]))
            self.gsrc = torch.zeros_like(self.result) if (p.need_grad and not no_grad) else None

    def emit_bwd(self):
        p, x = self.plan, self.x
        if not x.needs_grad or self.no_grad:
            return
        g = p.buf(x.N, x.H, x.W, x.C)
        if self.rows:
            rec = p.launch(p.bwd, "add_into", 0, _ptr(g), x.P * x.C, 0)
        else:
            rec = p.launch(p.bwd, "nchw_to_nhwc", 0, x.N, x.C, x.H, x.W, _ptr(g))
        p.dynamic("gout%d" % self.index, rec, 0)
        p.contribute(x, g, True)
