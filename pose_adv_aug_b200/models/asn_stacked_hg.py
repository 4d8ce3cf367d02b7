"""Drop-in mirror of the reference module `models/asn_stacked_hg.py` (zhiqiangdon/pose-adv-aug):
same constructors (`create_hg`, `create_asn`), class names, attribute names (hence identical
`state_dict()` keys and shapes), forward signatures and return structures -- but `forward`
plans and launches the hand-written sm_100a kernels of libhgk instead of torch ops.

Reference lines cited as `ref:<line>` refer to /root/reference/models/asn_stacked_hg.py.
torch.nn.Conv2d / BatchNorm2d / Linear objects are used only as *parameter containers* (their
own forward is never called): that keeps parameter registration, state_dict schema and
`isinstance` checks identical to the reference.
"""
import math

import torch
import torch.nn as nn

from ..engine import Plan, ParamStore, HGKError

__all__ = ["_Residual", "_Hourglass", "_Hourglass_Wrapper", "ASN", "create_hg", "create_asn", "Hourglass"]

# convolution kernel selection for newly built plans: 0 = tcgen05 tensor cores where the shape is covered
# (3xTF32 forward), 1 = fp32 SIMT kernels everywhere
CONV_PATH = 0
# False: gradients use plain TF32 operands on the tensor cores (fast; error below the whole-net fp32 noise
# floor).  True: 3xTF32 data gradients + fp32 weight gradients (fp32-class gradients, slower).
PRECISE_GRADS = False
# True: build the hourglass skip branches after the down chain and hold them back in the multi-stream schedule (see
# _Hourglass._build_down); a pure scheduling choice, results are identical (measured 12.27 -> 12.00 ms/step with 8
# streams, 3 of them low priority).  HGK_DEFER_SKIPS=0 restores the reference's build order.
import os as _os
DEFER_SKIPS = _os.environ.get("HGK_DEFER_SKIPS", "1") == "1"
# Inter-stack head: forth_conv(y) + in_conv(out_conv(y)) as ONE C->C convolution with the combined weights Wf + Wi Wo
# (engine.Plan.head_comb, csrc/heads.cu): the 16->C pass over the activation, its data gradient and its weight gradient
# become weight-space products.  HGK_FUSE_HEAD=0 restores the three separate convolutions of ref:332-334.
FUSE_HEAD = _os.environ.get("HGK_FUSE_HEAD", "1") == "1"
# Forward products: TF32 + 2xBF16 (csrc/conv_tc2.cu, conv_tc3.cu) is ~3 * 2^-20 per product against ~2^-22 for 3xTF32.  A
# train-mode hourglass amplifies forward rounding by ~2x every 1.5 stacks (measured at 2 images, heat-map error relative to the
# map maximum vs the fp64 oracle, stacks 1..8: fp32 reference 1e-5 .. 4e-4, 3xTF32 2e-5 .. 5e-4, TF32 + 2xBF16 3e-5 .. 1.2e-3):
# inside the 1e-3 tolerance with a wide margin up to 4 stacks (2.4e-4), over it at 8.  Deeper nets therefore keep 3xTF32.
X2_MAX_STACKS = int(_os.environ.get("HGK_X2_MAX_STACKS", "4"))


def _reference_init(root):
    """ref:258-270 / ref:381-393: conv W,b ~ U(+-1/sqrt(k*k*Cin)); BN gamma ~ U(0,1), beta = 0."""
    for m in root.modules():
        if isinstance(m, nn.Conv2d):
            n = m.kernel_size[0] * m.kernel_size[1] * m.in_channels
            stdv = 1 / math.sqrt(n)
            m.weight.data.uniform_(-stdv, stdv)
            if m.bias is not None:
                m.bias.data.uniform_(-stdv, stdv)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.uniform_()
            m.bias.data.zero_()


# ------------------------------------------------------------------------------------------
# plan execution glue (autograd boundary)
# ------------------------------------------------------------------------------------------
class _PlanFn(torch.autograd.Function):
    """One autograd node for a whole planned forward.  Parameter gradients are accumulated by the
    kernels directly into the flat .grad views (ParamStore); input gradients are returned."""

    @staticmethod
    def forward(ctx, plan, n_in, *tensors):
        inputs = tensors[:n_in]
        outs = plan.run_forward(inputs)
        ctx.plan = plan
        ctx.inputs = inputs
        ctx.count = plan.fwd_count
        ctx.n_extra = len(tensors) - n_in
        return tuple(o.clone() for o in outs)

    @staticmethod
    def backward(ctx, *gouts):
        plan = ctx.plan
        if plan.fwd_count != ctx.count:
            raise HGKError("backward() after a newer forward of the same plan: the saved activations were "
                           "overwritten (run backward before the next forward of this module/shape)")
        for st in plan.stores:
            st.attach_grads()
        gins = plan.run_backward(ctx.inputs, gouts)
        gins = list(gins) + [None] * (len(ctx.inputs) - len(gins))
        return (None, None) + tuple(gins[:len(ctx.inputs)]) + (None,) * ctx.n_extra


class _Root(object):
    """Per-root-module state: flat parameter store + plan cache."""

    def __init__(self):
        self.store = None
        self.plans = {}


def _state(module):
    st = module.__dict__.get("_hgk_root")
    if st is None:
        st = _Root()
        module.__dict__["_hgk_root"] = st
    return st


def _ensure_store(module, device):
    st = _state(module)
    if st.store is None or st.store.device != device or not st.store.valid():
        st.store = ParamStore(module, device)
        st.plans = {}
    return st.store


def _check_input(x):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise HGKError("pose_adv_aug_b200 runs on CUDA (sm_100a) only; got a %s tensor -- there is no CPU fallback"
                       % (x.device if isinstance(x, torch.Tensor) else type(x)))
    if x.dtype != torch.float32:
        raise ValueError("expected a float32 tensor, got %s" % x.dtype)
    if x.dim() != 4:
        raise ValueError("expected an NCHW tensor, got %d dims" % x.dim())


def _run(root, extra_roots, key, inputs, build):
    """Build (or fetch) the plan `key` of `root` and run it through autograd."""
    for x in inputs:
        _check_input(x)
    device = inputs[0].device
    stores = [_ensure_store(m, device) for m in [root] + list(extra_roots)]
    mods = [root] + list(extra_roots)
    need_grad = torch.is_grad_enabled() and (any(x.requires_grad for x in inputs) or
                                             any(p.requires_grad for m in mods for p in m.parameters()))
    shapes = tuple(tuple(x.shape) for x in inputs)
    modes = tuple(m.training for m in mods)
    ids = tuple(id(s) for s in stores)
    # which inputs want a gradient is baked into the plan (input_nchw(needs_grad=...)): it is part of the key, so a module
    # first called on a constant input and later inside a larger autograd graph gets a plan with the input-gradient path
    in_grad = tuple(bool(x.requires_grad) and need_grad for x in inputs)
    pkey = (key, shapes, modes, need_grad, in_grad, ids, CONV_PATH, PRECISE_GRADS, FUSE_HEAD, X2_MAX_STACKS)
    cache = _state(root).plans
    plan = cache.get(pkey)
    if plan is None:
        plan = Plan(stores, device, root.training, need_grad, conv_path=CONV_PATH, precise_grads=PRECISE_GRADS)
        plan.in_requires_grad = [x.requires_grad for x in inputs]
        plan.structure = build(plan)
        plan.finish()
        cache[pkey] = plan
    inputs = [x.contiguous() for x in inputs]
    if need_grad:
        for st in stores:
            st.attach_grads()
        anchor = None
        for m in mods:
            for p in m.parameters():
                if p.requires_grad:
                    anchor = p
                    break
            if anchor is not None:
                break
        extra = (anchor,) if anchor is not None else ()
        outs = _PlanFn.apply(plan, len(inputs), *(tuple(inputs) + extra))
    else:
        outs = tuple(o.clone() for o in plan.run_forward(inputs))
    return plan, list(outs)


# ------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------
class _Residual(nn.Module):
    """ref:11-49 -- post-activation bottleneck: c1(1x1) bn relu c2(3x3) bn relu c3(1x1) (+x | +adapter(x)) bn relu."""

    def __init__(self, in_num, out_num, adapter=None):
        super(_Residual, self).__init__()
        self.in_num = in_num
        self.out_num = out_num
        self.conv1 = nn.Conv2d(in_num, out_num // 2, kernel_size=1, stride=1, bias=True)
        self.bn1 = nn.BatchNorm2d(out_num // 2)
        self.conv2 = nn.Conv2d(out_num // 2, out_num // 2, kernel_size=3, stride=1, padding=1, bias=True)
        self.bn2 = nn.BatchNorm2d(out_num // 2)
        self.conv3 = nn.Conv2d(out_num // 2, out_num, kernel_size=1, stride=1, bias=True)
        self.bn3 = nn.BatchNorm2d(out_num)
        self.relu = nn.ReLU(inplace=True)
        self.adapter = adapter

    def _build(self, plan, x):
        if self.adapter is None:
            if self.in_num != self.out_num:
                raise ValueError("identity shortcut needs in_num == out_num (ref:31-32)")
            shortcut = x
        else:
            shortcut = plan.conv(x, self.adapter)                       # ref:34
        out = plan.conv(x, self.conv1, bn=self.bn1)                     # ref:36-38
        out = plan.conv(out, self.conv2, bn=self.bn2)                   # ref:40-42
        return plan.conv(out, self.conv3, bn=self.bn3, res=shortcut)    # ref:44-47

    def forward(self, x):
        def build(plan):
            t = plan.input_nchw(*x.shape, needs_grad=x.requires_grad)
            plan.output_nchw(self._build(plan, t))
        return _run(self, (), "residual", [x], build)[1][0]


def _build_stack(seq, plan, x):
    for m in seq:
        x = m._build(plan, x)
    return x


class _Hourglass(nn.Module):
    """ref:51-213."""

    def __init__(self, chan, num_modules):
        super(_Hourglass, self).__init__()
        self.num_modules = num_modules
        self.chan = chan
        self.down1 = self._stack_residual()
        self.down2 = self._stack_residual()
        self.down3 = self._stack_residual()
        self.down4 = self._stack_residual()
        self.up1 = self._stack_residual()
        self.up2 = self._stack_residual()
        self.up3 = self._stack_residual()
        self.up4 = self._stack_residual()
        self.skip1 = self._stack_residual()
        self.skip2 = self._stack_residual()
        self.skip3 = self._stack_residual()
        self.skip4 = self._stack_residual()
        self.neck = self._stack_residual()
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.upsample = nn.Upsample(scale_factor=2)
        self.keys = ['neck', 'skip1', 'skip2', 'skip3', 'skip4']

    def _stack_residual(self):
        return nn.Sequential(*[_Residual(self.chan, self.chan) for _ in range(self.num_modules)])

    def _build_down(self, plan, x):
        """ref:140-157; returns (neck, skip1..4)."""
        if DEFER_SKIPS:
            # same graph, different build order: the down chain first, then the three big skip branches, held back (by a
            # scheduling-only edge in the multi-stream CUDA graph) until the chain enters its 16x16 rung: they then fill
            # the SMs while the 16x16 / 8x8 / 4x4 rungs run as a latency chain on a few SMs
            x64 = x
            x32 = _build_stack(self.down1, plan, plan.maxpool(x64))
            x16 = _build_stack(self.down2, plan, plan.maxpool(x32))
            anchor = plan.fwd[-1] if plan.fwd else None
            with plan.defer_scope(anchor):
                s1 = _build_stack(self.skip1, plan, x64)
                s2 = _build_stack(self.skip2, plan, x32)
                s3 = _build_stack(self.skip3, plan, x16)
            x = _build_stack(self.down3, plan, plan.maxpool(x16))
            with plan.low_scope():
                s4 = _build_stack(self.skip4, plan, x)
        else:
            with plan.low_scope():
                s1 = _build_stack(self.skip1, plan, x)
            x = _build_stack(self.down1, plan, plan.maxpool(x))
            with plan.low_scope():
                s2 = _build_stack(self.skip2, plan, x)
            x = _build_stack(self.down2, plan, plan.maxpool(x))
            with plan.low_scope():
                s3 = _build_stack(self.skip3, plan, x)
            x = _build_stack(self.down3, plan, plan.maxpool(x))
            with plan.low_scope():
                s4 = _build_stack(self.skip4, plan, x)
        x = _build_stack(self.down4, plan, plan.maxpool(x))
        x = _build_stack(self.neck, plan, x)
        return x, s1, s2, s3, s4

    def _build_up(self, plan, x, s1, s2, s3, s4):
        """ref:192-203: up residual -> nearest x2 -> + skip (one fused kernel per rung)."""
        x = plan.add(_build_stack(self.up4, plan, x), s4, upsample_a=True)
        x = plan.add(_build_stack(self.up3, plan, x), s3, upsample_a=True)
        x = plan.add(_build_stack(self.up2, plan, x), s2, upsample_a=True)
        x = plan.add(_build_stack(self.up1, plan, x), s1, upsample_a=True)
        return x

    def _build(self, plan, x, asn=None, is_half_hg=False, is_dropout=False, mask=None):
        """Returns (y or None, agent outputs or None, dropout mask tensor or None)."""
        neck, s1, s2, s3, s4 = self._build_down(plan, x)
        agent = None
        if asn is not None:
            assert mask is None                                                      # ref:160
            feats = {'neck': plan.detach(neck), 'skip1': plan.detach(s1), 'skip2': plan.detach(s2),
                     'skip3': plan.detach(s3), 'skip4': plan.detach(s4)}           # ref:161-162
            mode = plan.training
            plan.training = asn.training
            agent = asn._build(plan, feats, is_dropout=is_dropout)
            plan.training = mode
            if is_half_hg:
                return None, agent, None                                             # ref:169-171,174-176
            if is_dropout:
                mask = plan.sample_mask(agent[0])                                    # ref:178
        if mask is not None:                                                         # ref:179-190
            neck = plan.dropout(neck, mask)
            s1, s2, s3, s4 = (plan.dropout(t, mask) for t in (s1, s2, s3, s4))
        return self._build_up(plan, neck, s1, s2, s3, s4), agent, mask

    def forward(self, x, asn=None, is_half_hg=False, is_aug=False, is_dropout=False, dropout_masks=None):
        if asn is not None:
            assert dropout_masks is None                                             # ref:160
            assert is_aug != is_dropout                                              # ref:165
        inputs = [x] if dropout_masks is None else [x, dropout_masks]

        def build(plan):
            t = plan.input_nchw(*x.shape, needs_grad=x.requires_grad)
            mask = None
            if dropout_masks is not None:
                if dropout_masks.dim() != 4 or dropout_masks.shape[1] != 1:
                    raise ValueError("dropout_masks must be [N,1,h,w] (ref:81)")
                mask = plan.input_plain(dropout_masks.shape[0], dropout_masks.shape[2], dropout_masks.shape[3], 1, "dropout_masks")
            y, agent, m = self._build(plan, t, asn, is_half_hg, is_dropout, mask)
            if y is not None:
                plan.output_nchw(y)
            if agent is not None:
                for a in agent:
                    (plan.output_plain if is_dropout else plan.output_rows)(a)
            return m
        extra = (asn,) if asn is not None else ()
        key = ("hourglass", asn is not None, is_half_hg, is_dropout, dropout_masks is not None)
        plan, outs = _run(self, extra, key, inputs, build)
        if asn is None:
            return outs[0]
        if is_aug:
            if is_half_hg:
                return outs[0], outs[1]
            return outs[0], outs[1], outs[2]                                         # ref:208
        if is_half_hg:
            return outs[0]                                                           # ref:176
        m = plan.structure
        return outs[0], outs[1], plan.mask_indexes.clone(), m.z.view(m.N, 1, m.H, m.W).clone()   # ref:210


class _Hourglass_Wrapper(nn.Module):
    """ref:215-342."""

    def __init__(self, num_modules, num_stacks, chan=256, num_classes=16):
        super(_Hourglass_Wrapper, self).__init__()
        self.chan = chan
        self.num_modules = num_modules
        self.num_classes = num_classes
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=True)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.residual1 = self._make_adapter_residual(64, 128)
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.residual2 = _Residual(128, 128)
        self.residual3 = self._make_adapter_residual(128, chan)
        self.num_stacks = num_stacks
        hg, post_res, linear, out_conv, forth_conv, in_conv = [], [], [], [], [], []
        for i in range(0, num_stacks):
            hg.append(_Hourglass(chan=chan, num_modules=num_modules))
            post_res.append(self._stack_residual())
            linear.append(nn.Sequential(nn.Conv2d(chan, chan, kernel_size=1, stride=1, bias=True),
                                        nn.BatchNorm2d(chan), nn.ReLU(inplace=True)))
            out_conv.append(nn.Conv2d(chan, num_classes, kernel_size=1, stride=1, bias=True))
            if i < num_stacks - 1:
                forth_conv.append(nn.Conv2d(chan, chan, kernel_size=1, stride=1, bias=True))
                in_conv.append(nn.Conv2d(num_classes, chan, kernel_size=1, stride=1, bias=True))
        self.hg = nn.ModuleList(hg)
        self.post_res = nn.ModuleList(post_res)
        self.linear = nn.ModuleList(linear)
        self.out_conv = nn.ModuleList(out_conv)
        self.forth_conv = nn.ModuleList(forth_conv)
        self.in_conv = nn.ModuleList(in_conv)
        _reference_init(self)

    def _stack_residual(self):
        return nn.Sequential(*[_Residual(self.chan, self.chan) for _ in range(self.num_modules)])

    def _make_adapter_residual(self, in_num, out_num):
        adapter = nn.Conv2d(in_num, out_num, kernel_size=1, stride=1, bias=True)
        return _Residual(in_num, out_num, adapter)

    def _build(self, plan, img, asn=None, is_half_hg=False, is_dropout=False):
        """ref:282-342.  Returns (list of per-stack heat-map tensors, agent outputs or None)."""
        plan.x2_ok = self.num_stacks <= X2_MAX_STACKS                   # forward numerics by depth, see X2_MAX_STACKS
        x = plan.stem(img, self.conv1, self.bn1)                        # ref:283-285
        x = self.residual1._build(plan, x)
        x = plan.maxpool(x)
        x = self.residual2._build(plan, x)
        x = self.residual3._build(plan, x)
        outs, agent, mask = [], None, None
        for i in range(self.num_stacks):
            if i == 0 and asn is not None:
                y, agent, mask = self.hg[i]._build(plan, x, asn, is_half_hg, is_dropout)
                if is_half_hg:
                    return outs, agent                                  # ref:302-304,311-313
            else:
                y, _, _ = self.hg[i]._build(plan, x, mask=mask)         # ref:320-322: later stacks reuse the masks
            y = _build_stack(self.post_res[i], plan, y)                 # ref:327
            y = plan.conv(y, self.linear[i][0], bn=self.linear[i][1])   # ref:328
            fused = FUSE_HEAD and i < self.num_stacks - 1
            if fused:
                # ref:332-334 as one convolution, built BEFORE out_conv: the backward then meets out_conv's (SIMT) data gradient
                # first and this tensor-core one last, which carries y's BatchNorm-backward reduction in its epilogue
                comb = plan.head_comb(self.forth_conv[i], self.in_conv[i], self.out_conv[i])
                x = plan.conv(y, comb, res=x)
            o = plan.conv(y, self.out_conv[i])                          # ref:329
            outs.append(o)
            if i < self.num_stacks - 1 and not fused:
                t = plan.conv(y, self.forth_conv[i], res=x)             # ref:332,334
                x = plan.conv(o, self.in_conv[i], res=t)                # ref:333-334
        return outs, agent

    def forward(self, x, asn=None, is_half_hg=False, is_aug=False, is_dropout=False):
        if asn is not None:
            assert is_aug != is_dropout                                 # ref:297
        _check_input(x)
        if x.shape[1] != 3 or x.shape[2] % 64 or x.shape[3] % 64:
            raise ValueError("expected [N,3,H,W] images with H, W multiples of 64 (got %s)" % (tuple(x.shape),))
        if x.requires_grad and torch.is_grad_enabled():
            # the stem kernel has no image-gradient path (no script of the reference differentiates w.r.t. the image)
            raise HGKError("the input image requires grad, but this path produces no image gradient; pass x.detach()")

        def build(plan):
            img = plan.input_image(x.shape[0], x.shape[2], x.shape[3])
            outs, agent = self._build(plan, img, asn, is_half_hg, is_dropout and asn is not None)
            for o in outs:
                plan.output_nchw(o)
            if agent is not None:
                for a in agent:
                    (plan.output_plain if is_dropout else plan.output_rows)(a)
            return len(outs)
        extra = (asn,) if asn is not None else ()
        plan, outs = _run(self, extra, ("wrapper", asn is not None, is_half_hg, is_dropout), [x], build)
        n = plan.structure
        if asn is None:
            return outs[:n]                                             # ref:342
        if is_aug:
            if is_half_hg:
                return outs[n], outs[n + 1]                             # ref:304
            return outs[:n], outs[n], outs[n + 1]                       # ref:338
        if is_half_hg:
            return outs[n]                                              # ref:313
        return outs[:n], outs[n], plan.mask_indexes.clone()             # ref:340


def create_hg(num_stacks, num_modules, num_classes, chan):
    """ref:344-347."""
    return _Hourglass_Wrapper(num_stacks=num_stacks, num_modules=num_modules, num_classes=num_classes, chan=chan)


def Hourglass(nStack, nFeat, num_modules=1, num_classes=16):
    """north_star's `Hourglass(nStack, nFeat)` spelling of create_hg(nStack, 1, 16, nFeat)."""
    return create_hg(nStack, num_modules, num_classes, nFeat)


class ASN(nn.Module):
    """ref:349-439 -- the adversarial augmentation agent."""

    def __init__(self, chan_in, chan_out, scale_num, rotation_num, is_aug=False, is_dropout=False):
        assert is_aug != is_dropout                                     # ref:351
        super(ASN, self).__init__()
        self.num_modules = 3
        self.chan_in = chan_in
        self.chan_out = chan_out
        self.residual_skip1 = _Residual(chan_in, chan_out)
        self.residual_skip2 = _Residual(chan_in, chan_out)
        self.residual_skip3 = _Residual(chan_in, chan_out)
        self.residual_skip4 = _Residual(chan_in, chan_out)
        self.residual_neck = _Residual(chan_in, chan_out)
        self.merge1 = _Residual(chan_out, chan_out)
        self.merge2 = _Residual(chan_out, chan_out)
        self.merge3 = _Residual(chan_out, chan_out)
        self.merge4 = _Residual(chan_out, chan_out)
        self.deep_merge = self._stack_residual()
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        if is_aug:
            self.avgpool = nn.AvgPool2d(4)
            self.fc_scale = nn.Linear(chan_out, scale_num)
            self.fc_rotation = nn.Linear(chan_out, rotation_num)
        if is_dropout:
            self.out_conv = nn.Conv2d(chan_out, 1, kernel_size=1, stride=1, bias=True)
        _reference_init(self)

    def _stack_residual(self):
        return nn.Sequential(*[_Residual(self.chan_out, self.chan_out) for _ in range(self.num_modules)])

    def _build(self, plan, f, is_dropout=False):
        """ref:401-439.  f: dict of plan tensors.  Returns (scale, rotation) logits or (mask logits,)."""
        if is_dropout != hasattr(self, "out_conv"):
            raise ValueError("this ASN was created with is_%s=True (ref:351,400-405)" % ("dropout" if hasattr(self, "out_conv") else "aug"))
        skip1 = self.residual_skip1._build(plan, f['skip1'])
        skip2 = self.residual_skip2._build(plan, f['skip2'])
        skip3 = self.residual_skip3._build(plan, f['skip3'])
        skip4 = self.residual_skip4._build(plan, f['skip4'])
        neck = self.residual_neck._build(plan, f['neck'])
        x = self.merge1._build(plan, plan.add(plan.maxpool(skip1), skip2))
        x = self.merge2._build(plan, plan.add(plan.maxpool(x), skip3))
        x = self.merge3._build(plan, plan.add(plan.maxpool(x), skip4))
        x = self.merge4._build(plan, plan.add(plan.maxpool(x), neck))
        x = _build_stack(self.deep_merge, plan, x)
        if is_dropout:
            # ref:437-439: out_conv = Conv2d(chan_out, 1, 1x1): one dot product per cell of the 4x4 map
            return (plan.linear(plan.materialize(x), self.out_conv),)
        x = plan.avgpool(x, 4)                                          # ref:431
        if x.H != 1 or x.W != 1:
            raise ValueError("ASN head needs a 4x4 neck (256x256 input); got %dx%d after AvgPool2d(4)" % (x.H, x.W))
        return plan.linear(x, self.fc_scale), plan.linear(x, self.fc_rotation)   # ref:434-435

    def forward(self, x, is_aug=False, is_dropout=False):
        if is_aug == is_dropout:
            return None                                                 # ref:430-439: neither branch returns
        keys = ['neck', 'skip1', 'skip2', 'skip3', 'skip4']
        tensors = [x[k] for k in keys]

        def build(plan):
            f = dict((k, plan.input_nchw(*t.shape, needs_grad=t.requires_grad)) for k, t in zip(keys, tensors))
            for o in self._build(plan, f, is_dropout=is_dropout):
                (plan.output_plain if is_dropout else plan.output_rows)(o)
        outs = _run(self, (), ("asn", is_dropout), tensors, build)[1]
        return outs[0] if is_dropout else (outs[0], outs[1])


def create_asn(chan_in, chan_out, scale_num=None, rotation_num=None, is_aug=False, is_dropout=False):
    """ref:441-444."""
    return ASN(chan_in=chan_in, chan_out=chan_out, scale_num=scale_num, rotation_num=rotation_num,
               is_aug=is_aug, is_dropout=is_dropout)
