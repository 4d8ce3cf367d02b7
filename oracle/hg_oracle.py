"""CPU oracle for the stacked-hourglass (+ASN agent) hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package
(`pose_adv_aug_b200/`); only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may use it.

This is a *functional restatement* of the reference's algorithm: plain torch CPU
tensor arithmetic (any float dtype: fp32 = "the reference", fp64 = "truth") over a
name->tensor state dict, with each function citing the reference lines it follows
(paths relative to /root/reference).  It deliberately does not use nn.Module,
nn.BatchNorm2d or any code of the product package.

Parity pin: `tests/golden/*.npz` hold outputs of the *reference itself*
(`models/asn_stacked_hg.py`, made Python-3 runnable by `oracle/make_ref.py`) produced by
`oracle/gen_golden.py` in the build container; `tests/test_oracle.py` checks this
restatement against them (and against the live reference when `oracle/_ref/` is present).
The reference ships no tests / golden vectors of its own.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # torch.nn.BatchNorm2d default, used at models/asn_stacked_hg.py:19,22,25,224,243
BN_MOMENTUM = 0.1    # torch.nn.BatchNorm2d default


# --------------------------------------------------------------------------------------
# Schema: the parameter / buffer names and shapes the reference constructors register.
# --------------------------------------------------------------------------------------
def _conv_entries(prefix, cout, cin, k):
    return [(prefix + ".weight", (cout, cin, k, k)), (prefix + ".bias", (cout,))]


def _bn_entries(prefix, c):
    return [(prefix + ".weight", (c,)), (prefix + ".bias", (c,)),
            (prefix + ".running_mean", (c,)), (prefix + ".running_var", (c,)),
            (prefix + ".num_batches_tracked", ())]


def residual_schema(prefix, in_num, out_num, adapter):
    """_Residual.__init__, models/asn_stacked_hg.py:13-28 (adapter registered last, :28)."""
    m = out_num // 2
    e = []
    e += _conv_entries(prefix + ".conv1", m, in_num, 1) + _bn_entries(prefix + ".bn1", m)
    e += _conv_entries(prefix + ".conv2", m, m, 3) + _bn_entries(prefix + ".bn2", m)
    e += _conv_entries(prefix + ".conv3", out_num, m, 1) + _bn_entries(prefix + ".bn3", out_num)
    if adapter:
        e += _conv_entries(prefix + ".adapter", out_num, in_num, 1)
    return e


def _stack_schema(prefix, chan, num_modules):
    """_stack_residual, models/asn_stacked_hg.py:73-77 / :272-276 / :395-399."""
    e = []
    for i in range(num_modules):
        e += residual_schema("%s.%d" % (prefix, i), chan, chan, False)
    return e


HG_BRANCHES = ["down1", "down2", "down3", "down4", "up1", "up2", "up3", "up4",
               "skip1", "skip2", "skip3", "skip4", "neck"]   # registration order, :56-68


def hg_schema(num_stacks, num_modules, num_classes, chan):
    """_Hourglass_Wrapper.__init__, models/asn_stacked_hg.py:216-255 (state_dict order)."""
    e = []
    e += _conv_entries("conv1", 64, 3, 7) + _bn_entries("bn1", 64)
    e += residual_schema("residual1", 64, 128, True)
    e += residual_schema("residual2", 128, 128, False)
    e += residual_schema("residual3", 128, chan, True)
    for s in range(num_stacks):
        for b in HG_BRANCHES:
            e += _stack_schema("hg.%d.%s" % (s, b), chan, num_modules)
    for s in range(num_stacks):
        e += _stack_schema("post_res.%d" % s, chan, num_modules)
    for s in range(num_stacks):
        e += _conv_entries("linear.%d.0" % s, chan, chan, 1) + _bn_entries("linear.%d.1" % s, chan)
    for s in range(num_stacks):
        e += _conv_entries("out_conv.%d" % s, num_classes, chan, 1)
    for s in range(num_stacks - 1):
        e += _conv_entries("forth_conv.%d" % s, chan, chan, 1)
    for s in range(num_stacks - 1):
        e += _conv_entries("in_conv.%d" % s, chan, num_classes, 1)
    return e


def asn_schema(chan_in, chan_out, scale_num=None, rotation_num=None, is_aug=False, is_dropout=False):
    """ASN.__init__, models/asn_stacked_hg.py:350-379."""
    assert is_aug != is_dropout
    e = []
    for k in ("residual_skip1", "residual_skip2", "residual_skip3", "residual_skip4", "residual_neck"):
        e += residual_schema(k, chan_in, chan_out, False)
    for k in ("merge1", "merge2", "merge3", "merge4"):
        e += residual_schema(k, chan_out, chan_out, False)
    e += _stack_schema("deep_merge", chan_out, 3)
    if is_aug:
        e += [("fc_scale.weight", (scale_num, chan_out)), ("fc_scale.bias", (scale_num,)),
              ("fc_rotation.weight", (rotation_num, chan_out)), ("fc_rotation.bias", (rotation_num,))]
    if is_dropout:
        e += _conv_entries("out_conv", 1, chan_out, 1)
    return e


# --------------------------------------------------------------------------------------
# Layers
# --------------------------------------------------------------------------------------
class BNState(object):
    """Collects the running-stat updates a training-mode forward produces."""

    def __init__(self, training):
        self.training = training
        self.updates = OrderedDict()


def conv(sd, p, x, k):
    """nn.Conv2d(.., kernel_size=k, stride=1, padding=k//2, bias=True), :17,20,23,242-248,279."""
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=1, padding=k // 2)


def batchnorm(sd, p, x, st):
    """nn.BatchNorm2d forward (:19,22,25,224,243): batch statistics (biased variance) in
    training mode, running statistics in eval mode; running stats are updated with
    momentum 0.1 and the *unbiased* batch variance."""
    g = sd[p + ".weight"].view(1, -1, 1, 1)
    b = sd[p + ".bias"].view(1, -1, 1, 1)
    if st.training:
        n = x.numel() // x.size(1)
        mean = x.mean(dim=(0, 2, 3))
        var = ((x - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
        with torch.no_grad():
            unb = var * (float(n) / max(n - 1, 1))
            st.updates[p + ".running_mean"] = (1 - BN_MOMENTUM) * sd[p + ".running_mean"] + BN_MOMENTUM * mean
            st.updates[p + ".running_var"] = (1 - BN_MOMENTUM) * sd[p + ".running_var"] + BN_MOMENTUM * unb
            if (p + ".num_batches_tracked") in sd:
                st.updates[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
    else:
        mean = sd[p + ".running_mean"]
        var = sd[p + ".running_var"]
    xhat = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
    return xhat * g + b


def residual(sd, p, x, st):
    """_Residual.forward, models/asn_stacked_hg.py:30-49 (post-activation bottleneck)."""
    shortcut = conv(sd, p + ".adapter", x, 1) if (p + ".adapter.weight") in sd else x   # :31-34
    out = F.relu(batchnorm(sd, p + ".bn1", conv(sd, p + ".conv1", x, 1), st))              # :36-38
    out = F.relu(batchnorm(sd, p + ".bn2", conv(sd, p + ".conv2", out, 3), st))            # :40-42
    out = conv(sd, p + ".conv3", out, 1) + shortcut                                         # :44-45
    return F.relu(batchnorm(sd, p + ".bn3", out, st))                                       # :46-47


def stack(sd, p, x, st, num_modules):
    """nn.Sequential of num_modules residuals, :73-77."""
    for i in range(num_modules):
        x = residual(sd, "%s.%d" % (p, i), x, st)
    return x


def maxpool(x):
    """nn.MaxPool2d(kernel_size=2, stride=2), :69,227,371."""
    return F.max_pool2d(x, 2, 2)


def upsample(x):
    """nn.Upsample(scale_factor=2) -- nearest neighbour, :70."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def hourglass_down(sd, p, x, st, M):
    """_Hourglass.forward first half, :140-157.  Returns neck and the 4 skips."""
    s1 = stack(sd, p + ".skip1", x, st, M)
    x = stack(sd, p + ".down1", maxpool(x), st, M)
    s2 = stack(sd, p + ".skip2", x, st, M)
    x = stack(sd, p + ".down2", maxpool(x), st, M)
    s3 = stack(sd, p + ".skip3", x, st, M)
    x = stack(sd, p + ".down3", maxpool(x), st, M)
    s4 = stack(sd, p + ".skip4", x, st, M)
    x = stack(sd, p + ".down4", maxpool(x), st, M)
    x = stack(sd, p + ".neck", x, st, M)
    return x, s1, s2, s3, s4


def hourglass_up(sd, p, x, s1, s2, s3, s4, st, M):
    """_Hourglass.forward second half, :192-203."""
    x = upsample(stack(sd, p + ".up4", x, st, M)) + s4
    x = upsample(stack(sd, p + ".up3", x, st, M)) + s3
    x = upsample(stack(sd, p + ".up2", x, st, M)) + s2
    x = upsample(stack(sd, p + ".up1", x, st, M)) + s1
    return x


def hg_stem(sd, x, st):
    """_Hourglass_Wrapper.forward stem, :283-289."""
    x = F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"], stride=2, padding=3)   # :223,283
    x = F.relu(batchnorm(sd, "bn1", x, st))
    x = residual(sd, "residual1", x, st)
    x = maxpool(x)
    x = residual(sd, "residual2", x, st)
    x = residual(sd, "residual3", x, st)
    return x


def dropout(x, masks):
    """_Hourglass._dropout, :79-100: x * nearest-upsampled [N,1,4,4] masks."""
    scale = x.size(2) // 4
    if scale != 1:
        masks = masks.repeat_interleave(scale, dim=2).repeat_interleave(scale, dim=3)
    return x * masks.to(x.dtype).expand(x.size())


def sample_mask(pred_masks, dropout_num=2):
    """_Hourglass._sample_mask, :102-136: softmax over the 16 cells, np.random.choice(16, 2, replace=False) per sample
    from the GLOBAL numpy RandomState (seed it before calling), chosen cells zeroed in an all-ones mask."""
    import numpy as np
    n, _, h, w = pred_masks.shape
    masks = torch.ones(pred_masks.size(), dtype=pred_masks.dtype)
    probs = F.softmax(pred_masks.view(n, -1).float(), dim=1).detach().cpu().numpy()
    indexes = torch.zeros(n, dropout_num).long()
    for i in range(n):
        idx = np.random.choice(h * w, dropout_num, p=probs[i], replace=False)
        for j in range(len(idx)):
            masks[i, 0, int(idx[j]) // w, int(idx[j]) % w] = 0
            indexes[i, j] = int(idx[j])
    return masks, indexes


def hg_forward_dropout(sd, x, num_stacks, asn_sd, num_modules=1, training=True, asn_training=False, is_half_hg=False):
    """_Hourglass_Wrapper.forward in is_dropout mode, :308-322,340: half -> pred_mask; whole ->
    (outs, pred_mask, indexes, masks) with the masks of stack 0 re-applied in every later stack."""
    st = BNState(training)
    x = hg_stem(sd, x, st)
    outs, masks, pred, indexes = [], None, None, None
    for i in range(num_stacks):
        p = "hg.%d" % i
        neck, s1, s2, s3, s4 = hourglass_down(sd, p, x, st, num_modules)
        if i == 0:
            feats = {"neck": neck.detach(), "skip1": s1.detach(), "skip2": s2.detach(),
                     "skip3": s3.detach(), "skip4": s4.detach()}
            pred = asn_forward(asn_sd, feats, training=asn_training)[0]
            if is_half_hg:
                return pred, st
            masks, indexes = sample_mask(pred)
        neck, s1, s2, s3, s4 = (dropout(t, masks) for t in (neck, s1, s2, s3, s4))
        y = hourglass_up(sd, p, neck, s1, s2, s3, s4, st, num_modules)
        y = stack(sd, "post_res.%d" % i, y, st, num_modules)
        y = F.relu(batchnorm(sd, "linear.%d.1" % i, conv(sd, "linear.%d.0" % i, y, 1), st))
        o = conv(sd, "out_conv.%d" % i, y, 1)
        outs.append(o)
        if i < num_stacks - 1:
            x = x + conv(sd, "forth_conv.%d" % i, y, 1) + conv(sd, "in_conv.%d" % i, o, 1)
    return (outs, pred, indexes, masks), st


def hg_forward(sd, x, num_stacks, num_modules=1, training=True, asn_sd=None, is_half_hg=False):
    """_Hourglass_Wrapper.forward, models/asn_stacked_hg.py:282-342.

    asn_sd=None: returns (list of S heatmaps, BNState).
    asn_sd given (is_aug mode): half -> ((scale, rot), st); whole -> ((outs, scale, rot), st).
    The ASN runs on detached features (:161-162) with its own BN mode = asn_training False
    here (joint-train calls it under agent.eval(), joint-train-pose-s-r-agent.py:207-208).
    """
    st = BNState(training)
    x = hg_stem(sd, x, st)
    outs = []
    agent = None
    for i in range(num_stacks):
        p = "hg.%d" % i
        neck, s1, s2, s3, s4 = hourglass_down(sd, p, x, st, num_modules)
        if i == 0 and asn_sd is not None:
            feats = {"neck": neck.detach(), "skip1": s1.detach(), "skip2": s2.detach(),
                     "skip3": s3.detach(), "skip4": s4.detach()}                   # :161-162
            agent = asn_forward(asn_sd, feats, training=False)[0]
            if is_half_hg:
                return agent, st                                                    # :169-171, :302-304
        y = hourglass_up(sd, p, neck, s1, s2, s3, s4, st, num_modules)
        y = stack(sd, "post_res.%d" % i, y, st, num_modules)                        # :327
        y = F.relu(batchnorm(sd, "linear.%d.1" % i, conv(sd, "linear.%d.0" % i, y, 1), st))   # :328
        o = conv(sd, "out_conv.%d" % i, y, 1)                                       # :329
        outs.append(o)
        if i < num_stacks - 1:
            x = x + conv(sd, "forth_conv.%d" % i, y, 1) + conv(sd, "in_conv.%d" % i, o, 1)   # :331-334
    if agent is not None:
        return (outs, agent[0], agent[1]), st
    return outs, st


def asn_forward(sd, feats, training=True):
    """ASN.forward (is_aug=True), models/asn_stacked_hg.py:401-436."""
    st = BNState(training)
    skip1 = residual(sd, "residual_skip1", feats["skip1"], st)
    skip2 = residual(sd, "residual_skip2", feats["skip2"], st)
    skip3 = residual(sd, "residual_skip3", feats["skip3"], st)
    skip4 = residual(sd, "residual_skip4", feats["skip4"], st)
    neck = residual(sd, "residual_neck", feats["neck"], st)
    x = residual(sd, "merge1", maxpool(skip1) + skip2, st)
    x = residual(sd, "merge2", maxpool(x) + skip3, st)
    x = residual(sd, "merge3", maxpool(x) + skip4, st)
    x = residual(sd, "merge4", maxpool(x) + neck, st)
    x = stack(sd, "deep_merge", x, st, 3)
    if "fc_scale.weight" in sd:
        x = F.avg_pool2d(x, 4)                                                      # :375,431
        x = x.view(x.size(0), -1)
        scale = F.linear(x, sd["fc_scale.weight"], sd["fc_scale.bias"])             # :434
        rot = F.linear(x, sd["fc_rotation.weight"], sd["fc_rotation.bias"])         # :435
        return (scale, rot), st
    return F.conv2d(x, sd["out_conv.weight"], sd["out_conv.bias"]), st              # :438


# --------------------------------------------------------------------------------------
# Loss / optimiser / criterion
# --------------------------------------------------------------------------------------
def mse_loss(outs, target):
    """Inline loss of stack-hg.py:156-159: sum over stacks of mean squared error."""
    loss = 0
    for o in outs:
        t = (o - target) ** 2
        loss = loss + t.sum() / t.numel()
    return loss


def weighted_L2(pred, gt, weight):
    """pylib/Criterion.py:12-18."""
    loss = (pred - gt) ** 2 * weight
    return loss.sum() / loss.numel()


def weighted_sigmoid_crossentropy(pred, gt, weight):
    """pylib/Criterion.py:4-10."""
    loss = (gt * torch.log(pred + 1e-6) + (1 - gt) * torch.log(1 - pred + 1e-6)) * weight
    return -loss.sum() / loss.numel()


def agent_kl_loss(logits, target, num_bins):
    """joint-train-pose-s-r-agent.py:399-407: F.kl_div(log(softmax+1e-7), target) * bins
    (kl_div default reduction of torch 0.3 = mean over all elements)."""
    logp = torch.log(F.softmax(logits, dim=1) + 1e-7)
    pointwise = torch.where(target > 0, target * (torch.log(target.clamp_min(1e-30)) - logp),
                            torch.zeros_like(target))
    return pointwise.mean() * num_bins


def rmsprop_step(params, grads, square_avg, lr=2.5e-4, alpha=0.99, eps=1e-8):
    """torch.optim.RMSprop(lr, alpha=.99, eps=1e-8, momentum=0, weight_decay=0), stack-hg.py:51-52,165:
    v = alpha v + (1-alpha) g^2 ;  p -= lr * g / (sqrt(v) + eps)."""
    with torch.no_grad():
        for k in params:
            g = grads[k]
            square_avg[k].mul_(alpha).addcmul_(g, g, value=1 - alpha)
            params[k].addcdiv_(g, square_avg[k].sqrt().add_(eps), value=-lr)


# --------------------------------------------------------------------------------------
# Convenience: one full training step on a state dict (fwd + loss + bwd [+ RMSprop]).
# --------------------------------------------------------------------------------------
def is_trainable(name):
    return not (name.endswith("running_mean") or name.endswith("running_var")
                or name.endswith("num_batches_tracked"))


def train_step(sd, x, target, num_stacks, num_modules=1, square_avg=None, lr=2.5e-4):
    """stack-hg.py:153-165.  `sd` tensors are used in place; returns
    (outs, loss, grads dict, BNState).  If square_avg is given the RMSprop update and the
    running-stat updates are applied to `sd` as well."""
    leaves = OrderedDict()
    for k, v in sd.items():
        if is_trainable(k):
            leaves[k] = v.detach().requires_grad_(True)
    work = OrderedDict((k, leaves.get(k, v)) for k, v in sd.items())
    outs, st = hg_forward(work, x, num_stacks, num_modules, training=True)
    loss = mse_loss(outs, target)
    gs = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    grads = OrderedDict()
    for (k, v), g in zip(leaves.items(), gs):
        grads[k] = g if g is not None else torch.zeros_like(v)
    if square_avg is not None:
        params = OrderedDict((k, sd[k]) for k in leaves)
        rmsprop_step(params, grads, square_avg, lr=lr)
        with torch.no_grad():
            for k, v in st.updates.items():
                sd[k].copy_(v)
    return [o.detach() for o in outs], loss.detach(), grads, st
