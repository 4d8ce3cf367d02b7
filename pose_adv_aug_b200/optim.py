"""Flat multi-tensor RMSprop: torch.optim.RMSprop(lr, alpha=.99, eps=1e-8, momentum=0,
weight_decay=0) of the reference (stack-hg.py:51-52,165) as ONE kernel launch over the flat
parameter / gradient / square-average buffers (the reference issues ~5 pointwise launches for
each of its 396 parameter tensors)."""
import torch

from ._lib import get_lib, HGKError
from .models.asn_stacked_hg import _ensure_store


class FlatRMSprop(object):
    def __init__(self, module, lr=2.5e-4, alpha=0.99, eps=1e-8, grad_scale=1.0, device=None):
        self.module = module
        self.lr, self.alpha, self.eps, self.grad_scale = lr, alpha, eps, grad_scale
        self.device = device
        self._store = None
        self.square_avg = None
        self.param_groups = [{"lr": lr, "alpha": alpha, "eps": eps}]     # adjust_lr-style access (utils/util.py:105)

    def _sync(self):
        dev = self.device
        if dev is None:
            p = next(self.module.parameters())
            dev = p.device
        if dev.type != "cuda":
            raise HGKError("FlatRMSprop needs the module on a CUDA device")
        store = _ensure_store(self.module, dev)
        if store is not self._store:
            old = self.square_avg
            self.square_avg = torch.zeros_like(store.flat)
            if old is not None and old.numel() == self.square_avg.numel():
                self.square_avg.copy_(old)
            self._store = store
        return store

    @property
    def store(self):
        return self._sync()

    def zero_grad(self, set_to_none=False):
        self._sync().grad.zero_()

    def step(self, grad_scale=None):
        st = self._sync()
        st.attach_grads()
        lib = get_lib()
        g = self.param_groups[0]
        stream = torch.cuda.current_stream(st.device).cuda_stream
        lib.check(lib.rmsprop_flat(st.flat.data_ptr(), st.grad.data_ptr(), self.square_avg.data_ptr(), st.numel,
                                   float(g["lr"]), float(g["alpha"]), float(g["eps"]),
                                   float(self.grad_scale if grad_scale is None else grad_scale), stream),
                  "hgk_rmsprop_flat")

    def state_dict(self):
        return pack_state(self._sync(), self.square_avg, self.param_groups[0])

    def load_state_dict(self, sd):
        unpack_state(self._sync(), self.square_avg, sd, self.param_groups[0])


def pack_state(store, square_avg, group):
    """torch.optim-style state dict of the flat square-average buffer: state[i]['square_avg'] per parameter i (the order
    of module.parameters()), param_groups[0] = hyper-parameters + 'params': [0..n-1]."""
    state = {}
    for i, (p, o) in enumerate(zip(store.params, store.offsets)):
        state[i] = {"square_avg": square_avg[o:o + p.numel()].view(p.shape).clone()}
    g = dict((k, group[k]) for k in ("lr", "alpha", "eps"))
    g["params"] = list(range(len(store.params)))
    return {"state": state, "param_groups": [g]}


def unpack_state(store, square_avg, sd, group):
    """Inverse of pack_state.  Saved state keys are mapped through param_groups[0]['params'] BY POSITION, as
    torch.optim.Optimizer.load_state_dict does (so checkpoints keyed by id(p), e.g. the reference's torch-0.3 files, load
    too); missing entries and shape mismatches raise instead of leaving square_avg silently at zero."""
    groups = sd.get("param_groups") or [{}]
    keys = groups[0].get("params")
    if keys is None:
        keys = sorted(sd["state"].keys(), key=lambda k: (str(type(k)), k))
    if len(keys) != len(store.params):
        raise HGKError("optimizer state has %d parameters, the model has %d" % (len(keys), len(store.params)))
    for key, p, o in zip(keys, store.params, store.offsets):
        s = sd["state"].get(key)
        if s is None:
            if sd["state"]:          # a partially filled state would silently reset some square averages
                raise HGKError("optimizer state has no entry for parameter key %r" % (key,))
            continue
        v = s.get("square_avg")
        if v is None:
            raise HGKError("optimizer state entry %r has no 'square_avg' (not an RMSprop checkpoint?)" % (key,))
        if v.numel() != p.numel():
            raise HGKError("square_avg of parameter key %r has %d elements, the parameter has %d" % (key, v.numel(), p.numel()))
        square_avg[o:o + p.numel()].view(p.shape).copy_(v.reshape(p.shape))
    for k in ("lr", "alpha", "eps"):
        if k in groups[0]:
            group[k] = groups[0][k]
