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
        st = self._sync()
        state = {}
        for i, (p, o) in enumerate(zip(st.params, st.offsets)):
            state[i] = {"square_avg": self.square_avg[o:o + p.numel()].view(p.shape).clone()}
        return {"state": state, "param_groups": [dict(self.param_groups[0], params=list(range(len(st.params))))]}

    def load_state_dict(self, sd):
        st = self._sync()
        for i, (p, o) in enumerate(zip(st.params, st.offsets)):
            s = sd["state"].get(i)
            if s is not None and "square_avg" in s:
                self.square_avg[o:o + p.numel()].view(p.shape).copy_(s["square_avg"])
        for k in ("lr", "alpha", "eps"):
            if k in sd["param_groups"][0]:
                self.param_groups[0][k] = sd["param_groups"][0][k]
