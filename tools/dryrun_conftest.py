"""Developer aid (no GPU here): run the -m gpu tests on CPU tensors with every libhgk call
executed for argument conversion / validation only (launch errors ignored), so Python-side
mistakes (wrong arity, bad names, shape logic) surface before GPU time is spent.
Numeric assertions are expected to fail; look for non-AssertionError exceptions.

    HGK_DRYRUN=1 python -m pytest tests -m gpu -q -p tools.dryrun_conftest 2>&1 | grep -v AssertionError
"""
import os
import types

import torch


def pytest_configure(config):
    if not os.environ.get("HGK_DRYRUN"):
        return
    import pose_adv_aug_b200._lib as L
    lib = L.get_lib()
    for name in L.SIGNATURES:
        fn = getattr(lib, name[4:])

        def make(fn):
            def wrapped(*a):
                rc = fn(*a)
                return 0 if rc == -2 else rc
            return wrapped
        setattr(lib, name[4:], make(fn))
    lib.cdll.hgk_device_ok = lambda: 1
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0, synchronize=lambda: None)
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.cuda.current_device = lambda: 0
    import pose_adv_aug_b200.models.asn_stacked_hg as M

    def chk(x):
        if x.dim() != 4:
            raise ValueError("dims")
    M._check_input = chk
    import pose_adv_aug_b200.pylib.Evaluation as EV
    import pose_adv_aug_b200.pylib.HumanAug as HA
    import pose_adv_aug_b200.agent as AG
    EV._cuda_f32 = lambda t, what: t.contiguous().float()
    HA._check = lambda t, what: t.contiguous().float()
    type(torch.empty(0)).is_cuda = property(lambda self: True) if False else type(torch.empty(0)).is_cuda
    _orig_ss = AG.softmax_sample

    def _ss(logits, u):
        class _Fake(torch.Tensor):
            is_cuda = True
        return _orig_ss(logits.as_subclass(_Fake), u)
    AG.softmax_sample = _ss
    import pose_adv_aug_b200.trainer as TR
    orig = TR.HourglassTrainer.__init__

    def init(self, net, batch, res, **kw):
        kw["device"] = torch.device("cpu")
        kw["use_graph"] = False
        orig(self, net, batch, res, **kw)
    TR.HourglassTrainer.__init__ = init


def pytest_collection_modifyitems(config, items):
    if not os.environ.get("HGK_DRYRUN"):
        return
    for item in items:
        item.own_markers = [m for m in item.own_markers if m.name != "skip"]
