"""pose_adv_aug_b200 -- B200-native (sm_100a) stacked-hourglass (+ASN agent) training path,
a drop-in for models/asn_stacked_hg.py and pylib/Criterion.py of zhiqiangdon/pose-adv-aug.

    from pose_adv_aug_b200.models.asn_stacked_hg import create_hg, create_asn
    from pose_adv_aug_b200.pylib.Criterion import weighted_L2, weighted_sigmoid_crossentropy
    from pose_adv_aug_b200 import FlatRMSprop, HourglassTrainer

Everything numerical runs in libhgk.so (hand-written CUDA, C-ABI in include/hgk.h); there is
no CPU fallback.
"""
from ._lib import get_lib, HGKError, LIB_PATH          # noqa: F401
from .models.asn_stacked_hg import create_hg, create_asn, Hourglass   # noqa: F401
from .optim import FlatRMSprop                          # noqa: F401
from .trainer import HourglassTrainer                   # noqa: F401
from . import dist                                      # noqa: F401
