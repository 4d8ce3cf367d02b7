"""ctypes binding of libhgk.so (the C-ABI declared in include/hgk.h).

There is no CPU / PyTorch fallback: if the shared library is missing or a kernel call
fails, the product path raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhgk.so")

P, I, L, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float

# name -> argument types (the trailing `void* stream` included).  Mirrors include/hgk.h.
SIGNATURES = {
    "hgk_conv_nhwc": [P, P, P, I, I, I, I, I, P, I, I, P, I, P, P, P, I, P, I, P, P, I, P],
    "hgk_conv_wgrad_nhwc": [P, P, P, I, I, I, I, I, P, I, I, P, L, L, L, P, P],
    "hgk_pack_weights": [P, P, P, I, P],
    "hgk_conv_tc_nhwc": [P, P, P, I, I, I, I, I, P, P, I, P, I, P, P, P, I, P, I, P, P, P],
    "hgk_conv_tc_dgrad_bnstats_nhwc": [P, I, I, I, I, P, P, I, I, P, P, I, P, P, P, I, P, P, P, P, P],
    "hgk_conv_tc_bn_nhwc": [P, P, P, I, I, I, I, I, P, P, I, P, I, P, P, P, I, P, I, P, P,
                            P, P, F, F, P, P, P, P, P, P, P, P],
    "hgk_conv_tc_x2_nhwc": [P, P, P, I, I, I, I, I, P, P, I, P, I, P, P, P, I, P, I, P, P, P],
    "hgk_conv_tc_bn_x2_nhwc": [P, P, P, I, I, I, I, I, P, P, I, P, I, P, P, P, I, P, I, P, P,
                               P, P, F, F, P, P, P, P, P, P, P, P],
    "hgk_conv_tc_dgrad_bnfin_nhwc": [P, I, I, I, I, P, P, I, I, P, P, I, P, P, P, I, P, P, P, P,
                                     P, I, P, P, P, P, P, P, P],
    "hgk_conv_tc_dgrad_bnapply_nhwc": [P, P, P, P, I, P, P, P, P, P, I, I, I, I, P, I, I, P, P, I,
                                       P, P, P, I, P, P, P, P, P, I, P, P, P, P, P, P, P],
    "hgk_bn_bwd_reduce_fin": [P, P, P, P, I, P, P, L, I, P, P, P, I, P, P, P, P, P, P, P],
    "hgk_pack_weights_tc": [P, P, P, I, P],
    "hgk_conv_wgrad_tc_nhwc": [P, P, P, I, I, I, I, I, P, I, I, P, P, P],
    "hgk_unpack_add_grads": [P, P, P, I, P],
    "hgk_stem_conv7_fwd": [P, I, I, I, P, P, I, P, P, P, P],
    "hgk_stem_conv7_wgrad": [P, I, I, I, P, I, P, P, P],
    "hgk_stem_conv7_wgrad_bnapply": [P, I, I, I, P, P, P, P, I, P, P, P, P, I, P, P, P],
    "hgk_bn_finalize": [P, P, L, P, P, F, F, P, P, P, P, P, P, I, P],
    "hgk_bn_eval_prepare": [P, P, P, P, F, P, P, P, P, I, P],
    "hgk_bn_bwd_reduce": [P, P, P, P, I, P, P, L, I, P, P, P],
    "hgk_bn_bwd_finalize": [P, P, L, P, P, P, I, P, P, P, P, P, I, P],
    "hgk_bn_bwd_apply": [P, P, P, P, I, P, P, P, P, L, I, P],
    "hgk_maxpool2_fwd": [P, P, P, I, I, I, I, I, P, P],
    "hgk_maxpool2_bwd": [P, P, P, I, I, I, I, I, P, P, I, P],
    "hgk_add_fwd": [P, P, P, I, I, P, P, P, I, I, I, I, I, P, P],
    "hgk_upsample2_bwd": [P, I, I, I, I, P, I, P],
    "hgk_maxpool2_bwd_bnred": [P, P, P, I, I, I, I, I, P, P, I, P, P, P, P, P, I, P, P, P, P, P, P, P],
    "hgk_upsample2_bwd_bnred": [P, I, I, I, I, P, I, P, P, P, I, P, P, P, P, P, I, P, P, P, P, P, P, P],
    "hgk_add_into": [P, P, L, I, P],
    "hgk_head_combine_fwd": [P, P, P, P, P, P, P, P, I, I, P],
    "hgk_head_combine_bwd": [P, P, P, P, P, P, P, P, P, P, I, I, P],
    "hgk_nchw_to_nhwc": [P, I, I, I, I, P, P],
    "hgk_nhwc_to_nchw": [P, P, P, I, I, I, I, I, P, P],
    "hgk_avgpool_fwd": [P, P, P, I, I, I, I, I, I, P, P],
    "hgk_avgpool_bwd": [P, I, I, I, I, I, P, I, P],
    "hgk_linear_fwd": [P, P, P, I, I, I, P, P],
    "hgk_linear_bwd": [P, P, P, I, I, I, P, P, P, P],
    "hgk_mse_fwd_bwd": [P, P, L, F, F, P, I, P, P],
    "hgk_criterion_fwd": [I, P, P, P, L, P, P],
    "hgk_criterion_bwd": [I, P, P, P, L, P, P, P],
    "hgk_rmsprop_flat": [P, P, P, L, F, F, F, F, P],
    "hgk_rmsprop_flat_dev": [P, P, P, L, P, P],
    "hgk_f64_to_f32": [P, P, I, F, P],
    "hgk_heatmap_peaks": [P, I, I, I, I, I, I, I, P, P, P, P],
    "hgk_pts2heatmap": [P, I, I, I, P, I, P, P, P],
    "hgk_pck_accuracy": [P, P, P, I, I, F, F, P, I, P, P, P],
    "hgk_dist_acc": [P, I, F, P, P],
    "hgk_per_person_pckh": [P, P, I, I, P, I, F, P, P],
    "hgk_flip_merge_nchw": [P, P, I, I, I, I, P, I, I, P, P],
    "hgk_softmax_sample": [P, I, I, P, P, P, P],
    "hgk_mask_mul_fwd": [P, P, P, I, P, I, I, I, I, I, I, P, P],
    "hgk_mask_mul_bwd": [P, P, I, I, I, I, I, I, P, I, P],
    "hgk_stem_s2d_image": [P, I, I, I, P, P],
    "hgk_stem_s2d_weight": [P, I, P, P],
    "hgk_aug_minmax": [P, I, I, I, I, I, I, I, I, P, P, P],
    "hgk_aug_window_bytes": [P, I, I, I, I, I, I, I, I, I, P, I, I, P, P],
    "hgk_aug_image_bytes_f32": [P, I, I, P, P, P],
    "hgk_aug_resample_coeffs": [I, I, I, P, P, P],
    "hgk_aug_resize_h": [P, I, I, I, I, I, I, I, P, P, I, P, P],
    "hgk_aug_resize_v": [P, I, I, I, P, P, I, P, P],
    "hgk_aug_rotate": [P, I, I, P, P, P],
    "hgk_aug_to_chw_float": [P, I, I, P, P],
    "hgk_aug_crop_batch": [P, P, P, I, I, P, P, P, P, P, P],
}


class HGKError(RuntimeError):
    pass


class _Lib(object):
    def __init__(self, path):
        if not os.path.exists(path):
            raise HGKError(
                "libhgk.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the CUDA path)" % path)
        self.path = path
        self.cdll = ctypes.CDLL(path)
        self.cdll.hgk_last_error.restype = ctypes.c_char_p
        self.cdll.hgk_last_error.argtypes = []
        self.cdll.hgk_version.restype = I
        self.cdll.hgk_device_ok.restype = I
        self.cdll.hgk_pdl_arm.restype = I
        self.cdll.hgk_pdl_arm.argtypes = [I]
        self.pdl_arm = self.cdll.hgk_pdl_arm
        self.cdll.hgk_conv_tc_supported.restype = I
        self.cdll.hgk_conv_tc_supported.argtypes = [I, I, I]
        self.cdll.hgk_conv_tc_bnapply_supported.restype = I
        self.cdll.hgk_conv_tc_bnapply_supported.argtypes = [I, I, I, I, I, I]
        self.cdll.hgk_conv_tc_x2_supported.restype = I
        self.cdll.hgk_conv_tc_x2_supported.argtypes = [I, I, I, I, I, I]
        self.cdll.hgk_conv_wgrad_tc_supported.restype = I
        self.cdll.hgk_conv_wgrad_tc_supported.argtypes = [I, I, I]
        self.cdll.hgk_aug_resample_ksize.restype = I
        self.cdll.hgk_aug_resample_ksize.argtypes = [I, I]
        self.aug_resample_ksize = self.cdll.hgk_aug_resample_ksize
        self.cdll.hgk_aug_desc_fields.restype = I
        self.cdll.hgk_aug_desc_fields.argtypes = []
        for name, args in SIGNATURES.items():
            fn = getattr(self.cdll, name)      # AttributeError if the symbol is not exported
            fn.argtypes = args
            fn.restype = I
            setattr(self, name[4:], fn)

    def conv_tc_supported(self, cin, cout, k):
        return bool(self.cdll.hgk_conv_tc_supported(cin, cout, k))

    def conv_tc_bnapply_supported(self, n, h, w, cin, cout, k):
        return bool(self.cdll.hgk_conv_tc_bnapply_supported(n, h, w, cin, cout, k))

    def conv_tc_x2_supported(self, n, h, w, cin, cout, k):
        return bool(self.cdll.hgk_conv_tc_x2_supported(n, h, w, cin, cout, k))

    def conv_wgrad_tc_supported(self, cin, cout, k):
        return bool(self.cdll.hgk_conv_wgrad_tc_supported(cin, cout, k))

    def last_error(self):
        return self.cdll.hgk_last_error().decode()

    def check(self, rc, name="hgk"):
        if rc != 0:
            raise HGKError("%s failed (%d): %s" % (name, rc, self.last_error()))


_lib = None


def get_lib():
    global _lib
    if _lib is None:
        _lib = _Lib(LIB_PATH)
    return _lib
