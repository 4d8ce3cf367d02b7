"""CPU checks of the C-ABI: libhgk.so loads without a GPU, exports every symbol include/hgk.h declares,
argument validation works without launching, and the product path refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hgk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hgk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pose_adv_aug_b200._lib import get_lib, SIGNATURES, LIB_PATH
    assert os.path.exists(LIB_PATH), "build with __graft_entry__.build()"
    lib = get_lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib.cdll, n), "symbol %s declared in include/hgk.h is not exported" % n
    # every entry point that launches work is bound with a signature
    launching = [n for n in names if n not in ("hgk_last_error", "hgk_version", "hgk_device_ok", "hgk_conv_tc_supported",
                                               "hgk_conv_wgrad_tc_supported", "hgk_debug_set_timeline",
                                               "hgk_conv_tc_bnapply_supported", "hgk_pdl_arm", "hgk_conv_tc_x2_supported",
                                               "hgk_aug_resample_ksize", "hgk_aug_desc_fields")]
    assert sorted(launching) == sorted(SIGNATURES.keys())
    assert lib.cdll.hgk_version() >= 100


def test_plan_with_tensor_core_stem(monkeypatch):
    """HGK_STEM_TC=1: the stem is planned as space-to-depth + the image-tile tensor-core kernel with ksize = 4, fed by its own
    one-entry weight repack in front of the global one (so that it does not wait for it)."""
    from pose_adv_aug_b200.engine import Plan, ParamStore, schedule_streams
    from pose_adv_aug_b200.models import asn_stacked_hg as M
    monkeypatch.setenv("HGK_STEM_TC", "1")
    net = M.create_hg(1, 1, 16, 128)
    dev = torch.device("cpu")
    st = ParamStore(net, dev)
    plan = Plan([st], dev, True, True)
    img = plan.input_image(2, 256, 256)
    outs, _ = net._build(plan, img)
    for o in outs:
        plan.output_nchw(o, no_grad=True)
    plan.finish()
    names = [r[2] for r in plan.fwd]
    stem = [r for r in plan.fwd if r[2] in ("conv_tc_bn_x2_nhwc", "conv_tc_bn_nhwc") and r[1][10] == 4]
    assert len(stem) == 1 and stem[0][1][7] == 16 and stem[0][1][12] == 64 and names.count("stem_s2d_image") == 1
    assert names.index("stem_s2d_image") < plan.fwd.index(stem[0]) and names.count("stem_conv7_fwd") == 0
    assert names.count("bn_finalize") == 0
    pre = [r[2] for r in plan.pre]
    assert pre.index("stem_s2d_weight") < pre.index("pack_weights_tc") and id(stem[0]) in plan.no_pack_dep
    own = [r for r in plan.pre if r[2] == "pack_weights_tc"][0]
    assert stem[0][1][8] in plan.rw_override[id(own)][1] and stem[0][1][9] in plan.rw_override[id(own)][1]


def test_argument_validation_without_gpu():
    from pose_adv_aug_b200._lib import get_lib
    lib = get_lib()
    buf = (ctypes.c_float * 16)()
    p = ctypes.addressof(buf)
    assert lib.conv_nhwc(p, 0, 0, 0, 1, 1, 1, 6, p, 1, 0, 0, 4, 0, 0, 0, 0, p, 0, 0, 0, 0, 0) == -1
    assert "multiples of 4" in lib.last_error()
    assert lib.conv_tc_nhwc(p, 0, 0, 0, 1, 1, 1, 48, p, 0, 1, 0, 64, 0, 0, 0, 0, p, 0, 0, 0, 0) == -1
    assert "unsupported shape" in lib.last_error()
    assert lib.maxpool2_fwd(p, 0, 0, 0, 1, 3, 3, 4, p, 0) == -1
    assert lib.stem_conv7_fwd(p, 1, 64, 64, p, 0, 32, p, 0, 0, 0) == -1
    assert lib.rmsprop_flat(0, p, p, 4, 1e-3, 0.99, 1e-8, 1.0, 0) == -1
    assert lib.conv_tc_supported(128, 128, 3) and not lib.conv_tc_supported(16, 256, 1)
    assert lib.conv_wgrad_tc_supported(256, 16, 1) and not lib.conv_wgrad_tc_supported(16, 256, 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_product_path_fails_loudly_without_cuda():
    from pose_adv_aug_b200 import create_hg, HGKError
    from pose_adv_aug_b200.pylib.Criterion import weighted_L2
    net = create_hg(1, 1, 16, 32)
    with pytest.raises(HGKError):
        net(torch.rand(1, 3, 64, 64))
    with pytest.raises(HGKError):
        weighted_L2(torch.rand(4), torch.rand(4), torch.ones(4))


def test_module_schema_matches_reference_on_cpu():
    """state_dict keys/shapes of the drop-in modules == the reference's (golden meta), no GPU needed."""
    import json
    from pose_adv_aug_b200 import create_hg, create_asn, Hourglass
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "meta.json")))
    net = create_hg(2, 1, 16, 256)
    assert [[k, list(v.shape)] for k, v in net.state_dict().items()] == meta["schema_hg_s2_m1_k16_c256"]
    assert sum(p.numel() for p in net.parameters()) == meta["n_params_hg_s2_c256"]
    asn = create_asn(256, 256, 7, 7, is_aug=True)
    assert [[k, list(v.shape)] for k, v in asn.state_dict().items()] == meta["schema_asn_aug_c256"]
    asn_d = create_asn(256, 256, is_dropout=True)
    assert [[k, list(v.shape)] for k, v in asn_d.state_dict().items()] == meta["schema_asn_dropout_c256"]
    assert [k for k, _ in Hourglass(2, 256).state_dict().items()] == [k for k, _ in net.state_dict().items()]


def test_plan_builds_and_covers_every_parameter():
    """Dry plan construction on CPU tensors: launch counts of the config-2 step and that every parameter's
    gradient pointer is written by some backward launch."""
    from pose_adv_aug_b200.engine import Plan, ParamStore
    from pose_adv_aug_b200.models import asn_stacked_hg as M
    net = M.create_hg(2, 1, 16, 256)
    dev = torch.device("cpu")
    st = ParamStore(net, dev)
    plan = Plan([st], dev, True, True)
    img = plan.input_image(2, 256, 256)
    tgt = plan.target_nchw(2, 16, 64, 64)
    acc = torch.zeros(1, dtype=torch.float64)
    outs, _ = net._build(plan, img)
    for o in outs:
        plan.mse_loss(o, tgt, acc)
        plan.output_nchw(o, no_grad=True)
    plan.finish()
    names = [r[2] for r in plan.fwd]
    # 101 convolutions in the reference graph; the inter-stack in_conv is folded into forth_conv (Plan.head_comb)
    n_conv = 100 if M.FUSE_HEAD else 101
    n_x2 = names.count("conv_tc_bn_x2_nhwc")          # large 3x3 / 1x1 layers: TF32 + 2xBF16 products
    assert (names.count("conv_tc_nhwc") + names.count("conv_tc_x2_nhwc") + names.count("conv_tc_bn_nhwc") + n_x2 + names.count("conv_nhwc") == n_conv
            and names.count("stem_conv7_fwd") == 1)
    assert [r[2] for r in plan.pre] == (["head_combine_fwd"] if M.FUSE_HEAD else [])
    # 96 BatchNorms: 95 finalised by the last CTA of their tensor-core convolution, the stem's by its own launch
    assert names.count("conv_tc_bn_nhwc") + n_x2 == 95 and names.count("bn_finalize") == 1 and n_x2 >= 10
    assert names.count("maxpool2_fwd") == 9 and names.count("add_fwd") == 8
    bnames = [r[2] for r in plan.bwd]
    assert bnames.count("conv_wgrad_tc_nhwc") + bnames.count("conv_wgrad_nhwc") == n_conv
    assert bnames.count("head_combine_bwd") == (1 if M.FUSE_HEAD else 0)
    assert "add_into" not in bnames                                             # all gradient fan-ins are aliased/fused
    # 96 BatchNorm backwards: the apply is evaluated on load by the image-tile data-gradient kernels on the large layers;
    # the stem's by its weight-gradient kernel; a bn_bwd_apply launch remains for the small layers, which take the cluster
    # split-K kernel (at this dry plan's batch of 2 the 16x16 .. 4x4 rungs are small: <= 12 tiles of 128 pixels)
    n_ap = bnames.count("conv_tc_dgrad_bnapply_nhwc") + bnames.count("stem_conv7_wgrad_bnapply")
    assert n_ap + bnames.count("bn_bwd_apply") == 96 and n_ap >= 20 and bnames.count("stem_conv7_wgrad_bnapply") == 1
    # every BN-backward finaliser rides on the kernel that produced its sums
    n_ap_red = sum(1 for r in plan.bwd if r[2] == "conv_tc_dgrad_bnapply_nhwc" and r[1][20] != 0)
    n_pw_red = bnames.count("maxpool2_bwd_bnred") + bnames.count("upsample2_bwd_bnred")     # ... or on the pool / up-sample backward
    assert bnames.count("conv_tc_dgrad_bnfin_nhwc") + bnames.count("bn_bwd_reduce_fin") + n_ap_red + n_pw_red == 96
    assert n_pw_red >= 10 and bnames.count("bn_bwd_reduce_fin") <= 10
    assert "bn_bwd_finalize" not in bnames and "bn_bwd_reduce" not in bnames
    written = set()
    for fn, a, name in plan.bwd:
        written.update(x for x in a if isinstance(x, int))
    base = st.grad.data_ptr()
    direct = 0
    for i, p in enumerate(st.params):
        if st.grad_ptr(i) in written:
            direct += 1
    # 3x3 conv weights are reached through the tap-major scratch + unpack table, everything else directly
    assert direct + len(plan.wg_entries) == len(st.params)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package (nor the GPU arm of bench.py before its
    cpu_baseline leg) may import it."""
    pkg = os.path.join(ROOT, "pose_adv_aug_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
    bench = open(os.path.join(ROOT, "bench.py")).read()
    ours = bench[bench.index("def run_ours"):bench.index("def parity_vs_cpu")]
    assert "oracle" not in ours.split("cpu_baseline leg")[0]
