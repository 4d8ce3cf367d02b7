"""CPU check of engine.schedule_streams: for the full config-2 step, every data dependency between two
launches (RAW / WAW / WAR on any device pointer) is enforced by stream order or by a chain of event waits."""
import torch

from pose_adv_aug_b200.engine import Plan, ParamStore, schedule_streams, _WRITES
from pose_adv_aug_b200.models import asn_stacked_hg as M


def _plan():
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
    keep = (torch.zeros(2, 3, 256, 256), torch.zeros(2, 16, 64, 64))
    plan.patch("image", keep[0].data_ptr())
    plan.patch("target", keep[1].data_ptr())
    return plan, keep, plan.head_launches() + plan.fwd + plan.bwd


def test_schedule_respects_every_dependency():
    plan, keep, L = _plan()
    for ns in (2, 4):
        so, cross = schedule_streams(L, ns)
        assert len(so) == len(L) and max(so) < ns
        # vector clocks: vc[i][k] = newest launch on stream k that is ordered before (or is) launch i
        vc, tail = [], [-1] * ns
        for i in range(len(L)):
            k = so[i]
            c = list(vc[tail[k]]) if tail[k] >= 0 else [-1] * ns
            for d in cross[i]:
                assert so[d] != k and d < i
                c = [max(x, y) for x, y in zip(c, vc[d])]
            c[k] = i
            vc.append(c)
            tail[k] = i
        # recompute true dependencies independently of the scheduler
        PTR_MIN = 1 << 32
        last_write, readers, barrier = {}, {}, -1
        barriers = ("pack_weights", "pack_weights_tc", "unpack_add_grads")
        n_dep = 0
        for i, (fn, args, name) in enumerate(L):
            wpos = _WRITES.get(name)
            rd = [a for j, a in enumerate(args) if isinstance(a, int) and a >= PTR_MIN and wpos is not None and j not in wpos]
            wr = [a for j, a in enumerate(args) if isinstance(a, int) and a >= PTR_MIN and (wpos is None or j in wpos)]
            deps = set()
            if name in barriers:
                deps = set(range(i))
            else:
                for p in rd:
                    if p in last_write:
                        deps.add(last_write[p])
                for p in wr:
                    if p in last_write:
                        deps.add(last_write[p])
                    deps.update(readers.get(p, ()))
                # (the stem convolution and the target transpose read neither packed-weight arena: they may start
                #  while the weight repack at the head of the step is still running)
                if barrier >= 0 and not (name in ("stem_conv7_fwd", "stem_s2d_image", "nchw_to_nhwc") and L[barrier][2].startswith("pack_weights")):
                    deps.add(barrier)
            for d in deps:
                if d == i:
                    continue
                n_dep += 1
                assert vc[i][so[d]] >= d, "launch %d (%s) not ordered after %d (%s)" % (i, name, d, L[d][2])
            for p in rd:
                readers.setdefault(p, []).append(i)
            for p in wr:
                last_write[p] = i
                readers[p] = []
            if name in barriers:
                barrier = i
        assert n_dep > 600      # sanity: the recomputed dependency set is not trivially empty
        # and there is real parallelism: no stream holds more than 60 % of the launches
        assert max(so.count(k) for k in range(ns)) < 0.6 * len(L) or ns == 2


def test_write_table_covers_plan_entry_points():
    plan, keep, L = _plan()
    names = set(r[2] for r in L)
    for n in names:
        assert n in _WRITES or n in ("pack_weights", "pack_weights_tc", "unpack_add_grads"), n


def test_deferred_skip_branches_are_low_priority_and_held_back():
    """The hourglass skip branches are built after the down chain, confined to the low-priority streams and ordered
    (by a scheduling-only edge) after the launch that ends the 32x32 rung -- and the data dependencies still hold."""
    assert M.DEFER_SKIPS
    plan, keep, L = _plan()
    assert plan.low_recs and plan.after
    ns, n_low = 8, 3
    so, cross = schedule_streams(L, ns, n_low=n_low, low_ids=plan.low_recs, after=plan.after,
                                 low_names=("conv_wgrad_tc_nhwc", "conv_wgrad_nhwc", "stem_conv7_wgrad"))
    index = dict((id(r), i) for i, r in enumerate(L))
    vc, tail = [], [-1] * ns
    for i in range(len(L)):
        k = so[i]
        c = list(vc[tail[k]]) if tail[k] >= 0 else [-1] * ns
        for d in cross[i]:
            c = [max(x, y) for x, y in zip(c, vc[d])]
        c[k] = i
        vc.append(c)
        tail[k] = i
    n_edges = 0
    for i, r in enumerate(L):
        if id(r) in plan.low_recs:
            assert so[i] >= ns - n_low, "skip-branch launch %d (%s) on a normal-priority stream" % (i, r[2])
        a = plan.after.get(id(r))
        if a is not None:
            d = index[a]
            assert d < i and vc[i][so[d]] >= d, "launch %d is not held back behind its anchor %d" % (i, d)
            n_edges += 1
    assert n_edges >= 2 * 3 * 3          # two hourglasses x (skip1, skip2, skip3) x three convolutions


def test_plan_structure_of_agent_and_dropout_modes_on_cpu():
    """Host logic only (no launches): the launch lists planned for the ASN aug / dropout modes have the structure the
    reference forward prescribes (models/asn_stacked_hg.py:159-190,308-322)."""
    C = 32
    net = M.create_hg(2, 1, 16, C)
    dev = torch.device("cpu")

    def build(asn, **kw):
        stores = [ParamStore(net, dev)] + ([ParamStore(asn, dev)] if asn is not None else [])
        plan = Plan(stores, dev, True, True)
        img = plan.input_image(2, 256, 256)
        outs, agent = net._build(plan, img, asn, **kw)
        for o in outs:
            plan.output_nchw(o)
        if agent is not None:
            for a in agent:
                (plan.output_plain if kw.get("is_dropout") else plan.output_rows)(a)
        plan.finish()
        return plan, outs, agent

    names = lambda lst: [r[2] for r in lst]
    # aug, half hourglass: no up path, two [N,7] logits, gradients only inside the agent
    aug = M.create_asn(C, C, 7, 7, is_aug=True)
    plan, outs, agent = build(aug, is_half_hg=True)
    assert outs == [] and len(agent) == 2 and (agent[0].N, agent[0].C) == (2, 7)
    assert names(plan.fwd).count("linear_fwd") == 2 and names(plan.fwd).count("avgpool_fwd") == 1
    assert "add_fwd" in names(plan.fwd) and "mask_mul_fwd" not in names(plan.fwd)
    # dropout, half hourglass: one [N,4,4,1] map of mask logits
    drop = M.create_asn(C, C, is_dropout=True)
    plan, outs, agent = build(drop, is_half_hg=True, is_dropout=True)
    assert outs == [] and len(agent) == 1 and (agent[0].H, agent[0].W, agent[0].C) == (4, 4, 1)
    assert "softmax_sample" not in names(plan.fwd)
    # dropout, whole net: softmax -> host sampling -> 5 masked tensors in EACH of the two hourglasses (ref:179-190,320-322)
    plan, outs, agent = build(drop, is_dropout=True)
    f = names(plan.fwd)
    assert len(outs) == 2 and f.count("softmax_sample") == 1 and f.count("host_sample_mask") == 1
    assert f.count("mask_mul_fwd") == 10 and names(plan.bwd).count("mask_mul_bwd") == 10
    assert f.index("softmax_sample") < f.index("host_sample_mask") < f.index("mask_mul_fwd")
    assert tuple(plan.mask_indexes.shape) == (2, 2)
    with torch.no_grad():
        assert M.create_asn(C, C, 7, 7, is_aug=True)({"x": 0}, is_aug=False, is_dropout=False) is None   # ref:430-439


def test_bucketed_gradient_unpack_and_allreduce_points():
    """In-graph bucketed all-reduce (trainer.py): with `grad_splits` the 3x3 weight-gradient unpack becomes one launch per
    bucket behind that bucket's last weight-gradient kernel, scheduled by its real read / write sets; every bucket of the
    flat gradient buffer has a last writer, and the buckets of the later layers are final well before the end of the
    backward pass (that is what the all-reduce overlaps with)."""
    from pose_adv_aug_b200.trainer import plan_bucket_splits, bucket_last_writers
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
    plan.finish(grad_splits=lambda pl: plan_bucket_splits(pl, st, 5))
    splits = plan.grad_splits
    assert len(splits) == 4 and all(s in st.offsets for s in splits) and splits == sorted(splits)
    sizes = [b - a for a, b in zip([0] + splits, splits + [st.numel])]
    keep = (torch.zeros(2, 3, 256, 256), torch.zeros(2, 16, 64, 64))
    plan.patch("image", keep[0].data_ptr())
    plan.patch("target", keep[1].data_ptr())
    L = plan.head_launches() + plan.fwd + plan.bwd
    unp = [i for i, r in enumerate(L) if r[2] == "unpack_add_grads"]
    assert len(unp) == 5 and all(id(L[i]) in plan.rw_override for i in unp)
    # every scratch range an unpack reads was written by earlier launches only, every row appears exactly once
    n_rows = 0
    for i in unp:
        rd, wr = plan.rw_override[id(L[i])]
        n_rows += len(rd)
        for p in rd:
            writers = [j for j, r in enumerate(L) if r[2].startswith("conv_wgrad") and p in r[1]]
            assert writers and max(writers) < i
    assert n_rows == len(plan.wg_entries)
    last = bucket_last_writers(L, plan.rw_override, st.grad.data_ptr(), st.numel, splits)
    assert all(l >= 0 for l in last)
    n_head, n_fwd = len(plan.head_launches()), len(plan.fwd)
    assert all(l >= n_head + n_fwd for l in last)                 # gradients are written by the backward list
    frac = [(l - n_head - n_fwd) / float(len(plan.bwd)) for l in last]
    # back-to-front completion: at least half of the bytes are final before 55 % of the backward list has been issued,
    # and the bucket that is final last (the stem side of the buffer) is a small one
    early = sum(sz for sz, f in zip(sizes, frac) if f < 0.55)
    assert early > 0.5 * st.numel, (sizes, frac)
    assert sizes[frac.index(max(frac))] < 0.15 * st.numel, (sizes, frac)
    # the scheduler orders each unpack after the weight-gradient launches it reads, without making it a barrier
    ns = 8
    so, cross = schedule_streams(L, ns, n_low=3, low_ids=plan.low_recs, after=plan.after, rw_override=plan.rw_override,
                                 low_names=("conv_wgrad_tc_nhwc", "conv_wgrad_nhwc", "stem_conv7_wgrad", "stem_conv7_wgrad_bnapply"))
    vc, tail = [], [-1] * ns
    for i in range(len(L)):
        k = so[i]
        c = list(vc[tail[k]]) if tail[k] >= 0 else [-1] * ns
        for d in cross[i]:
            c = [max(x, y) for x, y in zip(c, vc[d])]
        c[k] = i
        vc.append(c)
        tail[k] = i
    for i in unp:
        rd, wr = plan.rw_override[id(L[i])]
        for p in rd:
            if p < (1 << 32):
                continue        # schedule_streams tells pointers from sizes by magnitude (device pointers are > 2^32); a CPU
                                # heap block can lie below that when this test runs late in a long pytest process
            for j, r in enumerate(L[:i]):
                if r[2].startswith("conv_wgrad") and p in r[1]:
                    assert vc[i][so[j]] >= j
        # not a barrier: some launch issued before it is NOT ordered before it
        assert any(vc[i][so[j]] < j for j in range(i))


def test_weight_repack_has_real_dependencies_and_the_stem_is_issued_first():
    """Plan.step_launches(): target transpose and FFMA stem are issued in front of the two weight repacks, which are scheduled
    by their real read / write sets (derived weights in, packed blocks out) instead of as barriers: every launch that points
    at a packed block is ordered behind its repack, the stem is not."""
    plan, keep, _ = _plan()
    L = plan.step_launches()
    names = [r[2] for r in L]
    assert sorted(map(id, L)) == sorted(map(id, plan.head_launches() + plan.fwd + plan.bwd))
    i_stem, i_pack, i_tc = names.index("stem_conv7_fwd"), names.index("pack_weights"), names.index("pack_weights_tc")
    assert i_stem < i_pack < i_tc and names.index("head_combine_fwd") < i_pack
    ns = 8
    so, cross = schedule_streams(L, ns, n_low=3, low_ids=plan.low_recs, after=plan.after, rw_override=plan.rw_override,
                                 no_pack_ids=plan.no_pack_dep,
                                 low_names=("conv_wgrad_tc_nhwc", "conv_wgrad_nhwc", "stem_conv7_wgrad", "stem_conv7_wgrad_bnapply"))
    vc, tail = [], [-1] * ns
    for i in range(len(L)):
        k = so[i]
        c = list(vc[tail[k]]) if tail[k] >= 0 else [-1] * ns
        for d in cross[i]:
            c = [max(x, y) for x, y in zip(c, vc[d])]
        c[k] = i
        vc.append(c)
        tail[k] = i
    n_checked = 0
    for i_p in (i_pack, i_tc):
        rd, wr = plan.rw_override[id(L[i_p])]
        wr = set(p for p in wr if p >= (1 << 32))          # (pointers below 2^32 are not recognised as such, see above)
        for i, r in enumerate(L):
            if i != i_p and any(isinstance(a, int) and a in wr for a in r[1]):
                assert i > i_p and vc[i][so[i_p]] >= i_p, "launch %d (%s) reads a packed block but is not behind the repack" % (i, r[2])
                n_checked += 1
        # the repack itself waits for the derived weights it reads
        for j in range(i_p):
            if L[j][2] == "head_combine_fwd" and any(p in L[j][1] for p in rd if p >= (1 << 32)):
                assert vc[i_p][so[j]] >= j
    assert n_checked > 150 or not wr
    assert vc[i_stem][so[i_pack]] < i_pack if so[i_stem] != so[i_pack] else True


def test_bucket_split_dynamic_programme_is_optimal_on_small_cases():
    """plan_bucket_splits minimises sum(bucket elements x completion position of the bucket) over contiguous partitions:
    compared with brute force over every partition of small random instances."""
    import itertools
    import random
    from pose_adv_aug_b200 import trainer as T

    class _Store(object):
        pass

    class _Plan(object):
        pass

    rnd = random.Random(7)
    for trial in range(30):
        P = rnd.randint(2, 9)
        sizes = [rnd.randint(1, 50) * 32 for _ in range(P)]
        pos = [rnd.randint(-1, 40) for _ in range(P)]
        nb = rnd.randint(1, 4)
        st = _Store()
        st.offsets = [sum(sizes[:i]) for i in range(P)]
        st.numel = sum(sizes)
        plan = _Plan()
        plan.bwd = [None] * 41
        orig = T.param_completion
        T.param_completion = lambda pl, s_: pos
        try:
            cuts = T.plan_bucket_splits(plan, st, nb)
        finally:
            T.param_completion = orig
        assert cuts == sorted(cuts) and all(c in st.offsets[1:] for c in cuts) and len(cuts) <= nb - 1
        c = [(p + 1) / 41.0 for p in pos]

        def cost(cut_idx):
            edges = [0] + list(cut_idx) + [P]
            return sum(sum(sizes[a:b]) * max(c[a:b]) for a, b in zip(edges[:-1], edges[1:]))

        best = min(cost(ci) for k in range(0, nb) for ci in itertools.combinations(range(1, P), k))
        got = cost([st.offsets.index(x) for x in cuts])
        assert abs(got - best) < 1e-9 * max(best, 1.0), (sizes, pos, nb, cuts, got, best)
