"""Per-launch CUDA-event timing of one real train step (no profiler, real clocks).
usage: python tools/time_plan.py [--batch 24] [--top 25]"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pose_adv_aug_b200 import synth                           # noqa: E402
from pose_adv_aug_b200.models import asn_stacked_hg as M      # noqa: E402
from pose_adv_aug_b200 import HourglassTrainer                # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=24)
ap.add_argument("--res", type=int, default=256)
ap.add_argument("--chan", type=int, default=256)
ap.add_argument("--stacks", type=int, default=2)
ap.add_argument("--top", type=int, default=25)
ap.add_argument("--conv-path", type=int, default=0)
ap.add_argument("--filter", default="")
args = ap.parse_args()
M.CONV_PATH = args.conv_path
dev = torch.device("cuda", 0)
net = M.create_hg(args.stacks, 1, 16, args.chan)
net.load_state_dict(synth.make_state_dict(synth.schema_of(net), seed=1, perturb_bn=False))
tr = HourglassTrainer(net, args.batch, args.res, device=dev, use_graph=False)
tr.x.copy_(synth.make_images(args.batch, args.res, seed=100))
tr.t.copy_(synth.make_heatmaps(args.batch, args.res, 16, seed=200))
for _ in range(3):
    tr.step_resident()
torch.cuda.synchronize()
plan = tr.plan
stream = torch.cuda.current_stream().cuda_stream
recs = plan.head_launches() + plan.fwd + plan.bwd
for a in plan.aux_zero:
    a.zero_()
plan.stat_f[:plan.stat_f_used].zero_()
plan.stat_b[:plan.stat_b_used].zero_()
tr.store.grad.zero_()
if plan.wg_buf is not None:
    plan.wg_buf.zero_()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(recs) + 1)]
evs[0].record()
for i, (fn, a, name) in enumerate(recs):
    rc = fn(*a, stream)
    assert rc == 0, name
    evs[i + 1].record()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
rows = []
for i, (fn, a, name) in enumerate(recs):
    ms = evs[i].elapsed_time(evs[i + 1])
    key = name
    if name in ("conv_tc_dgrad_bnfin_nhwc", "conv_tc_dgrad_bnstats_nhwc"):
        key = "%s k%d" % (name, a[7])
        desc = "%dx%d %d->%d" % (a[2], a[3], a[4], a[8])
    elif name == "conv_tc_dgrad_bnapply_nhwc":
        key = "%s k%d" % (name, a[15])
        desc = "%dx%d %d->%d%s" % (a[11], a[12], a[13], a[16], " +red" if a[20] else "")
    elif name in ("conv_nhwc", "conv_tc_nhwc", "conv_tc_bn_nhwc", "conv_tc_bn_x2_nhwc", "conv_tc_x2_nhwc"):
        N, H, W, Cin = a[4:8]
        if name != "conv_nhwc":
            k, Cout, split = a[10], a[12], a[9] != 0
            key = "%s k%d %s" % (name, k, "3xtf32" if split else "tf32")
        else:
            k, Cout = a[9], a[12]
            key = "%s k%d" % (name, k)
        desc = "%dx%d %d->%d" % (H, W, Cin, Cout)
    elif name in ("conv_wgrad_nhwc", "conv_wgrad_tc_nhwc"):
        N, H, W, Cin = a[4:8]
        Cout, k = a[9], a[10]
        key = "%s k%d" % (name, k)
        desc = "%dx%d %d->%d" % (H, W, Cin, Cout)
    else:
        desc = ""
    agg[key][0] += 1
    agg[key][1] += ms
    rows.append((ms, key, desc))
tot = sum(v[1] for v in agg.values())
print("step (event-timed, launches serialized on one stream): %.2f ms, %d launches" % (tot, len(recs)))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:args.top]:
    print("%-34s %4d %9.3f ms %5.1f%%" % (k, c, t, 100 * t / tot))
if args.filter:
    print("--- launches matching", args.filter)
    seen = collections.OrderedDict()
    for ms, key, desc in rows:
        if args.filter in key:
            seen.setdefault((key, desc), []).append(ms)
    for (key, desc), v in seen.items():
        print("%-30s %-22s n=%2d avg %7.1f us" % (key, desc, len(v), 1e3 * sum(v) / len(v)))
# critical path of the dependency DAG with these per-launch times (what infinitely many streams could reach)
from pose_adv_aug_b200.engine import _WRITES
PTR_MIN = 1 << 32
last_write, readers, barrier = {}, {}, -1
finish, pred = [], []
for i, (fn, a, name) in enumerate(recs):
    ms = evs[i].elapsed_time(evs[i + 1])
    wpos = _WRITES.get(name)
    rd = [x for j, x in enumerate(a) if isinstance(x, int) and x >= PTR_MIN and wpos is not None and j not in wpos]
    wr = [x for j, x in enumerate(a) if isinstance(x, int) and x >= PTR_MIN and (wpos is None or j in wpos)]
    deps = set()
    if wpos is None:
        deps = set(range(i))
    else:
        for p_ in rd:
            if p_ in last_write:
                deps.add(last_write[p_])
        for p_ in wr:
            if p_ in last_write:
                deps.add(last_write[p_])
            deps.update(readers.get(p_, ()))
        if barrier >= 0:
            deps.add(barrier)
    start = max([finish[d] for d in deps], default=0.0)
    pred.append(max(deps, key=lambda d: finish[d]) if deps else -1)
    finish.append(start + ms)
    for p_ in rd:
        readers.setdefault(p_, []).append(i)
    for p_ in wr:
        last_write[p_] = i
        readers[p_] = []
    if wpos is None:
        barrier = i
print("critical path of the launch DAG: %.2f ms (sum of launches %.2f ms)" % (max(finish), tot))
cp = collections.defaultdict(lambda: [0, 0.0])
i = max(range(len(finish)), key=lambda j: finish[j])
while i >= 0:
    cp[recs[i][2]][0] += 1
    cp[recs[i][2]][1] += evs[i].elapsed_time(evs[i + 1])
    i = pred[i]
print("--- critical path by entry point")
for k, (c, t) in sorted(cp.items(), key=lambda x: -x[1][1]):
    print("  %-32s %4d %8.3f ms" % (k, c, t))
print("--- slowest launches")
rows.sort(reverse=True)
for r in rows[:20]:
    print("%8.3f ms  %-30s %s" % r)
