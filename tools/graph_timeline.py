"""Timeline of ONE CUDA-graph replay of the config-2 train step (torch profiler / CUPTI kernel records):
where the step's wall time goes -- phases in which a machine-filling kernel runs vs phases in which only the
latency-bound small layers are in flight.  usage: python tools/graph_timeline.py [--streams 6 --low-streams 2]"""
import argparse, collections, json, os, sys, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from pose_adv_aug_b200 import synth
from pose_adv_aug_b200.models import asn_stacked_hg as M
from pose_adv_aug_b200 import HourglassTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=10)
ap.add_argument("--low-streams", type=int, default=3)
ap.add_argument("--out", default="gpurun_out/graph_timeline.json")
args = ap.parse_args()
dev = torch.device("cuda", 0)
net = M.create_hg(2, 1, 16, 256)
net.load_state_dict(synth.make_state_dict(synth.schema_of(net), seed=1, perturb_bn=False))
tr = HourglassTrainer(net, 24, 256, device=dev, use_graph=True, n_streams=args.streams, n_low=args.low_streams)
tr.x.copy_(synth.make_images(24, 256, seed=100)); tr.t.copy_(synth.make_heatmaps(24, 256, 16, seed=200))
for _ in range(4):
    tr.step_resident()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step_resident()
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        ev.append((e.time_range.start, e.time_range.end, e.name))
ev.sort()
if not ev:
    print("no CUDA kernel records"); sys.exit(0)
t0, t1 = ev[0][0], max(e[1] for e in ev)
print("kernels: %d, span %.3f ms, sum of kernel durations %.3f ms" % (len(ev), (t1 - t0) / 1e3, sum(e[1] - e[0] for e in ev) / 1e3))
def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("hgk::", "")
    return n[:48]
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ev:
    agg[short(n)][0] += 1; agg[short(n)][1] += (e - s)
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
    print("  %-48s %4d %9.1f us" % (n, c, t))
# concurrency profile: time with k kernels in flight
pts = []
for s, e, n in ev:
    pts.append((s, 1)); pts.append((e, -1))
pts.sort()
conc = collections.defaultdict(float); cur = 0; last = pts[0][0]
for t, d in pts:
    conc[cur] += t - last; last = t; cur += d
print("time with k kernels in flight (us):", {k: round(v, 1) for k, v in sorted(conc.items())})
# time during which at least one 'big' kernel (duration > 40 us) runs
big = sorted([(s, e) for s, e, n in ev if e - s > 40.0])
cov = 0.0; cs, ce = None, None
for s, e in big:
    if cs is None: cs, ce = s, e
    elif s <= ce: ce = max(ce, e)
    else: cov += ce - cs; cs, ce = s, e
if cs is not None: cov += ce - cs
print("time covered by kernels longer than 40 us: %.3f ms of %.3f ms" % (cov / 1e3, (t1 - t0) / 1e3))
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump([(s - t0, e - t0, short(n)) for s, e, n in ev], open(args.out, "w"))
