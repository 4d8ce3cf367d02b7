"""Which kernels run ALONE in a CUDA-graph replay (from the JSON written by tools/graph_timeline.py): time during which
exactly one kernel is in flight, attributed to that kernel and summed per kernel name -- the serial part of the step.
usage: python tools/alone_time.py gpurun_out/<tag>_graph_timeline.json"""
import collections, json, sys
ev = json.load(open(sys.argv[1]))
pts = []
for i, (s, e, n) in enumerate(ev):
    pts.append((s, 1, i)); pts.append((e, -1, i))
pts.sort()
live = set(); last = pts[0][0]
alone = collections.defaultdict(float); cnt = collections.defaultdict(int)
tot = collections.defaultdict(float)
for s, e, n in ev:
    tot[n] += e - s; cnt[n] += 1
for t, d, i in pts:
    if len(live) == 1:
        alone[ev[next(iter(live))][2]] += t - last
    last = t
    if d > 0: live.add(i)
    else: live.discard(i)
print("span %.1f us; alone time total %.1f us" % (max(e for s, e, n in ev) - ev[0][0], sum(alone.values())))
print("%-52s %5s %10s %10s" % ("kernel", "n", "alone us", "total us"))
for n, a in sorted(alone.items(), key=lambda x: -x[1])[:30]:
    print("%-52s %5d %10.1f %10.1f" % (n, cnt[n], a, tot[n]))
