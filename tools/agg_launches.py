"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for ONE step
(between two pack_weights_kernel launches).  usage: python tools/agg_launches.py file.csv [top_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
hdr, data = None, []
for r in rows:
    if r and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
idx = [i for i, d in enumerate(data) if 'pack_weights_kernel' in d['Kernel Name']]
seg = data[idx[0]:idx[1]] if len(idx) > 1 else data
agg = collections.defaultdict(lambda: [0, 0.0])
big = []
for d in seg:
    n = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('hgk::', '')
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[n][0] += 1
    agg[n][1] += v
    big.append((v, n, d['Grid Size']))
tot = sum(v[1] for v in agg.values())
print('one step: %d launches, %.1f us summed kernel time' % (len(seg), tot))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:topn]:
    print('%-44s %4d %10.1f us %5.1f%%' % (n[:44], c, t, 100 * t / tot))
big.sort(reverse=True)
print('--- slowest launches')
for b in big[:14]:
    print('%9.1f us  %-36s grid %s' % b)
