"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python scripts/launch_summary.py launches.csv [skip_first_n_launches]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 1:][skip:]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    name = r[kn].split('(')[0].replace('<unnamed>::', '').replace('void ', '')[-60:]
    v = float(r[mv].replace(',', ''))
    v = v / 1000.0 if r[mu] in ('ns', 'nsecond') else v * 1000.0 if r[mu] in ('ms', 'msecond') else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print('| kernel | launches | total us | share |')
print('|---|---:|---:|---:|')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('| %s | %d | %.1f | %.1f%% |' % (k, len(v), sum(v), 100 * sum(v) / tot))
print('| **all** | %d | %.1f | 100%% |' % (sum(len(v) for v in agg.values()), tot))
