"""Per-kernel launch count and mean duration from an ncu --metrics gpu__time_duration.sum --csv launch list.
    python profiles/summarize_launches.py profiles/r2_launches_mhd_p_basic.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(d['Metric Value'].replace(',', ''))
    if d['Metric Unit'] in ('ns', 'nsecond'):
        v /= 1e3
    elif d['Metric Unit'] in ('ms', 'msecond'):
        v *= 1e3
    agg.setdefault(d['Kernel Name'][:90], []).append(v)
for k, v in agg.items():
    print('%-92s n=%4d  mean %9.2f us  total %10.1f us' % (k, len(v), sum(v) / len(v), sum(v)))
