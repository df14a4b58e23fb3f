"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv, collections, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
tot = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = row['Kernel Name']
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)
    short = name.split('(')[0]
    short = re.sub(r'^void ', '', short)
    if 'wu::' not in short:
        short = re.sub(r'<.*', '', short)
    tot.setdefault(short, [0, 0.0])
    tot[short][0] += 1
    tot[short][1] += v
    n += 1
s = sum(v[1] for v in tot.values())
ours = sum(v[1] for k, v in tot.items() if 'wu::' in k)
print(f"{n} launches, {s/1e3:.2f} ms of kernel time (ncu: cold cache, serialised); wu:: kernels {ours/1e3:.2f} ms = {100*ours/s:.1f}%")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t/1e3:9.3f} ms {100*t/s:5.1f}%  x{c:4d}  {k[:120]}")
